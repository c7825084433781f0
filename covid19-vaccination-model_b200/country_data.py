"""Offline-safe counterpart of the reference's get_country_data (plot.py:8-36).

The reference downloads Our World in Data's vaccinations.csv when `plot` is imported by the
Dash app (app.py:30 calls get_country_data() at import time), which blocks -- or crashes --
app.py on an air-gapped box.  This version never touches the network unless it is given an
http(s) source, and with `offline_ok` it degrades to an EMPTY table of the same shape instead
of raising, so the model side of the app keeps working.

Same return value: (avail_countries, country_data) where country_data is the pivot table
plot_country_data reads (plot.py:160-203): columns (value, location) for the two observed
series, indexed by date; `country_data[k][country].dropna()` is what gets plotted.
"""
from __future__ import annotations

import os

import numpy as np
import pandas as pd

OWID_URL = ("https://raw.githubusercontent.com/owid/covid-19-data/master/public/data/vaccinations/"
            "vaccinations.csv")                      # plot.py:10-12
VALUE_COLUMNS = ["people_vaccinated_per_hundred", "daily_vaccinations_per_million"]   # plot.py:16-21
ENV_SOURCE = "VAXMC_COUNTRY_CSV"


def empty_country_data():
    """What get_country_data returns when no data can be had: no countries and a table without
    columns, so plot_country_data's `k in country_data.columns` (plot.py:169) is False for every
    series and it takes its "No data to show" branch instead of indexing a country that is not
    there (the app preselects Germany / United States / Russia, app.py:204)."""
    cols = pd.MultiIndex.from_arrays([[], []], names=[None, "location"])
    return np.array([], dtype=object), pd.DataFrame(index=pd.DatetimeIndex([], name="date"), columns=cols,
                                                    dtype=np.float64)


def get_country_data(source: str | os.PathLike | None = None, offline_ok: bool = True):
    """plot.py:8-36 on `source`: a local CSV path, an http(s) URL, or None.

    None means $VAXMC_COUNTRY_CSV if set, otherwise -- only when offline_ok is False -- the OWID
    URL the reference uses.  With offline_ok (the default) a missing/unreachable source yields
    empty_country_data() instead of an exception, and nothing is fetched unless an URL was asked
    for explicitly.
    """
    if source is None:
        source = os.environ.get(ENV_SOURCE)
    if source is None:
        if offline_ok:
            return empty_country_data()
        source = OWID_URL
    try:
        df = pd.read_csv(source)
    except Exception:
        if offline_ok:
            return empty_country_data()
        raise
    df["date"] = pd.to_datetime(df["date"])
    df = df.sort_values(["location", "date"])
    df = df.loc[:, ["location", "date"] + VALUE_COLUMNS]
    avail_countries = df["location"].unique()
    country_data = pd.pivot_table(df, columns="location", values=VALUE_COLUMNS, index="date")
    return avail_countries, country_data
