"""CPU: offline-safe get_country_data against the reference's own output.

tests/golden/country_data_pivot.csv was written by oracle/make_golden.py --country-data, which
runs the UNMODIFIED /root/reference/plot.py:get_country_data with pd.read_csv pointed at the
committed fixture tests/golden/owid_vaccinations_sample.csv (OWID's vaccinations.csv schema,
synthetic rows: there is no network here)."""
import importlib
import os
import socket

import numpy as np
import pandas as pd
import pytest

PKG = "covid19-vaccination-model_b200"
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FIXTURE = os.path.join(GOLDEN, "owid_vaccinations_sample.csv")


@pytest.fixture(scope="module")
def cd():
    return importlib.import_module(PKG + ".country_data")


def _golden():
    want = pd.read_csv(os.path.join(GOLDEN, "country_data_pivot.csv"), header=[0, 1], index_col=0)
    want.index = pd.to_datetime(want.index)
    countries = open(os.path.join(GOLDEN, "country_data_countries.txt")).read().split("\n")[:-1]
    return countries, want


def test_matches_reference_pivot_on_the_fixture(cd):
    countries, want = _golden()
    got_countries, got = cd.get_country_data(FIXTURE)
    assert list(got_countries) == countries
    assert list(got.columns) == list(want.columns)
    assert got.index.equals(want.index) and got.index.name == "date"
    assert np.array_equal(got.to_numpy(), want.to_numpy(), equal_nan=True)
    # what plot_country_data does with it (plot.py:169-175)
    for k in cd.VALUE_COLUMNS:
        assert k in got.columns
        g = got[k]["Andorra"].dropna()
        assert len(g) > 0 and g.index.is_monotonic_increasing
    assert "cum_number_vac_received_per_hundred" not in got.columns      # model-only series: 'No data to show'


def test_never_touches_the_network_by_default(cd, monkeypatch):
    def no_network(*a, **k):
        raise AssertionError("network access attempted")

    monkeypatch.setattr(socket, "create_connection", no_network)
    monkeypatch.setattr(socket.socket, "connect", no_network)
    monkeypatch.delenv(cd.ENV_SOURCE, raising=False)
    countries, data = cd.get_country_data()                  # what app.py:30 would call at import time
    assert len(countries) == 0 and data.empty
    for k in cd.VALUE_COLUMNS:
        assert k not in data.columns                          # plot.py:169 -> its "No data to show" branch
    monkeypatch.setenv(cd.ENV_SOURCE, FIXTURE)
    countries, data = cd.get_country_data()
    assert "United States" in countries and not data.empty


def test_missing_source_degrades_or_raises(cd, tmp_path):
    countries, data = cd.get_country_data(tmp_path / "nope.csv")
    assert len(countries) == 0 and data.empty
    with pytest.raises(Exception):
        cd.get_country_data(tmp_path / "nope.csv", offline_ok=False)
