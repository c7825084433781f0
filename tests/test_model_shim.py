"""CPU: host-side mirror of the reference interface (flags, messages, seeding, wire format)."""
import importlib
import json

import numpy as np
import pandas as pd
import pytest

PKG = "covid19-vaccination-model_b200"


@pytest.fixture(scope="module")
def m():
    return importlib.import_module(PKG + ".model")


def test_cli_flags_match_reference(m):
    """model.py:419-500: same flags, same splitting, single value -> [v, v]."""
    p = m.build_parser()
    a = p.parse_args("--pro=35,45 --anti=12,25 --pressure=0.,0.02 --dupl_time=5,7 --init_stock=1,1.2 "
                     "--max_delivery=5,10 --date_range=2020-12-15,2022-07-1".split())
    assert a.pro == [35.0, 45.0] and a.anti == [12.0, 25.0] and a.pressure == [0.0, 0.02]
    assert a.dupl_time == [5.0, 7.0] and a.init_stock == [1.0, 1.2] and a.max_delivery == [5.0, 10.0]
    assert a.date_range == ["2020-12-15", "2022-07-1"] and a.mc_samples == 100 and a.CI == 0.95
    b = p.parse_args("--pro=40 --anti=20 --pressure=0.01 --dupl_time=6 --init_stock=1.1 --max_delivery=7 "
                     "--date_range=2021-01-01 --mc_samples=7 --CI=95".split())
    assert b.pro == [40.0, 40.0] and b.date_range == ["2021-01-01", "2021-01-01"]
    assert b.mc_samples == 7 and b.CI == 95.0
    for flag in ("--pro", "--anti", "--pressure", "--dupl_time", "--init_stock", "--max_delivery",
                 "--date_range"):
        assert [x for x in p._actions if flag in x.option_strings][0].required
    with pytest.raises(SystemExit):
        p.parse_args(["--pro=1,2"])


def test_agnostics_message_formatting(m):
    """model.py:339-349 incl. the collapse to one value when the interval is < 1 %."""
    assert m._agnostics_message(0.415, 0.0473) == "Agnosticts: 37 - 46%"
    assert m._agnostics_message(0.40, 0.0) == "Agnosticts: 40%"
    assert m._agnostics_message(0.40, 0.004) == "Agnosticts: 40%"
    assert m._agnostics_message(0.01, 0.05) == "Agnosticts: 0 - 6%"
    assert m.MSG_ERR_BOUNDS.startswith("ERROR: The pertentages of pro- and anti-vaccines")


def test_dates_match_reference_ranges(m):
    daily, weekly = m._dates("2020-12-15", "2022-07-1")
    assert len(daily) == 564 and len(weekly) == 81
    assert weekly[0] == daily[0] and (weekly[1] - weekly[0]).days == 7
    d5, w5 = m._dates("2020-12-15", "2025-12-14")
    assert len(d5) == 1826 and len(w5) == 261
    assert isinstance(weekly, pd.DatetimeIndex)


def test_seed_state_plays_np_random_seed(m):
    m.seed(12345)
    a = [m._next_seed(None) for _ in range(3)]
    m.seed(12345)
    b = [m._next_seed(None) for _ in range(3)]
    assert a == b and a[0] == (12345, True) and len({x[0] for x in a}) == 3
    assert m._next_seed(7) == (7, True)
    m.seed(None)
    s1, det = m._next_seed(None)
    assert not det
    assert m._next_seed(None)[0] != s1


def test_results_wire_format_round_trips_json(m):
    """app.py:240,383,483: model_results must survive the dcc.Store JSON round trip."""
    daily, weekly = m._dates("2021-01-01", "2021-01-20")
    res = {k: {"upper": np.arange(len(d)) + 1.0, "lower": np.arange(len(d)) - 1.0,
               "mean": np.arange(len(d)) * 1.0, "dates": d}
           for k, d in zip(m.SERIES, (daily, weekly, daily, daily))}
    res["number_finished_samples"] = 100
    back = json.loads(json.dumps(m.results_to_jsonable(res)))
    assert back["number_finished_samples"] == 100
    for k in m.SERIES:
        assert back[k]["mean"] == list(res[k]["mean"])
        assert pd.DatetimeIndex(back[k]["dates"]).equals(res[k]["dates"])
