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


# ------------------------------------------------------------------ result cache (N1)
class _FakeJob:
    """Stands in for distributed.run_job: counts calls, returns host buffers (no GPU)."""

    def __init__(self, m):
        self.m = m
        self.calls = 0

    def __call__(self, cfg, device=None):
        self.calls += 1
        out = self.m.nat.ResultBuffers(cfg.T)
        out.c.n_finished = cfg.n_samples if not cfg.max_running_time else cfg.n_samples // 2
        out.c.soft_no_mean, out.c.soft_no_std = 0.4, 0.05
        out.mean[0][:] = float(cfg.seed % 1000) + self.calls
        return out


USA_ARGS = ([35, 45], [12, 25], [0.0, 0.02], [5, 7], [1, 1.2], [5, 10])
USA_DATES = dict(start_date="2020-12-15", end_date="2021-01-31")


def test_result_cache_hit_miss_eviction(m, monkeypatch):
    """model.py:139 (lru_cache(maxsize=10) on run_sampling) + app.py:435-447: identical slider state
    with a fixed seed must re-plot without recomputing; anything that changes the sampled tuples
    or the reduction must miss."""
    from importlib import import_module

    dist_ = import_module(PKG + ".distributed")
    fake = _FakeJob(m)
    monkeypatch.setattr(dist_, "run_job", fake)
    m.cache_clear()
    base = dict(CI=95, n_rep=1000, N=50000, date_range=USA_DATES, seed=12345)

    def call(args=USA_ARGS, **kw):
        k = {**base, **kw}
        return m.run_model(*args, k.pop("CI"), k.pop("n_rep"), k.pop("N"), k.pop("date_range"), **k)

    r1 = call()
    assert fake.calls == 1
    r2 = call()
    assert fake.calls == 1 and r2 is r1                                   # hit: the very same triple
    # one miss per key field: each bound, CI, n_rep, N, either date, the seed, the mode
    for i in range(6):
        for j in range(2):
            a = [list(x) for x in USA_ARGS]
            a[i][j] = a[i][j] * 1.01 + 0.001
            before = fake.calls
            call(args=tuple(a))
            assert fake.calls == before + 1, (i, j)
    m.cache_clear()
    call()
    n0 = fake.calls
    for kw in (dict(CI=90), dict(n_rep=1001), dict(N=50001), dict(seed=12346), dict(mode="expected"),
               dict(date_range=dict(start_date="2020-12-16", end_date="2021-01-31")),
               dict(date_range=dict(start_date="2020-12-15", end_date="2021-02-01")),
               dict(direct_max=5), dict(pilot_samples=4096)):
        call(**kw)
    assert fake.calls == n0 + 9
    assert len(m._CACHE) == 10
    call()                                                                # 10 entries: the first is still there
    assert fake.calls == n0 + 9
    call(seed=999)                                                        # 11th distinct key evicts the LRU one ...
    assert len(m._CACHE) == 10 and fake.calls == n0 + 10
    call(CI=90)                                                           # ... which is CI=90 (the base was just touched)
    assert fake.calls == n0 + 11
    call()
    assert fake.calls == n0 + 11


def test_result_cache_is_skipped_when_the_job_is_not_reproducible(m, monkeypatch):
    from importlib import import_module

    dist_ = import_module(PKG + ".distributed")
    fake = _FakeJob(m)
    monkeypatch.setattr(dist_, "run_job", fake)
    m.cache_clear()
    m.seed(None)
    for _ in range(2):                        # entropy seed (the reference's unseeded CLI): never cached
        m.run_model(*USA_ARGS, 95, 100, 50000, USA_DATES)
    assert fake.calls == 2 and len(m._CACHE) == 0
    for _ in range(2):                        # a time budget makes the result depend on the clock
        res, err, _ = m.run_model(*USA_ARGS, 95, 100, 50000, USA_DATES, max_running_time=30, seed=1)
        assert "Maximum computation time of 30s exceeded. Only 50 of the desired 100" in err
    assert fake.calls == 4 and len(m._CACHE) == 0
    m.seed(12345)                             # app.py:442: np.random.seed(12345) on every callback
    a = m.run_model(*USA_ARGS, 95, 100, 50000, USA_DATES)
    m.seed(12345)
    b = m.run_model(*USA_ARGS, 95, 100, 50000, USA_DATES)
    assert b is a and fake.calls == 5
    m.seed(None)
    m.cache_clear()
