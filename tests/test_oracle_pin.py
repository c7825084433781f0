"""Pins the CPU oracle (oracle/vaxmodel_oracle.c) to the reference's own outputs.

The reference ships no tests; the fixtures under tests/golden/ were produced by
oracle/make_golden.py from the unmodified /root/reference/model.py.  Everything here
is bit-exact (np.array_equal on float64) unless stated.
"""
import os

import numpy as np
import pytest

import ref_oracle as ro

SERIES = ro.SERIES


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def test_numpy_streams(golden_dir):
    g = _load(golden_dir, "numpy_streams.npz")
    mt = ro.MT(5)
    got = np.array([mt.random_sample() for _ in range(len(g["doubles"]))])
    assert np.array_equal(got, g["doubles"])
    mt = ro.MT(99)
    pv = np.array([[mt.poisson(float(l)) for _ in range(g["poisson"].shape[1])] for l in g["lams"]])
    assert np.array_equal(pv, g["poisson"])


@pytest.mark.parametrize("case", ["usa", "reject", "abort", "degenerate"])
def test_param_sampling(golden_dir, case):
    g = _load(golden_dir, "param_sampling.npz")
    b = g[f"{case}/bounds"]
    mt = ro.MT(int(g[f"{case}/seed"]))
    params, soft = ro.sample_param_combinations(mt, *b, int(g[f"{case}/n_rep"]))
    if bool(g[f"{case}/aborted"]):
        assert params is None and soft is None
    else:
        assert np.array_equal(params, g[f"{case}/params"])
        assert np.array_equal(soft, g[f"{case}/p_soft_no"])
    # the oracle consumed exactly as many stream words as the reference
    assert mt.random_sample() == float(g[f"{case}/next_double"])


def test_single_realizations(golden_dir):
    g = _load(golden_dir, "single_realizations.npz")
    plist = g["params"]
    n_checked = 0
    for key in g.files:
        if key == "params":
            continue
        kind, i, T, N = key.split("/")
        i, T, N = int(i), int(T), int(N)
        if kind == "exp":
            r = ro.run_single_realization(None, 1, plist[i], T, N)
        else:
            r = ro.run_single_realization(ro.MT(1000 + i), 0, plist[i], T, N)
        got = np.array([r[k] for k in SERIES])
        assert np.array_equal(got, g[key]), key
        n_checked += 1
    assert n_checked == 2 * len(plist) * 3


@pytest.mark.parametrize("name,expected", [
    ("usa_seed12345_ci95_n100.npz", False),
    ("usa_seed2024_ci0p95_n64.npz", False),
    ("wide5y_seed7_ci99_n24.npz", False),
    ("usa_expected_seed12345_ci95_n100.npz", True),
    ("wide5y_expected_seed7_ci90_n32_N1e6.npz", True),
])
def test_run_model(golden_dir, name, expected):
    g = _load(golden_dir, name)
    b = g["bounds"]
    mt = ro.MT(int(g["seed"]))
    res, err, msg = ro.run_model(
        mt, b[0], b[1], b[2], b[3], b[4], b[5], float(g["CI"]), int(g["n_rep"]), int(g["N"]),
        dict(start_date=str(g["start_date"]), end_date=str(g["end_date"])), mode=1 if expected else 0)
    assert err == str(g["msg_error"])
    assert msg == str(g["msg_agnostics_pct"])
    assert res["number_finished_samples"] == int(g["number_finished_samples"])
    for k in SERIES:
        for f in ("mean", "lower", "upper"):
            assert np.array_equal(res[k][f], g[f"{k}/{f}"]), (k, f)
        assert [str(x) for x in res[k]["dates"]] == list(g[f"{k}/dates"])


def test_survey_known_answers(golden_dir):
    """expected-value known answers and the seeded USA run."""
    r = ro.run_single_realization(None, 1, [0.4, 0.2, 0.01, 6.0, 0.011, 0.07], 564, 50000)
    idx = [0, 6, 7, 100, 300, 563]
    assert np.allclose(r["people_vaccinated_per_hundred"][idx],
                       [0.12571428571428572, 0.6263358758006715, 0.8184963936159436,
                        30.04348693865135, 68.11442484595923, 78.35613916778212], rtol=1e-14)
    assert np.allclose(r["cum_number_vac_received_per_hundred"][idx],
                       [1.1, 1.1, 2.334, 41.814, 237.02, 503.02], rtol=1e-14)
    g = _load(golden_dir, "usa_seed12345_ci95_n100.npz")
    assert g["people_vaccinated_per_hundred/mean"][-1] == 74.18210000000003
    assert str(g["msg_agnostics_pct"]) == "Agnosticts: 37 - 46%"


def test_bulk_counts_consistency():
    """Independent-stream fan-out obeys the model's conservation laws and is thread-count independent."""
    mt = ro.MT(3)
    params, _ = ro.sample_param_combinations(mt, [0.35, 0.45], [0.12, 0.25], [0, 0.02], [5, 7],
                                             [0.01, 0.012], [0.05, 0.1], 64)
    d1, w1 = ro.bulk_counts(11, params, 200, 50000, threads=1)
    d2, w2 = ro.bulk_counts(11, params, 200, 50000, threads=5)
    assert np.array_equal(d1, d2) and np.array_equal(w1, w2)
    n_vacc, received, stock = d1
    assert np.array_equal(n_vacc + stock, received)
    assert (np.diff(n_vacc, axis=0) >= 0).all()
    # sample 0 of the fan-out is MT(seed+0) run through the scalar path
    r, cnt = ro.run_single_realization(ro.MT(11), 0, params[0], 200, 50000, want_counts=True)
    assert np.array_equal(cnt[0], n_vacc[:, 0])
    assert np.array_equal(cnt[1][0::7][:w1.shape[0]] +
                          np.concatenate([np.zeros(5, np.int32), cnt[1][5::7]])[:w1.shape[0]] *
                          (np.arange(w1.shape[0]) * 7 >= 30), w1[:, 0])
