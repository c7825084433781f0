"""oracle/make_golden.py -- generates tests/golden/*.npz from the reference itself.

Run in the build container (needs /root/reference):  python oracle/make_golden.py
Every fixture is an output of the unmodified reference model.py (imported through
oracle/reference_harness.py); the oracle (vaxmodel_oracle.c) and the CUDA path are
both tested against these files.  Fixtures are small (tens of KB) and committed.
"""
from __future__ import annotations

import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from reference_harness import expected_value_mode, load_reference  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
SERIES = ("people_vaccinated_per_hundred", "daily_vaccinations_per_million",
          "cum_number_vac_received_per_hundred", "vaccines_in_stock_per_hundred")

USA = dict(pro=[35, 45], anti=[12, 25], pressure=[0.0, 0.02], tau=[5, 7], nv0=[1, 1.2],
           nvmax=[5, 10], dates=("2020-12-15", "2022-07-1"))
WIDE5Y = dict(pro=[30, 50], anti=[5.5, 31.5], pressure=[0.0, 0.1], tau=[4, 8], nv0=[0.9, 1.3],
              nvmax=[2.5, 12.5], dates=("2020-12-15", "2025-12-14"))


def pack_results(res):
    d = {}
    for k in SERIES:
        for f in ("mean", "lower", "upper"):
            d[f"{k}/{f}"] = np.asarray(res[k][f], dtype=np.float64)
        d[f"{k}/dates"] = np.array([str(x) for x in res[k]["dates"]])
    d["number_finished_samples"] = np.int64(res["number_finished_samples"])
    return d


def run_model_case(model, cfg, CI, n_rep, N, seed, expected=False):
    np.random.seed(seed)
    args = (cfg["pro"], cfg["anti"], cfg["pressure"], cfg["tau"], cfg["nv0"], cfg["nvmax"], CI,
            n_rep, N, dict(start_date=cfg["dates"][0], end_date=cfg["dates"][1]))
    model.run_sampling.cache_clear()
    if expected:
        with expected_value_mode():
            res, err, msg = model.run_model(*args)
    else:
        res, err, msg = model.run_model(*args)
    # the parameter tuples the run consumed (same seed -> same first draws)
    np.random.seed(seed)
    params, soft = model.sample_param_combinations(
        np.array(cfg["pro"]) / 100, np.array(cfg["anti"]) / 100, np.array(cfg["pressure"]),
        np.array(cfg["tau"]), np.array(cfg["nv0"]) / 100, np.array(cfg["nvmax"]) / 100, n_rep)
    d = pack_results(res)
    d.update(msg_error=np.array(err), msg_agnostics_pct=np.array(msg),
             params=np.array(params, dtype=np.float64), p_soft_no=np.array(soft, dtype=np.float64),
             CI=np.float64(CI), n_rep=np.int64(n_rep), N=np.int64(N), seed=np.int64(seed),
             bounds=np.array([cfg["pro"], cfg["anti"], cfg["pressure"], cfg["tau"], cfg["nv0"],
                              cfg["nvmax"]], dtype=np.float64),
             start_date=np.array(cfg["dates"][0]), end_date=np.array(cfg["dates"][1]))
    return d


def main():
    os.makedirs(OUT, exist_ok=True)
    model = load_reference()

    # 1. seeded stochastic run_model, USA README scenario (SURVEY 8c known answers)
    np.savez_compressed(os.path.join(OUT, "usa_seed12345_ci95_n100.npz"),
                        **run_model_case(model, USA, 95, 100, 50000, 12345))
    # 2. same with the CLI default --CI 0.95 (D5 quirk: quantiles 0.49525/0.50475)
    np.savez_compressed(os.path.join(OUT, "usa_seed2024_ci0p95_n64.npz"),
                        **run_model_case(model, USA, 0.95, 64, 50000, 2024))
    # 3. wide 5-year scenario, CI=99, N=50000
    np.savez_compressed(os.path.join(OUT, "wide5y_seed7_ci99_n24.npz"),
                        **run_model_case(model, WIDE5Y, 99, 24, 50000, 7))
    # 4. expected-value mode, run_model level (north_star check 1)
    np.savez_compressed(os.path.join(OUT, "usa_expected_seed12345_ci95_n100.npz"),
                        **run_model_case(model, USA, 95, 100, 50000, 12345, expected=True))
    np.savez_compressed(os.path.join(OUT, "wide5y_expected_seed7_ci90_n32_N1e6.npz"),
                        **run_model_case(model, WIDE5Y, 90, 32, 1000000, 7, expected=True))

    # 5. single realizations: stochastic (seeded) and expected-value
    singles = {}
    plist = np.array([
        [0.4, 0.2, 0.01, 6.0, 0.011, 0.07],
        [0.35, 0.25, 0.0, 5.0, 0.01, 0.05],      # pressure 0 -> lam2 == 0 always
        [0.45, 0.12, 0.02, 7.0, 0.012, 0.10],
        [0.30, 0.055, 0.1, 4.0, 0.009, 0.125],   # wide sweep corner, strong pressure
        [0.999, 0.001, 0.05, 3.0, 0.002, 0.10],  # no agnostics at all
        [0.05, 0.9, 0.02, 6.0, 0.3, 0.3],        # tiny pro group, huge stock: clamps dominate
    ])
    singles["params"] = plist
    for i, p in enumerate(plist):
        for T, N in ((564, 50000), (100, 1000), (1826, 1000000)):
            with expected_value_mode():
                r = model.run_single_realization(*p, T, N)
            singles[f"exp/{i}/{T}/{N}"] = np.array([r[k] for k in SERIES], dtype=np.float64)
            np.random.seed(1000 + i)
            r = model.run_single_realization(*p, T, N)
            singles[f"sto/{i}/{T}/{N}"] = np.array([r[k] for k in SERIES], dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "single_realizations.npz"), **singles)

    # 6. parameter sampling incl. rejection and abort (model.py:280-287)
    ps = {}
    cases = {
        "usa": ([0.35, 0.45], [0.12, 0.25], [0.0, 0.02], [5, 7], [0.01, 0.012], [0.05, 0.10], 50, 1),
        "reject": ([0.5, 0.9], [0.3, 0.6], [0.0, 0.1], [3, 4], [0.002, 0.0024], [0.1, 0.1], 40, 2),
        "abort": ([0.7, 0.9], [0.5, 0.6], [0.0, 0.1], [3, 4], [0.002, 0.0024], [0.1, 0.1], 5, 3),
        "degenerate": ([0.4, 0.4], [0.2, 0.2], [0.01, 0.01], [6, 6], [0.011, 0.011], [0.07, 0.07], 3, 4),
    }
    for name, (a, b, c, d, e, f, n, seed) in cases.items():
        np.random.seed(seed)
        params, soft = model.sample_param_combinations(
            np.array(a), np.array(b), np.array(c), np.array(d), np.array(e), np.array(f), n)
        ps[f"{name}/bounds"] = np.array([a, b, c, d, e, f], dtype=np.float64)
        ps[f"{name}/n_rep"] = np.int64(n)
        ps[f"{name}/seed"] = np.int64(seed)
        ps[f"{name}/aborted"] = np.bool_(params is None)
        if params is not None:
            ps[f"{name}/params"] = np.array(params, dtype=np.float64)
            ps[f"{name}/p_soft_no"] = np.array(soft, dtype=np.float64)
        ps[f"{name}/next_double"] = np.float64(np.random.random_sample())  # stream position check
    np.savez_compressed(os.path.join(OUT, "param_sampling.npz"), **ps)

    # 7. raw numpy legacy streams: doubles and poisson variates over the lambda range the path sees
    np.random.seed(99)
    lams = np.array([0.0, 1e-3, 0.5, 3.0, 9.999, 10.0, 10.5, 63.0, 300.0, 2857.0, 1e4, 1e5])
    pv = np.array([[np.random.poisson(l) for _ in range(40)] for l in lams], dtype=np.int64)
    np.random.seed(5)
    dbl = np.random.random_sample(16)
    np.savez_compressed(os.path.join(OUT, "numpy_streams.npz"), lams=lams, poisson=pv, doubles=dbl,
                        numpy_version=np.array(np.__version__))
    print("golden fixtures written to", os.path.abspath(OUT))
    for f in sorted(os.listdir(OUT)):
        print(f"  {f}  {os.path.getsize(os.path.join(OUT, f))} B")


def country_data_golden():
    """N4: the reference's own get_country_data (plot.py:8-36) on the committed fixture CSV.
    plot.py imports plotly (absent here) -> stub modules; pd.read_csv is pointed at the fixture
    instead of the OWID URL.  Output: tests/golden/country_data_pivot.csv."""
    import sys
    import types

    import pandas as pd

    for name in ("plotly", "plotly.express", "plotly.graph_objects", "plotly.subplots"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["plotly.express"].colors = types.SimpleNamespace(qualitative=types.SimpleNamespace(Safe=["#000"] * 12))
    sys.modules["plotly.subplots"].make_subplots = lambda *a, **k: None
    sys.modules.pop("plot", None)                      # the harness' stub `plot` must not shadow the real one
    sys.path.insert(0, "/root/reference")
    import plot as ref_plot

    fixture = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden",
                           "owid_vaccinations_sample.csv")
    real = pd.read_csv
    pd.read_csv = lambda *a, **k: real(fixture)
    try:
        countries, pivot = ref_plot.get_country_data()
    finally:
        pd.read_csv = real
    pivot.to_csv(os.path.join(os.path.dirname(fixture), "country_data_pivot.csv"))
    with open(os.path.join(os.path.dirname(fixture), "country_data_countries.txt"), "w") as f:
        f.write("\n".join(countries) + "\n")
    print("country data golden:", pivot.shape, list(countries))


if __name__ == "__main__":
    if "--country-data" in sys.argv:       # only the N4 fixture (does not touch the model goldens)
        country_data_golden()
    else:
        main()
        country_data_golden()
