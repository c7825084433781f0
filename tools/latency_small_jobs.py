#!/usr/bin/env python
"""Wall-clock latency of run_model() for the sizes the reference's callers use (CLI default 100
samples, Dash app up to 10 000, app.py:144-152) and beyond; median / min / max of 20 calls, distinct seeds."""
import importlib
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
vax = importlib.import_module("covid19-vaccination-model_b200")
DATES = dict(start_date="2020-12-15", end_date="2022-07-1")
for n in (100, 1000, 10_000, 65_536, 100_000, 1_000_000):
    for N in (50_000, 1_000_000):
        ts = []
        for i in range(23):
            t0 = time.perf_counter()
            res, err, msg = vax.run_model([35, 45], [12, 25], [0.0, 0.02], [5, 7], [1, 1.2], [5, 10], 95, n, N,
                                          DATES, seed=100 + i)
            ts.append(time.perf_counter() - t0)
        t = sorted(ts[3:])
        print(json.dumps({"n_rep": n, "N": N, "ms_median": 1e3 * t[10], "ms_min": 1e3 * t[0], "ms_max": 1e3 * t[-1]}))
