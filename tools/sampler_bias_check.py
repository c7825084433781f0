#!/usr/bin/env python
"""GPU-vs-GPU bias check of the default Poisson sampler against the exact (PTRS) validation mode.

Single-value bounds (every sample has the same tuple, so the spread across samples is sampler
noise only), n samples per arm through the streaming fold; compares the per-date means and the
2.5 % / 97.5 % order statistics.  The means of n = 1e8 samples resolve a relative bias of the
sampler of ~1e-4 of the per-date standard deviation.
"""
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
    only = sys.argv[2] if len(sys.argv) > 2 else ""          # substring of the scenario name
    vax = importlib.import_module("covid19-vaccination-model_b200")
    nat = vax._native
    ctx = vax.default_context(0)
    out = []
    for name, tup in (("survey tuple", (0.4, 0.2, 0.01, 6.0, 0.011, 0.07)),
                      ("tiny pressure", (0.36, 0.24, 0.0005, 5.2, 0.0101, 0.052)),
                      ("strong pressure", (0.30, 0.055, 0.1, 4.0, 0.009, 0.125)),
                      # the README USA box itself (mixed parameters): here sd is the spread across samples
                      ("USA bounds", (0.35, 0.45, 0.12, 0.25, 0.0, 0.02, 5, 7, 0.01, 0.012, 0.05, 0.10))):
        if only and only not in name:
            continue
        b = np.array(tup) if len(tup) == 12 else np.repeat(np.array(tup), 2)
        res = {}
        for exact in (0, 1):
            ctx.set_option("exact_poisson", exact)
            cfg = nat.make_config(b, 50000, 564, n, 1234 + exact, 0.025, 0.975)
            res[exact] = ctx.run_bounds(cfg)
        ctx.set_option("exact_poisson", 0)
        worst, dq, z2, cells = 0.0, 0.0, 0.0, 0
        for s in (0, 1, 3):                                    # vaccinated, weekly doses, stock
            m0, m1 = res[0].mean[s], res[1].mean[s]
            sd = np.maximum((res[1].upper[s] - res[1].lower[s]) / 3.92, 1e-12)   # sampler-noise sd per date
            z = np.abs(m0 - m1) / (sd * np.sqrt(2.0 / n))
            z = z[(res[1].upper[s] - res[1].lower[s]) > 0]
            worst = max(worst, float(z.max()) if len(z) else 0.0)
            z2 += float((z * z).sum()); cells += len(z)
            unit = {0: 100.0 / 50000, 1: 1e6 / 50000, 3: 100.0 / 50000}[s]     # one person / one dose
            dq = max(dq, float(np.abs(res[0].lower[s] - res[1].lower[s]).max() / unit),
                     float(np.abs(res[0].upper[s] - res[1].upper[s]).max() / unit))
        out.append({"tuple": name, "n_per_arm": n, "max_abs_z_of_mean_difference": worst,
                    "cells": cells, "mean_z2": z2 / max(cells, 1),     # ~1 when the two samplers have the same means
                    "max_band_difference_in_counts": dq,
                    "ms_default": res[0].c.device_ms, "ms_exact": res[1].c.device_ms})
        print(json.dumps(out[-1]))
    return 0


if __name__ == "__main__":
    sys.exit(main())
