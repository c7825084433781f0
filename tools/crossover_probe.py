"""Direct (materialise + select) vs streamed (pilot + fold) wall time of run_model-sized jobs."""
import importlib, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
vax = importlib.import_module("covid19-vaccination-model_b200"); nat = vax._native
ctx = vax.default_context(0)
b = bench.usa_bounds_frac()
for n in (49152, 65536, 98304, 131072, 163840, 196608, 262144):
    row = [n]
    for dm in (n + 1, 1):
        ts = []
        for i in range(15):
            cfg = nat.make_config(b, 50000, 564, n, 900 + i, 0.025, 0.975, direct_max=dm)
            t0 = time.perf_counter(); out = ctx.run_bounds(cfg); ts.append((time.perf_counter() - t0) * 1e3)
        row.append(round(float(np.median(ts[3:])), 3))
    print("n %d direct %.3f ms streamed %.3f ms" % tuple(row), flush=True)
