"""Probe: one parameter tuple for all samples (pure sampler noise) at several sizes: passes, pilot, fold."""
import importlib, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
vax = importlib.import_module("covid19-vaccination-model_b200"); nat = vax._native
ctx = vax.default_context(0)
tup = (0.4, 0.2, 0.01, 6.0, 0.011, 0.07)
b = np.repeat(np.array(tup), 2)
for n in (int(1e7), int(1e8), int(3e8), int(1e9)):
    cfg = nat.make_config(b, 50000, 564, n, 77, 0.025, 0.975)
    t0 = time.perf_counter(); out = ctx.run_bounds(cfg); w = time.perf_counter() - t0
    print(n, "wall %.3f s device %.1f ms passes %d pilot %.1f ms fold %.1f ms n_pilot %d launches %d" % (
        w, out.c.device_ms, out.c.n_passes, out.c.pilot_ms, out.c.fold_ms, out.c.n_pilot, out.c.fold_launches), flush=True)
