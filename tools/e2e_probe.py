import importlib, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.getcwd())
import bench
vax = importlib.import_module("covid19-vaccination-model_b200"); nat = vax._native
ctx = vax.default_context(0)
b = bench.usa_bounds_frac()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")
def run(tag, do_flush, api):
    ws, ds = [], []
    for i in range(13):
        if do_flush: flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if api == "c":
            cfg = nat.make_config(b, 50000, 564, 1_000_000, 500 + i, 0.025, 0.975)
            out = ctx.run_bounds(cfg); dev = out.c.device_ms
        else:
            res, err, msg = vax.run_model(bench.USA["pro"], bench.USA["anti"], bench.USA["pressure"], bench.USA["tau"], bench.USA["nv0"], bench.USA["nvmax"], 95, 1_000_000, 50000, bench.DATES, seed=700 + i)
            dev = float("nan")
        w = (time.perf_counter() - t0) * 1e3
        if i >= 3: ws.append(w); ds.append(dev)
    print(tag, "wall median %.3f min %.3f | device median %.3f" % (np.median(ws), min(ws), np.median(ds)))
run("C ABI   no flush", False, "c")
run("C ABI   flush   ", True, "c")
run("run_model no flush", False, "py")
run("run_model flush   ", True, "py")
