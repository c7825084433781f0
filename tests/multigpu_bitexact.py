#!/usr/bin/env python
"""north_star check (3) on real hardware: for a fixed seed the sharded (NCCL) result on N GPUs is
bit-identical to the single-GPU result.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tests/multigpu_bitexact.py

Every rank runs the sharded job through run_model (one ncclAllReduce of the integer slab, issued
by libvaxmc on its own stream); rank 0 additionally runs the same job alone on its GPU through
the single-GPU entry point and compares every array with np.array_equal.  Also: the same job
with the slab merged by torch.distributed from Python (VAXMC_MERGE=torch) gives the same bytes,
an entropy seed (seed=None) is rank 0's on every rank, and an infeasible parameter box aborts on
every rank without a hang.  Prints one JSON line per scenario; exit code 0 iff all are ok.
"""
import importlib
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    vax = importlib.import_module("covid19-vaccination-model_b200")
    nat = vax._native
    scen = [
        ("usa n=3e6 CI=95", ([35, 45], [12, 25], [0.0, 0.02], [5, 7], [1, 1.2], [5, 10]), 95, 3_000_000,
         dict(start_date="2020-12-15", end_date="2022-07-1")),
        ("usa n=200003 CI=0.95 (CLI default quirk)", ([35, 45], [12, 25], [0.0, 0.02], [5, 7], [1, 1.2], [5, 10]),
         0.95, 200_003, dict(start_date="2020-12-15", end_date="2022-07-1")),
        ("wide5y n=400000 CI=99", ([30, 50], [5.5, 31.5], [0.0, 0.1], [4, 8], [0.9, 1.3], [2.5, 12.5]), 99,
         400_000, dict(start_date="2020-12-15", end_date="2025-12-14")),
        ("usa n=1000 (direct job, no collective)", ([35, 45], [12, 25], [0.0, 0.02], [5, 7], [1, 1.2], [5, 10]),
         95, 1000, dict(start_date="2020-12-15", end_date="2022-07-1")),
        # the Dash app's maximum population (app.py:157-166): deliveries/stock spread over millions of doses,
        # the widest windows and the largest slab the all-reduce ever carries
        ("usa N=1e6 population n=1e6 CI=95", ([35, 45], [12, 25], [0.0, 0.02], [5, 7], [1, 1.2], [5, 10]),
         95, 1_000_000, dict(start_date="2020-12-15", end_date="2022-07-1"), 1_000_000),
    ]
    ok_all = True
    for name, b, ci, n, dates, *rest in scen:
        N = rest[0] if rest else 50000
        t0 = time.perf_counter()
        res, err, msg = vax.run_model(*b, ci, n, N, dates, seed=20261020)         # sharded over `world` ranks
        wall_ms = 1e3 * (time.perf_counter() - t0)
        digest = float(sum(res[k][f].sum() for k in vax.SERIES for f in ("mean", "lower", "upper")))
        t = torch.tensor([digest], dtype=torch.float64, device=f"cuda:{local}")
        lo, hi = t.clone(), t.clone()
        if world > 1:
            dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        same_across_ranks = bool(lo.item() == hi.item())
        if rank == 0:
            T = len(res[vax.SERIES[0]]["mean"])
            bounds = vax.model._bounds12(np.array(b[0]) / 100, np.array(b[1]) / 100, np.array(b[2]),
                                         np.array(b[3]), np.array(b[4]) / 100, np.array(b[5]) / 100)
            cfg = nat.make_config(bounds, N, T, n, 20261020, (1 - ci / 100) / 2, (1 + ci / 100) / 2)
            single = vax.default_context(local).run_bounds(cfg)                    # one GPU, no torch.distributed
            equal = all(np.array_equal(res[k][f], getattr(single, f)[s])
                        for s, k in enumerate(vax.SERIES) for f in ("mean", "lower", "upper"))
            ok = equal and same_across_ranks and err == "" and res["number_finished_samples"] == n
            ok_all = ok_all and ok
            print(json.dumps({"scenario": name, "n_gpus": world, "bit_identical_to_single_gpu": equal,
                              "identical_on_all_ranks": same_across_ranks, "passes": int(single.c.n_passes),
                              "sharded_wall_ms": round(wall_ms, 3), "single_gpu_device_ms": round(float(single.c.device_ms), 3),
                              "agnostics": msg, "ok": ok}))
    usa = ([35, 45], [12, 25], [0.0, 0.02], [5, 7], [1, 1.2], [5, 10])
    dates = dict(start_date="2020-12-15", end_date="2022-07-1")

    def digest_of(res):
        return float(sum(res[k][f].sum() for k in vax.SERIES for f in ("mean", "lower", "upper")))

    def same_on_all_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local}")
        lo, hi = t.clone(), t.clone()
        if world > 1:
            dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        return bool(lo.item() == hi.item())

    # (a) the Python-driven merge (torch.distributed all_reduce on the slab) gives the same bytes
    a_res, _, _ = vax.run_model(*usa, 95, 500_000, 50000, dates, seed=77)
    os.environ["VAXMC_MERGE"] = "torch"
    vax.cache_clear()
    b_res, _, _ = vax.run_model(*usa, 95, 500_000, 50000, dates, seed=77)
    del os.environ["VAXMC_MERGE"]
    eq = all(np.array_equal(a_res[k][f], b_res[k][f]) for k in vax.SERIES for f in ("mean", "lower", "upper"))
    ok = eq and same_on_all_ranks(digest_of(a_res))
    ok_all = ok_all and ok
    if rank == 0:
        print(json.dumps({"scenario": "in-library ncclAllReduce == torch.distributed all_reduce of the slab",
                          "n_gpus": world, "ok": ok}))
    # (b) seed=None (the reference's default call): one key for the whole job
    vax.seed(None)
    c_res, c_err, _ = vax.run_model(*usa, 95, 300_000, 50000, dates)
    ok = c_err == "" and c_res["number_finished_samples"] == 300_000 and same_on_all_ranks(digest_of(c_res))
    ok_all = ok_all and ok
    if rank == 0:
        print(json.dumps({"scenario": "entropy seed (seed=None): identical bands on every rank", "n_gpus": world, "ok": ok}))
    # (c) a parameter box with no acceptable (pro, anti) pair, and one that is almost infeasible
    for name, pro, anti in (("infeasible box", [70, 90], [50, 60]), ("almost infeasible box (abort by the merged count)", [60, 99], [38, 99])):
        d_res, d_err, d_msg = vax.run_model(pro, anti, [0, 0.1], [3, 4], [0.2, 0.24], [10, 10], 95, 200_000, 50000, dates, seed=3)
        ok = d_res is None and d_err.startswith("ERROR: The pertentages") and same_on_all_ranks(1.0 if d_res is None else 0.0)
        ok_all = ok_all and ok
        if rank == 0:
            print(json.dumps({"scenario": name + ": every rank aborts, nobody hangs", "n_gpus": world, "ok": ok}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0 if ok_all else 1


if __name__ == "__main__":
    sys.exit(main())
