#!/usr/bin/env python
"""north_star check (3) on real hardware: for a fixed seed the sharded (NCCL) result on N GPUs is
bit-identical to the single-GPU result.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tests/multigpu_bitexact.py

Every rank runs the sharded job through run_model (one NCCL all-reduce of the integer slab);
rank 0 additionally runs the same job alone on its GPU through the single-GPU entry point and
compares every array with np.array_equal.  Prints one JSON line per scenario.
"""
import importlib
import json
import os
import sys

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
    ]
    ok_all = True
    for name, b, ci, n, dates in scen:
        res, err, msg = vax.run_model(*b, ci, n, 50000, dates, seed=20261020)     # sharded over `world` ranks
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
            cfg = nat.make_config(bounds, 50000, T, n, 20261020, (1 - ci / 100) / 2, (1 + ci / 100) / 2)
            single = vax.default_context(local).run_bounds(cfg)                    # one GPU, no torch.distributed
            equal = all(np.array_equal(res[k][f], getattr(single, f)[s])
                        for s, k in enumerate(vax.SERIES) for f in ("mean", "lower", "upper"))
            ok = equal and same_across_ranks and err == "" and res["number_finished_samples"] == n
            ok_all = ok_all and ok
            print(json.dumps({"scenario": name, "n_gpus": world, "bit_identical_to_single_gpu": equal,
                              "identical_on_all_ranks": same_across_ranks, "passes": int(single.c.n_passes),
                              "agnostics": msg, "ok": ok}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0 if ok_all else 1


if __name__ == "__main__":
    sys.exit(main())
