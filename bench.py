#!/usr/bin/env python
"""bench.py -- MC sample-weeks/s of the Monte Carlo hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N ...            # CPU arm (oracle port, host cores)

A "step" is one whole job of the hot path: parameter draw, daily stepper over every sample,
fold, quantile bands and means on the host (what run_model returns).  Workload at N=1 is
BASELINE.json configs[1]: README USA scenario, 1e6 samples, N=50000, 564 days; under
torchrun every rank folds 1e6 samples of a global range of N*1e6 (weak scaling) and the
integer accumulators are merged with one NCCL all-reduce per pass.

value   whole-job throughput from CUDA events on the library's stream (inputs -- 12 bounds --
        resident; nothing else is read), max over ranks.
e2e     the same job through the public Python API run_model() with host arguments and host
        result arrays, wall clock around the call, per-step.
One sample-week = 7 day-steps of model.py:82-127 for one sample.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

PKG = "covid19-vaccination-model_b200"
PKG_DIR = PKG
USA = dict(pro=[35, 45], anti=[12, 25], pressure=[0.0, 0.02], tau=[5, 7], nv0=[1, 1.2], nvmax=[5, 10])
DATES = dict(start_date="2020-12-15", end_date="2022-07-1")
T_DAYS = 564
N_POP = 50000
CI_PCT = 95
# BASELINE.json configs[3]: all bounds x2 width about the same centres, pressure 0-0.1, 5 years
WIDE = dict(pro=[30, 50], anti=[5.5, 31.5], pressure=[0.0, 0.1], tau=[4, 8], nv0=[0.9, 1.3], nvmax=[2.5, 12.5])
WIDE_DATES = dict(start_date="2020-12-15", end_date="2025-12-14")
WIDE_T = 1826
BYTES_PER_SAMPLE_WEEK = 112.0   # 4 int32 counts per sample-day if materialised
HBM_FALLBACK_GBS = 6650.0       # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent
N_SM, SCHEDULERS_PER_SM = 148, 4
# ncu counters of the dominant kernel (instructions and DRAM bytes per launch) live in a committed
# file keyed by the sha256 of the kernel source they were captured from: a kernel edit without a
# new capture makes the roofline fraction null ("stale_counters") instead of silently wrong.
COUNTERS_JSON = os.path.join(ROOT, "profiles", "kernel_counters.json")
KERNEL_SOURCE = os.path.join(ROOT, PKG_DIR, "csrc", "vaxmc_kernels.cuh")
# north_star check (3) on the driver's path: one FIXED global job whose digest must not depend on N
BITEXACT_JOB = dict(samples=3_000_000, seed=12345, ci=95)


def bounds_frac(b):
    import numpy as np

    return np.concatenate([np.array(b["pro"]) / 100, np.array(b["anti"]) / 100,
                           np.array(b["pressure"], dtype=float), np.array(b["tau"], dtype=float),
                           np.array(b["nv0"]) / 100, np.array(b["nvmax"]) / 100])


def usa_bounds_frac():
    return bounds_frac(USA)


def grid_boxes(n_boxes, seed=20261019):
    """configs[4]: parameter-bound boxes around the USA centres (generator recorded here):
    every bound pair = centre * (1 + 0.25*u1) -/+ halfwidth * (0.5 + u2), u ~ U(-1,1) / U(0,1),
    numpy default_rng(seed); pressure keeps lo = 0."""
    import numpy as np

    rng = np.random.default_rng(seed)
    keys = ["pro", "anti", "pressure", "tau", "nv0", "nvmax"]
    out = []
    for _ in range(n_boxes):
        box = {}
        for k in keys:
            lo, hi = USA[k]
            c, h = (lo + hi) / 2, (hi - lo) / 2
            c2 = c * (1 + 0.25 * rng.uniform(-1, 1))
            h2 = h * (0.5 + rng.uniform(0, 1))
            box[k] = [max(c2 - h2, 0.0), c2 + h2] if k != "pressure" else [0.0, c2 + h2]
        out.append(box)
    return out


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.stop_flag = threading.Event()
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
                if self.stop_flag.is_set():
                    break
        except Exception:
            pass

    def finish(self) -> dict:
        self.stop_flag.set()
        if self.proc:
            try:
                self.proc.terminate()
            except Exception:
                pass
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n_, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n_)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


def load_counters():
    """profiles/kernel_counters.json if it was captured from the kernel source that is built now."""
    import hashlib

    try:
        with open(KERNEL_SOURCE, "rb") as f:
            sha = hashlib.sha256(f.read()).hexdigest()
    except OSError:
        return None, {"stale_counters": True, "why": "kernel source not found"}
    try:
        with open(COUNTERS_JSON) as f:
            c = json.load(f)
    except Exception:
        return None, {"stale_counters": True, "why": "profiles/kernel_counters.json missing", "kernels_sha256": sha}
    if c.get("kernels_sha256") != sha:
        return None, {"stale_counters": True, "kernels_sha256": sha, "counters_sha256": c.get("kernels_sha256"),
                      "why": "vaxmc_kernels.cuh changed since the ncu capture behind profiles/kernel_counters.json"}
    return c, {"stale_counters": False, "kernels_sha256": sha, "source": c.get("source")}


def workload_config(args, world):
    """The `config` block: identical for the native arm and the reference arm of one run."""
    wide = args.workload == "wide5y"
    return {"workload": ("wide-uncertainty 5-year sweep at %g samples per GPU (BASELINE configs[3]); " if wide else
                         "README USA scenario at %g samples per GPU (BASELINE configs[1]); ") % args.samples
                        + "global range = n_gpus x samples_per_gpu, sharded by sample id",
            "samples_per_gpu": args.samples, "samples_global": args.samples * world,
            "days": WIDE_T if wide else T_DAYS, "N": N_POP, "CI": args.ci if args.ci is not None else CI_PCT,
            "seed": "2000+step", "l2": "flushed between steps (256 MiB memset)"}


def cpu_port_rate(n_samples: int, threads: int, seed: int = 1, bounds=None, T=None):
    """sample-weeks/s of the oracle port (C restatement of model.py:29-136 with numpy's own
    MT19937 + Poisson algorithms) on `threads` host threads, incl. the per-date reduction."""
    import numpy as np

    import ref_oracle as ro

    T = T or T_DAYS
    mt = ro.MT(seed)
    b = usa_bounds_frac() if bounds is None else bounds
    params, _ = ro.sample_param_combinations(mt, b[0:2], b[2:4], b[4:6], b[6:8], b[8:10], b[10:12],
                                             n_samples)
    ro.bulk_counts(seed, params[:256], T, N_POP, threads=threads)      # warm caches / build
    t0 = time.perf_counter()
    daily, weekly = ro.bulk_counts(seed, params, T, N_POP, threads=threads)
    for arr in (daily[0], daily[1], daily[2], weekly):                      # model.py:220-223
        np.quantile(arr, [0.025, 0.975], axis=1)
        arr.mean(axis=1)
    dt = time.perf_counter() - t0
    return n_samples * T / 7.0 / dt, dt


def run_reference_arm(args):
    """The reference's CPU algorithm for this path on the box's host cores (the C port: the
    reference itself is pure Python and cannot travel to the GPU box).  Same `config` as the
    native arm; each step is a bounded sample of that workload (--cpu-samples), and the line
    states how the rate behaves between that sample and the full per-GPU size (`linearity`) and
    what ONE core does (`single_thread`), as BASELINE.md section 3 asks."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return 0
    wide = args.workload == "wide5y"
    bounds, T = (bounds_frac(WIDE), WIDE_T) if wide else (usa_bounds_frac(), T_DAYS)
    threads = os.cpu_count() or 1
    n = args.cpu_samples
    rates, times = [], []
    for i in range(args.warmup + args.steps):
        r, dt = cpu_port_rate(n, threads, seed=100 + i, bounds=bounds, T=T)
        if i >= args.warmup:
            rates.append(r); times.append(dt)
    value = n * T / 7.0 * len(times) / sum(times)
    # the same port once at the native arm's full per-GPU size, and on one thread
    full_rate, full_dt = cpu_port_rate(args.samples, threads, seed=7, bounds=bounds, T=T) \
        if args.samples * T <= 1.2e9 else (None, None)
    one_rate, one_dt = cpu_port_rate(max(n // 10, 1000), 1, seed=8, bounds=bounds, T=T)
    sample = (f"{n} samples x {T} days per step, C port of model.py with numpy-identical MT19937/Poisson, "
              f"{threads} pthreads, np.quantile+mean per date")
    line = {
        "impl": "reference", "metric": "MC sample-weeks/sec", "value": value, "unit": "sample-weeks/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int64/f64", "data": "synthetic",
        "config": workload_config(args, world),
        "cpu_baseline": {"value": value, "unit": "sample-weeks/s", "cores": threads, "kind": "port",
                         "sample": sample},
        "linearity": {"samples": [n, args.samples], "rate": [value, full_rate],
                      "ratio_full_over_sample": (full_rate / value) if full_rate else None,
                      "full_size_s": full_dt,
                      "note": "CPU rate at the bounded per-step sample vs ONE pass at the native arm's full "
                              "per-GPU size: the rate is flat in n, so the bounded sample stands for the config"},
        "single_thread": {"value": one_rate, "unit": "sample-weeks/s", "cores": 1, "samples": max(n // 10, 1000),
                          "seconds": one_dt,
                          "note": "per-core figure (BASELINE.md 3); the reference's own pure-Python loop "
                                  "measured 3.5-4.0e4 sample-weeks/s per core in the survey container"},
        "e2e": {"value": value, "unit": "sample-weeks/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": ("the reference is pure Python and cannot travel to the GPU box (no /root/reference "
                 "there); this arm times the C restatement of its algorithm, which is ~25x faster "
                 "per core than the Python original (SURVEY.md 6: 3.5-4.0e4 sample-weeks/s/core)"),
    }
    print(json.dumps(line))
    return 0


def time_grid(vax, ctx, rank, world, local_rank, dist, n_boxes, samples, ci, wave, warmup, steps):
    """configs[4]: independent scenarios -> partition BOXES over ranks, no data-path collective.
    Host wall clock around the whole grid incl. the copy-back of every box's bands, max over ranks."""
    import torch

    boxes = grid_boxes(n_boxes)
    mine = list(range(rank, n_boxes, world))
    box_args = [(boxes[b]["pro"], boxes[b]["anti"], boxes[b]["pressure"], boxes[b]["tau"], boxes[b]["nv0"],
                 boxes[b]["nvmax"]) for b in mine]

    def one_pass(seed0):
        # the product's batched entry point: ONE library call per rank (vx_run_grid) pipelines the
        # boxes -- next boxes' pilots in the shadow of the current boxes' folds
        res = vax.run_model_grid(box_args, ci, samples, N_POP, DATES, seeds=[seed0 + b for b in mine],
                                 device=local_rank, wave=wave)
        checksum = 0.0
        for r, err, msg in res:
            assert err == ""
            checksum += float(r["people_vaccinated_per_hundred"]["mean"][-1])
        return checksum

    for i in range(warmup):
        one_pass(10_000 * i)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    l0 = ctx.launch_count
    per_step = []
    for i in range(steps):
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        one_pass(1_000_000 + 10_000 * i)
        per_step.append(time.perf_counter() - t0)
    t = torch.tensor(per_step, dtype=torch.float64, device=f"cuda:{local_rank}")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    per_step = [float(x) for x in t.tolist()]
    units = n_boxes * samples * T_DAYS / 7.0
    return {"workload": "batched multi-scenario grid (BASELINE configs[4]): %d bound boxes x %g samples, CI=%g, "
                        "boxes partitioned over ranks, one vx_run_grid call per rank" % (n_boxes, samples, ci),
            "boxes": n_boxes, "samples_per_box": samples, "days": T_DAYS, "N": N_POP,
            "generator": "bench.grid_boxes(seed=20261019)",
            "sample_weeks_per_s": units * steps / sum(per_step), "ms_per_step": 1e3 * sum(per_step) / steps,
            "step_ms": [1e3 * x for x in per_step], "median_step_ms": 1e3 * statistics.median(per_step),
            "step_spread": (max(per_step) - min(per_step)) / (sum(per_step) / steps) if steps > 1 else 0.0,
            "steps": steps, "warmup": warmup, "grid_wave": wave, "gpu_launches": ctx.launch_count - l0,
            "timing": "host wall clock around the whole grid incl. D2H of every box's bands (max over ranks)"}


def run_grid(args, vax, ctx, rank, world, local_rank, dist):
    ci = args.ci if args.ci is not None else 99.0
    g = time_grid(vax, ctx, rank, world, local_rank, dist, args.boxes, args.samples, ci, args.grid_wave,
                  args.warmup, args.steps)
    if rank == 0:
        print(json.dumps({
            "metric": "MC sample-weeks/sec", "value": g["sample_weeks_per_s"], "unit": "sample-weeks/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": g["ms_per_step"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "int32 state / fp32 sampler / fp64 arrivals", "data": "synthetic",
            "config": {k: g[k] for k in ("workload", "boxes", "samples_per_box", "days", "N", "generator")},
            "grid": g, "gpu_launches": g["gpu_launches"]}))
    if dist is not None:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--samples", type=int, default=1_000_000, help="samples per GPU per step")
    ap.add_argument("--cpu-samples", type=int, default=200_000, help="samples of the CPU baseline")
    ap.add_argument("--workload", default="usa", choices=["usa", "wide5y", "grid"],
                    help="usa: configs[1]/[2] (use --samples for 1e9/N); wide5y: configs[3]; grid: configs[4]")
    ap.add_argument("--boxes", type=int, default=1024, help="grid workload: number of bound boxes")
    ap.add_argument("--grid-wave", type=int, default=0, help="grid workload: boxes in flight per pipeline stage (0: library default)")
    ap.add_argument("--ci", type=float, default=None, help="CI in percent (default 95; grid: 99)")
    ap.add_argument("--direct", action="store_true",
                    help="force the materialise + radix-select path (HBM-bound passes) instead of the streaming fold")
    ap.add_argument("--target-samples", type=int, default=1_000_000_000,
                    help="global samples of the target_1e9 block (BASELINE configs[2]); 0 skips it")
    ap.add_argument("--grid-boxes", type=int, default=1024,
                    help="boxes of the grid_1024 block (BASELINE configs[4], 1e6 samples each, CI=99); 0 skips it")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    if args.impl == "reference":
        return run_reference_arm(args)

    import numpy as np
    import torch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the native path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    vax = importlib.import_module(PKG)
    nat = vax._native
    dmod = vax.distributed
    ctx = vax.default_context(local_rank)
    for kv in filter(None, os.environ.get("VX_OPTIONS", "").split(",")):   # e.g. VX_OPTIONS=fold_pad_smem=20000
        k, v = kv.split("=")
        ctx.set_option(k, float(v))

    if args.workload == "grid":
        return run_grid(args, vax, ctx, rank, world, local_rank, dist)
    global T_DAYS, CI_PCT
    wl_bounds, wl_dates = (USA, DATES) if args.workload == "usa" else (WIDE, WIDE_DATES)
    if args.workload == "wide5y":
        T_DAYS = WIDE_T
    if args.ci is not None:
        CI_PCT = args.ci
    n_global = args.samples * world
    bounds = bounds_frac(wl_bounds)
    q_lo, q_hi = (1 - CI_PCT / 100) / 2.0, (1 + CI_PCT / 100) / 2.0
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local_rank}")  # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    merge = os.environ.get("VAXMC_MERGE", "nccl")      # "torch": slab all-reduced from Python (cross-check)
    all_reduce = dmod.make_torch_reducer(dist) if dist else None
    if dist is not None and merge != "torch":
        dmod.ensure_comm(ctx, dist)                     # rendezvous only: 128-byte NCCL id over torch.distributed

    def one_job(seed):
        cfg = nat.make_config(bounds, N_POP, T_DAYS, n_global, seed, q_lo, q_hi,
                              direct_max=(n_global + 1) if args.direct else 0)
        if world == 1:
            return ctx.run_bounds(cfg)
        if merge == "torch":
            return dmod.run_sharded(dmod.ContextEngine(ctx), cfg, rank, world, all_reduce,
                                    lambda: nat.ResultBuffers(T_DAYS))
        return ctx.run_bounds_comm(cfg)      # libvaxmc's own ncclAllReduce, in-stream behind the fold

    # ---------------- device-timed leg ("value")
    for i in range(args.warmup):
        one_job(1000 + i)
    sampler = ClockSampler(local_rank)      # rank 0 only: nvidia-smi takes driver locks
    if rank == 0:
        sampler.start()
        time.sleep(0.15)
    barrier()
    launches0 = ctx.launch_count
    dev_ms, fold_ms, pilot_ms, merge_ms, fold_launches, passes, fold_samples = [], [], [], [], 0, [], 0
    t_wall0 = time.perf_counter()
    for i in range(args.steps):
        flush.zero_()                       # L2 flush between timed iterations (outside the events)
        torch.cuda.synchronize()
        out = one_job(2000 + i)
        dev_ms.append(out.c.device_ms); fold_ms.append(out.c.fold_ms); pilot_ms.append(out.c.pilot_ms)
        merge_ms.append(out.c.merge_ms); fold_samples += out.c.fold_samples
        fold_launches += out.c.fold_launches; passes.append(out.c.n_passes)
    n_pilot, side_samples = int(out.c.n_pilot), int(out.c.side_samples)
    barrier()
    wall_s = time.perf_counter() - t_wall0
    launches = ctx.launch_count - launches0
    clocks = sampler.finish() if rank == 0 else None

    tot_ms = torch.tensor([sum(dev_ms)], dtype=torch.float64, device=f"cuda:{local_rank}")
    if dist is not None:
        dist.all_reduce(tot_ms, op=dist.ReduceOp.MAX)
    total_ms = float(tot_ms.item())
    sample_weeks_per_step = n_global * T_DAYS / 7.0
    value = sample_weeks_per_step * args.steps / (total_ms * 1e-3)

    # ---------------- end-to-end leg through run_model (host args in, host arrays out)
    e2e = None
    if not args.no_e2e:
        e2e_s = []
        for i in range(min(args.warmup, 2) + args.steps):
            flush.zero_(); torch.cuda.synchronize()
            if dist is not None:
                dist.barrier()
            t0 = time.perf_counter()
            res, err, msg = vax.run_model(wl_bounds["pro"], wl_bounds["anti"], wl_bounds["pressure"],
                                          wl_bounds["tau"], wl_bounds["nv0"], wl_bounds["nvmax"], CI_PCT,
                                          n_global, N_POP, wl_dates, seed=3000 + i,
                                          direct_max=(n_global + 1) if args.direct else 0)
            checksum = float(res["people_vaccinated_per_hundred"]["mean"][-1])  # host read of the result
            dt = time.perf_counter() - t0
            assert err == "" and checksum > 0
            if i >= min(args.warmup, 2):
                e2e_s.append(dt)
        tt = torch.tensor([sum(e2e_s)], dtype=torch.float64, device=f"cuda:{local_rank}")
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        W = nat.weeks(T_DAYS)
        e2e = {"value": sample_weeks_per_step * len(e2e_s) / float(tt.item()), "unit": "sample-weeks/s",
               "h2d_bytes_per_step": 12 * 8 + 64,                     # the bounds + scalars of vx_config
               "d2h_bytes_per_step": 3 * (3 * T_DAYS + W) * 8 + 64,   # mean/lower/upper of 4 series
               "ms_per_step": 1e3 * float(tt.item()) / len(e2e_s),
               "ms_median_rank0": 1e3 * statistics.median(e2e_s), "ms_max_rank0": 1e3 * max(e2e_s),
               "api": "run_model(...) of covid19-vaccination-model_b200 (ctypes -> C ABI)"}

    # ---------------- north_star check (3): ONE fixed global job, whatever N is.  Its digest must be
    # the same on every line of a 1/2/4/8-GPU scaling run (integer slab merge => bit-identical).
    import hashlib

    bj = BITEXACT_JOB
    res, err, msg = vax.run_model(USA["pro"], USA["anti"], USA["pressure"], USA["tau"], USA["nv0"], USA["nvmax"],
                                  bj["ci"], bj["samples"], N_POP, DATES, seed=bj["seed"])
    h = hashlib.sha256()
    for k in vax.SERIES:
        for f in ("mean", "lower", "upper"):
            h.update(np.ascontiguousarray(res[k][f], dtype=np.float64).tobytes())
    digest = h.hexdigest()
    dg = torch.tensor([int(digest[:15], 16)], dtype=torch.int64, device=f"cuda:{local_rank}")
    lo, hi = dg.clone(), dg.clone()
    if dist is not None:
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    bitexact = {"job": "USA scenario, n=%d, seed=%d, CI=%g, sharded over n_gpus ranks" % (bj["samples"], bj["seed"], bj["ci"]),
                "n": bj["samples"], "sha256": digest, "arrays": "mean,lower,upper of the 4 series (float64 bytes)",
                "identical_on_all_ranks": bool(lo.item() == hi.item()), "agnostics": msg, "error": err}

    # ---------------- BASELINE configs[2] / second half of `metric`: wall time to CI bands at 1e9
    # samples (global, over the N ranks), through run_model
    target = None
    if args.target_samples > 0 and args.workload == "usa":
        tw = []
        for i in range(1 + 3):
            if dist is not None:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r9, e9, _ = vax.run_model(USA["pro"], USA["anti"], USA["pressure"], USA["tau"], USA["nv0"], USA["nvmax"],
                                      95, args.target_samples, N_POP, DATES, seed=9000 + i)
            chk = float(r9["people_vaccinated_per_hundred"]["upper"][-1])
            dt = time.perf_counter() - t0
            assert e9 == "" and chk > 0 and r9["number_finished_samples"] == args.target_samples
            if i >= 1:
                tw.append(dt)
        tt9 = torch.tensor([sum(tw)], dtype=torch.float64, device=f"cuda:{local_rank}")
        if dist is not None:
            dist.all_reduce(tt9, op=dist.ReduceOp.MAX)
        wall = float(tt9.item()) / len(tw)
        target = {"workload": "README USA scenario, %g samples GLOBAL over %d GPU(s), CI=95 (BASELINE configs[2])"
                              % (args.target_samples, world),
                  "samples_global": args.target_samples, "wall_s_to_ci_bands": wall,
                  "sample_weeks_per_s": args.target_samples * 564 / 7.0 / wall, "timed_calls": len(tw),
                  "warmup_calls": 1, "api": "run_model(..., n_rep=%d, CI=95), host wall clock, max over ranks"
                                            % args.target_samples}

    # ---------------- BASELINE configs[4] as a secondary block: 1024 boxes x 1e6 samples, CI=99
    grid = None
    if args.grid_boxes > 0 and args.workload == "usa":
        grid = time_grid(vax, ctx, rank, world, local_rank, dist, args.grid_boxes, 1_000_000, 99.0, args.grid_wave, 2, 3)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    # ---------------- roofline of the dominant kernel (the streaming-fold stepper)
    # The stepper reads and writes (almost) nothing: it is bound by instruction issue, so the
    # fraction is warp instructions issued per second over what 148 SMs x 4 schedulers can issue
    # at the SM clock sampled during this run.  Instructions per launch come from the committed ncu
    # capture of THIS kernel source (profiles/kernel_counters.json), the duration from this run's
    # CUDA events.
    peak_hbm, peak_src = measured_hbm_peak()
    fold_ms_avg = sum(fold_ms) / max(fold_launches, 1)       # per launch, rank 0, CUDA events
    counters, cstate = load_counters()
    sm_hz = 1e6 * (clocks.get("sm_mhz") or 1965.0)
    issue_peak = N_SM * SCHEDULERS_PER_SM * sm_hz
    sample_days = (fold_samples / max(fold_launches, 1)) * T_DAYS     # what one fold launch really stepped
    sd_per_s = sample_days / (fold_ms_avg * 1e-3) if fold_launches else None
    inst_per_launch = traffic = issue_ach = None
    if counters is not None and fold_launches and args.workload == "usa":
        scale = sample_days / float(counters["sample_days_per_launch"])
        inst_per_launch = counters["warp_inst_per_launch"] * scale
        traffic = counters["dram_bytes_per_launch"] * scale
        issue_ach = inst_per_launch / (fold_ms_avg * 1e-3)
    hbm_equiv = BYTES_PER_SAMPLE_WEEK * (sample_days / 7.0) / (fold_ms_avg * 1e-3) / 1e9 if fold_launches else None
    roofline = {
        "kernel": "vx_sim_kernel<SIM_FOLD>", "bound": "issue",
        "achieved": issue_ach, "peak": issue_peak, "unit": "warp-inst/s",
        "frac": (issue_ach / issue_peak) if issue_ach else None,
        "traffic": traffic,
        "peak_source": "148 SMs x 4 warp schedulers x SM clock sampled during the timed region (%.0f MHz)" % (sm_hz / 1e6),
        "ms_per_launch": fold_ms_avg, "share_of_step": sum(fold_ms) / sum(dev_ms) if sum(dev_ms) else None,
        "warp_inst_per_launch": inst_per_launch,
        "warp_inst_per_warp_day": (inst_per_launch / (sample_days / 32.0)) if inst_per_launch else None,
        "sample_days_per_s_per_gpu": sd_per_s,
        "counters": cstate,
        "hbm_equivalent": {"bound": "hbm", "achieved": hbm_equiv, "peak": peak_hbm, "unit": "GB/s",
                           "frac": (hbm_equiv / peak_hbm) if hbm_equiv else None, "peak_source": peak_src,
                           "note": "SURVEY 8(d)'s 112 B per sample-week (the int32 trajectories IF materialised) over "
                                   "the kernel time; the kernel folds them on the fly, so these bytes never move "
                                   "(see traffic): a derived figure, not the bound"},
        "note": "issue-bound stepper (Philox + sampler arithmetic in registers, DRAM traffic ~0 by design): "
                "achieved = ncu smsp__inst_executed.sum per launch (profiles/kernel_counters.json, same kernel "
                "source by sha256) / CUDA-event launch time of this run",
    }
    if cstate.get("stale_counters"):
        roofline["stale_counters"] = True

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        rate, dt = cpu_port_rate(args.cpu_samples, threads, bounds=bounds, T=T_DAYS)
        cpu = {"value": rate, "unit": "sample-weeks/s", "cores": threads, "kind": "port",
               "sample": f"{args.cpu_samples} samples x {T_DAYS} days of the same workload "
                         f"({dt:.1f} s), C port of model.py, {threads} pthreads"}

    line = {
        "metric": "MC sample-weeks/sec", "value": value, "unit": "sample-weeks/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int32 state / fp32 sampler / fp64 arrivals", "data": "synthetic",
        "config": workload_config(args, world),
        "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
        "roofline": roofline, "cpu_baseline": cpu,
        "bitexact": bitexact, "target_1e9": target, "grid_1024": grid,
        "wall_s_timed_region": wall_s,
        "passes_per_step": statistics.mean(passes) if passes else 0,
        "breakdown_ms_per_step": {"pilot": sum(pilot_ms) / args.steps, "fold": sum(fold_ms) / args.steps,
                                  "merge_allreduce": sum(merge_ms) / args.steps,
                                  "total_device": sum(dev_ms) / args.steps,
                                  "non_fold": (sum(dev_ms) - sum(fold_ms)) / args.steps},
        "job_split": {"pilot_samples": n_pilot, "side_samples_stepped_beside_the_pilot": side_samples,
                      "fold_samples_per_launch": fold_samples / max(fold_launches, 1),
                      "merge": ("single GPU: none" if world == 1 else
                                "torch.distributed all_reduce from Python" if merge == "torch" else
                                "ncclAllReduce(uint64 sum) issued by libvaxmc on its own stream")},
    }
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
