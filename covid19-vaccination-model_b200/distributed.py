"""One process per GPU: shard the Monte Carlo sample range over ranks and merge with NCCL.

The path shards by independent units: every sample is independent
(model.py:176-179) and its Philox stream is keyed by its GLOBAL id, so rank r simply folds
the contiguous range shard_range(n, r, world).  The only exchange step is, per streaming
pass, ONE all-reduce(sum) of the int64 accumulator slab (per-date sums, counts below the
windows, window histograms; the shard's rejection count and the fixed-point moments of
p_soft_no, model.py:280-287, 339-342, ride in its header) -- integer addition, hence results
that are bit-identical for 1/2/4/8 GPUs.

torch.distributed is plumbing only: it ships the 128-byte NCCL unique id once per process
(ensure_comm); the collective itself is libvaxmc's own ncclAllReduce on its own stream
(vx_run_bounds_comm).  run_sharded + the staged vx_job_* calls remain as the backend-agnostic
driver (gloo tests, VAXMC_MERGE=torch); all arithmetic is in libvaxmc.so.
"""
from __future__ import annotations

import os

import numpy as np

from . import _native as nat


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous [first, first+count) of global sample ids owned by `rank`."""
    base, rem = divmod(int(n), int(world))
    first = rank * base + min(rank, rem)
    count = base + (1 if rank < rem else 0)
    return first, count


class _DevicePtr:
    """Exposes a raw device pointer through __cuda_array_interface__ (tensor handoff)."""

    def __init__(self, ptr: int, n_words: int):
        self.__cuda_array_interface__ = {
            "shape": (n_words,), "typestr": "<i8", "data": (ptr, False), "version": 3,
            "strides": None,
        }


class ContextEngine:
    """Adapter: a libvaxmc Context as the engine run_sharded drives."""

    def __init__(self, ctx: nat.Context):
        self.ctx = ctx

    def begin(self, cfg, first, count):
        self.ctx.job_begin(cfg, first, count)

    def param_stats(self) -> np.ndarray:
        return self.ctx.job_param_stats()

    def set_param_stats(self, st) -> bool:
        return self.ctx.job_set_param_stats(st)

    def pilot(self):
        self.ctx.job_pilot()

    def is_direct(self) -> bool:
        return self.ctx.job_is_direct()

    def accumulate(self):
        self.ctx.job_accumulate()

    def acc_tensor(self):
        import torch

        ptr, n = self.ctx.job_acc()
        if n == 0:
            return None
        return torch.as_tensor(_DevicePtr(ptr, n), device=f"cuda:{self.ctx.device}")

    def finalize(self, out) -> int:
        return self.ctx.job_finalize(out)

    def end(self):
        self.ctx.job_end()

    def stats_tensor(self, st: np.ndarray):
        import torch

        return torch.from_numpy(st).to(f"cuda:{self.ctx.device}")


def bounds_infeasible(cfg) -> bool:
    """No acceptable (p_pro, p_anti) pair exists in the box: the reference burns 10*n_rep + 1
    rejections and returns (None, None) (model.py:280-287).  A property of cfg alone, so every
    rank takes the same decision without talking to the others."""
    b = cfg.bounds
    return min(b[0], b[1]) + min(b[2], b[3]) > 1.0


def run_sharded(engine, cfg, rank: int, world: int, all_reduce, make_out):
    """Drive one rank of a sharded job.  `all_reduce(tensor)` sums in place across ranks and
    returns when the result is usable.

    One collective per pass: the shard's parameter statistics (rejections, moments of
    p_soft_no) ride in the header of the int64 slab, and the 10*n_rep abort rule is decided on
    the MERGED count (a rank must never leave before a collective the others still enter).  A
    direct job (n <= direct_max) is computed whole on every rank and needs no collective at all.
    """
    direct = cfg.n_samples <= (cfg.direct_max if cfg.direct_max > 0 else nat.DIRECT_MAX_DEFAULT) \
        or cfg.mode == nat.MODE_EXPECTED
    first, count = (0, cfg.n_samples) if direct else shard_range(cfg.n_samples, rank, world)
    trace = _Trace(rank)
    engine.begin(cfg, first, count)
    trace("begin")
    try:
        st = engine.param_stats()
        trace("param_stats")
        out = make_out()
        if getattr(cfg, "bounds", None) is not None and bounds_infeasible(cfg):
            out.c.aborted = 1                     # certain on every rank: nobody enters a collective
            out.c.n_rejected = int(st[0])
            return out
        engine.pilot()
        trace("pilot")
        for _ in range(64):
            engine.accumulate()
            trace("accumulate")
            acc = engine.acc_tensor()
            if acc is not None:
                all_reduce(acc)
            trace("acc all-reduce")
            rc = engine.finalize(out)
            trace("finalize")
            if rc != nat.VX_REFINE:
                return out
        raise RuntimeError("window refinement did not converge")
    finally:
        engine.end()


class _Trace:
    """VX_TRACE=1: wall time of each stage of run_sharded on rank 0 (stderr)."""

    def __init__(self, rank):
        self.on = rank == 0 and bool(os.environ.get("VX_TRACE"))
        if self.on:
            import time

            self.clock = time.perf_counter
            self.t = self.clock()

    def __call__(self, what):
        if self.on:
            import sys

            now = self.clock()
            print(f"[vx trace py] {what:24s} {1e3 * (now - self.t):8.3f} ms", file=sys.stderr)
            self.t = now


def make_torch_reducer(dist):
    """all_reduce(tensor) over torch.distributed.

    The library synchronises its own stream before returning from every staged call, so the
    collective only has to be complete before the next library call."""
    import torch

    def all_reduce(t):
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        if t.is_cuda:
            torch.cuda.current_stream(t.device).synchronize()

    return all_reduce


def broadcast_seed(key: int) -> int:
    """Rank 0's Philox key on every rank.  An entropy seed (seed=None, the reference's default
    call) is drawn independently per process; the replicated pilot and the window plan must
    come from ONE key or the ranks would all-reduce slabs of different shapes."""
    try:
        import torch.distributed as dist
    except ImportError:
        return key
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return key
    box = [int(key)]
    dist.broadcast_object_list(box, src=0)
    return int(box[0])


def run_job(cfg, device=None):
    """Whole job for run_model: sharded when launched under torchrun, single GPU otherwise."""
    world = 1
    dist = None
    try:
        import torch.distributed as dist_mod

        if dist_mod.is_available() and dist_mod.is_initialized():
            dist = dist_mod
            world = dist.get_world_size()
    except ImportError:
        pass
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0")) if world > 1 else 0
    ctx = nat.default_context(int(device))
    if world == 1:
        return ctx.run_bounds(cfg)
    if os.environ.get("VAXMC_MERGE", "nccl") == "torch":
        # the slab merged by torch.distributed from Python (two host synchronisations around it):
        # kept for backends without NCCL and as the cross-check of the in-library path
        return run_sharded(ContextEngine(ctx), cfg, dist.get_rank(), world, make_torch_reducer(dist),
                           lambda: nat.ResultBuffers(cfg.T))
    ensure_comm(ctx, dist)
    return ctx.run_bounds_comm(cfg)


def ensure_comm(ctx, dist):
    """Give the context its NCCL communicator (once): rank 0 makes the unique id through the
    library, torch.distributed ships the 128 bytes -- rendezvous is all torch does on this path;
    the all-reduce itself is issued by libvaxmc on its own stream, right behind the fold."""
    world = dist.get_world_size()
    if ctx.comm_world == world:
        return
    box = [nat.comm_unique_id() if dist.get_rank() == 0 else None]
    dist.broadcast_object_list(box, src=0)
    ctx.comm_init(box[0], dist.get_rank(), world)
