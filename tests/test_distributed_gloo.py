"""CPU, world_size=2 over gloo: the multi-rank driver (covid19-vaccination-model_b200/distributed.py).

The arithmetic of a rank lives in libvaxmc.so and needs a GPU; here a FakeEngine with the
same interface folds a deterministic integer function of the GLOBAL sample id, which is
enough to pin what the driver is responsible for: disjoint contiguous shards that cover the
range, ONE int64 all-reduce of the accumulator slab per pass (the shard's parameter statistics
ride in its header), the VX_REFINE loop, the abort path decided on the merged rejection count
-- and that the merged integers do not depend on the number of ranks.
"""
import importlib
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
PKG = "covid19-vaccination-model_b200"


class _C:
    aborted = 0
    n_rejected = 0
    n_finished = 0


class _Out:
    def __init__(self):
        self.c = _C()
        self.merged = None


class FakeEngine:
    """Integer fold of f(id) = (id * 2654435761) % 1009 over the shard, two passes; the shard's
    rejection count rides in word 2 of the slab, as in libvaxmc."""

    def __init__(self, infeasible=False, skewed_rank=None):
        self.infeasible = infeasible
        self.skewed_rank = skewed_rank      # this rank ALONE exceeds 10*n rejections, the others see none
        self.passes = 0
        self.calls = []

    def begin(self, cfg, first, count):
        self.first, self.count, self.n = first, count, cfg.n_samples
        self.calls.append(("begin", first, count))

    def param_stats(self):
        ids = np.arange(self.first, self.first + self.count, dtype=np.int64)
        self.rej = int((ids % 3 == 0).sum()) if not self.infeasible else 11 * self.count
        if self.skewed_rank is not None:
            self.rej = 10 * self.n + 5 if dist.get_rank() == self.skewed_rank else 0
        s = (ids % 100) / 100.0
        return np.array([float(self.rej), s.sum(), (s * s).sum(), float(self.count)])

    def pilot(self):
        self.calls.append(("pilot",))

    def accumulate(self):
        ids = np.arange(self.first, self.first + self.count, dtype=np.uint64)
        f = (ids * np.uint64(2654435761)) % np.uint64(1009)
        hist = np.bincount(f.astype(np.int64), minlength=1009)
        self.acc = torch.from_numpy(np.concatenate([[self.count, int(f.sum()), self.rej], hist]).astype(np.int64))
        self.passes += 1

    def acc_tensor(self):
        return self.acc

    def finalize(self, out):
        out.merged = self.acc.clone().numpy()
        out.c.n_finished = int(out.merged[0])
        if out.merged[2] > 10 * self.n:              # model.py:282-287 on the MERGED count
            out.c.aborted = 1
            return 0
        return 1 if self.passes < 2 else 0          # VX_REFINE once, then VX_OK

    def end(self):
        self.calls.append(("end",))


class _Cfg:
    def __init__(self, n):
        self.n_samples = n
        self.T = 10
        self.direct_max = 1          # always streamed
        self.mode = 0


def _worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        d = importlib.import_module(PKG + ".distributed")
        eng = FakeEngine()
        n_reduces = [0]

        def all_reduce(t):
            n_reduces[0] += 1
            dist.all_reduce(t, op=dist.ReduceOp.SUM)

        out = d.run_sharded(eng, _Cfg(n), rank, world, all_reduce, _Out)
        bad = FakeEngine(infeasible=True)
        out_bad = d.run_sharded(bad, _Cfg(n), rank, world, all_reduce, _Out)
        # one shard alone breaks the 10*n rule, the other sees no rejection at all: both ranks must
        # still meet in the collective and BOTH report the abort (no early return, no hang)
        skew = FakeEngine(skewed_rank=0)
        out_skew = d.run_sharded(skew, _Cfg(n), rank, world, all_reduce, _Out)
        # entropy seeds differ per process; a collective entry point must end up with rank 0's
        key = d.broadcast_seed(1000 + rank)
        m = importlib.import_module(PKG + ".model")
        m.seed(None)
        k2, det = m._next_seed(None, collective=True)
        k3, _ = m._next_seed(None)                      # non-collective: stays local
        keys = [torch.tensor([k2 >> 32, k2 & 0xFFFFFFFF, k3 >> 32, k3 & 0xFFFFFFFF], dtype=torch.int64)
                for _ in range(world)]
        dist.all_gather(keys, keys[rank].clone())
        # result cache under torchrun: every rank must take the same hit/miss decision,
        # or one would sit in the job's all-reduce alone.  The stand-in job does a real collective.
        jobs = []

        def fake_job(cfg, device=None):
            t = torch.tensor([int(cfg.seed & 0x7FFFFFFF)], dtype=torch.int64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            jobs.append((int(t.item()), int(cfg.seed & 0x7FFFFFFF)))
            ob = m.nat.ResultBuffers(cfg.T)
            ob.c.n_finished = cfg.n_samples
            return ob

        d.run_job = fake_job
        m.cache_clear()
        a = ([35, 45], [12, 25], [0.0, 0.02], [5, 7], [1, 1.2], [5, 10], 95, 100, 50000,
             dict(start_date="2021-01-01", end_date="2021-01-10"))
        m.run_model(*a, seed=5); m.run_model(*a, seed=5)          # miss, hit
        m.run_model(*a); m.run_model(*a)                          # entropy seed: two jobs, one key each
        m.run_model(*a, seed=5)                                   # still a hit
        q.put((rank, out.merged.tolist(), eng.passes, n_reduces[0], eng.calls, out_bad.c.aborted, bad.passes,
               out_skew.c.aborted, skew.passes, key, det, [k.tolist() for k in keys], jobs))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _expected(n):
    ids = np.arange(n, dtype=np.uint64)
    f = (ids * np.uint64(2654435761)) % np.uint64(1009)
    hist = np.bincount(f.astype(np.int64), minlength=1009)
    rej = int((np.arange(n) % 3 == 0).sum())
    return np.concatenate([[n, int(f.sum()), rej], hist]).astype(np.int64)


@pytest.mark.parametrize("n", [1001, 64])
def test_two_rank_gloo_merge_equals_single_rank(n):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    merged = _expected(n)
    for rank, m, passes, n_red, calls, aborted, bad_passes, skew_aborted, skew_passes, key, det, keys, jobs in res:
        assert len(jobs) == 3 and all(mx == own for mx, own in jobs)     # same hit/miss and same key on every rank
        assert np.array_equal(np.array(m), merged)              # integer merge is exact
        assert passes == 2                                       # one VX_REFINE round trip
        assert n_red == 2 + 1 + 1                                # ONE collective per pass (+1 +1: the two aborting jobs)
        assert skew_aborted == 1 and skew_passes == 1            # decided on the MERGED count, on every rank
        assert key == 1000 and not det                           # rank 0's key everywhere
        assert keys[0][:2] == keys[1][:2] and keys[0][2:] != keys[1][2:]
        assert calls[0][0] == "begin" and calls[1] == ("pilot",) and calls[-1] == ("end",)
        # neither shard alone breaks the 10*n rule (11*n/2 each); the merged count does
        assert aborted == 1 and bad_passes == 1


def test_shard_range_partitions_exactly():
    d = importlib.import_module(PKG + ".distributed")
    for n in (0, 1, 7, 8, 9, 1000, 10**9 + 7):
        for world in (1, 2, 3, 4, 8):
            parts = [d.shard_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0
            for (f0, c0), (f1, c1) in zip(parts, parts[1:]):
                assert f0 + c0 == f1
            assert parts[-1][0] + parts[-1][1] == n
            counts = [c for _, c in parts]
            assert max(counts) - min(counts) <= 1


def test_single_process_driver_equals_sharded_sum():
    """Same fake engine, world=1: the merged slab must equal the 2-rank result (no collective)."""
    d = importlib.import_module(PKG + ".distributed")
    eng = FakeEngine()
    out = d.run_sharded(eng, _Cfg(1001), 0, 1, lambda t: None, _Out)
    assert np.array_equal(out.merged, _expected(1001))
