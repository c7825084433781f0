"""ctypes binding of libvaxmc.so (include/vaxmc.h).  Thin: no arithmetic happens here.

The library is the product; this module fails loudly when it is missing or when no CUDA
device is usable -- there is no CPU fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

SERIES = (
    "people_vaccinated_per_hundred",
    "daily_vaccinations_per_million",
    "cum_number_vac_received_per_hundred",
    "vaccines_in_stock_per_hundred",
)

VX_OK, VX_REFINE = 0, 1
VX_ERR_NODEVICE = -5
MODE_STOCHASTIC, MODE_EXPECTED = 0, 1
DIRECT_MAX_DEFAULT = 196608  # vx_config.direct_max == 0 (include/vaxmc.h: VX_DIRECT_MAX_DEFAULT)


class VaxmcError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libvaxmc error {code}: {msg}")
        self.code = code


class vx_config(C.Structure):
    _fields_ = [
        ("bounds", C.c_double * 12),
        ("N", C.c_int64),
        ("T", C.c_int32),
        ("mode", C.c_int32),
        ("n_samples", C.c_int64),
        ("seed", C.c_uint64),
        ("q_lo", C.c_double),
        ("q_hi", C.c_double),
        ("direct_max", C.c_int64),
        ("pilot_samples", C.c_int64),
        ("max_running_time", C.c_double),
    ]


_dp = C.POINTER(C.c_double)


class vx_result(C.Structure):
    _fields_ = [
        ("mean", _dp * 4),
        ("lower", _dp * 4),
        ("upper", _dp * 4),
        ("n_finished", C.c_int64),
        ("n_rejected", C.c_int64),
        ("soft_no_mean", C.c_double),
        ("soft_no_std", C.c_double),
        ("aborted", C.c_int32),
        ("n_passes", C.c_int32),
        ("device_ms", C.c_double),
        ("fold_ms", C.c_double),
        ("pilot_ms", C.c_double),
        ("fold_launches", C.c_int64),
        ("merge_ms", C.c_double),
        ("n_pilot", C.c_int64),
        ("fold_samples", C.c_int64),
        ("side_samples", C.c_int64),
    ]


class vx_nccl_id(C.Structure):
    _fields_ = [("internal", C.c_char * 128)]


ABI_VERSION = 2   # include/vaxmc.h: VX_ABI_VERSION


# every symbol include/vaxmc.h declares: (restype, argtypes)
_vp = C.c_void_p
_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)
_u32p = C.POINTER(C.c_uint32)
_cfgp = C.POINTER(vx_config)
_resp = C.POINTER(vx_result)
SYMBOLS = {
    "vx_create": (C.c_int, [C.POINTER(_vp), C.c_int]),
    "vx_destroy": (None, [_vp]),
    "vx_last_error": (C.c_char_p, [_vp]),
    "vx_version": (C.c_char_p, []),
    "vx_abi_version": (C.c_int, []),
    "vx_sizeof": (C.c_int64, [C.c_int]),
    "vx_comm_unique_id": (C.c_int, [C.POINTER(vx_nccl_id)]),
    "vx_comm_init": (C.c_int, [_vp, C.POINTER(vx_nccl_id), C.c_int, C.c_int]),
    "vx_comm_destroy": (C.c_int, [_vp]),
    "vx_comm_info": (C.c_int, [_vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "vx_run_bounds_comm": (C.c_int, [_vp, _cfgp, _resp]),
    "vx_run_grid": (C.c_int, [_vp, _cfgp, C.c_int32, _resp, C.c_int32]),
    "vx_host_shard_range": (None, [C.c_int64, C.c_int, C.c_int, _i64p, _i64p]),
    "vx_device_count": (C.c_int, []),
    "vx_set_option": (C.c_int, [_vp, C.c_char_p, C.c_double]),
    "vx_trim": (C.c_int, [_vp]),
    "vx_run_bounds": (C.c_int, [_vp, _cfgp, _resp]),
    "vx_run_params": (C.c_int, [_vp, _cfgp, _dp, _resp, _i32p]),
    "vx_sample_params": (C.c_int, [_vp, _cfgp, C.c_int64, C.c_int64, _dp, _dp]),
    "vx_dump_counts": (C.c_int, [_vp, _cfgp, C.c_int64, C.c_int64, _i32p]),
    "vx_expected_series": (C.c_int, [_vp, _cfgp, _dp, C.c_int64, _dp]),
    "vx_job_begin": (C.c_int, [_vp, _cfgp, C.c_int64, C.c_int64]),
    "vx_job_param_stats": (C.c_int, [_vp, _dp]),
    "vx_job_set_param_stats": (C.c_int, [_vp, _dp]),
    "vx_job_pilot": (C.c_int, [_vp]),
    "vx_job_is_direct": (C.c_int, [_vp]),
    "vx_job_accumulate": (C.c_int, [_vp]),
    "vx_job_acc": (C.c_int, [_vp, C.POINTER(_vp), _i64p]),
    "vx_job_finalize": (C.c_int, [_vp, _resp]),
    "vx_job_end": (C.c_int, [_vp]),
    "vx_host_lerp": (C.c_double, [C.c_double, C.c_double, C.c_double]),
    "vx_host_quantile_index": (None, [C.c_int64, C.c_double, _i64p, _i64p, _dp]),
    "vx_host_count_to_value": (C.c_double, [C.c_int, C.c_int64, C.c_int64]),
    "vx_host_weeks": (C.c_int32, [C.c_int32]),
    "vx_host_rows": (C.c_int64, [C.c_int32]),
    "vx_test_philox": (C.c_int, [_vp, _u32p, _u32p, _u32p]),
    "vx_test_poisson": (C.c_int, [_vp, C.c_double, C.c_int64, C.c_uint64, _i32p]),
    "vx_test_normal": (C.c_int, [_vp, C.c_int64, C.c_uint64, C.POINTER(C.c_float)]),
    "vx_test_select": (C.c_int, [_vp, _i32p, C.c_int64, C.c_int64, _i64p, C.c_int32, _i32p, _i64p]),
    "vx_launch_count": (C.c_int64, [_vp]),
}

_LIB = None


def lib_path() -> str:
    return _build.LIB


def load() -> C.CDLL:
    """Load libvaxmc.so (building it with nvcc if the in-tree .so is missing or stale)."""
    global _LIB
    if _LIB is None:
        path = _build.LIB
        if _build.needs_build():
            try:
                nvcc = _build.find_nvcc()
            except RuntimeError as e:   # a box without nvcc: the shipped in-tree .so is all there is
                nvcc = None
                if not os.path.exists(path):
                    raise RuntimeError(
                        "libvaxmc.so is missing and could not be built; this package has no "
                        f"CPU fallback ({e})") from e
            if nvcc is not None:
                path = _build.build()   # a compile error must surface, never a stale library
            else:
                import warnings

                warnings.warn("libvaxmc.so is older than its sources and nvcc is not available to rebuild it; "
                              "loading the shipped library (its ABI stamp is checked)")
        L = C.CDLL(path)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)  # AttributeError if the library does not export it
            fn.restype = res
            fn.argtypes = args
        # the structs are mirrored by hand above: a library built from another header must not be used
        got = (int(L.vx_abi_version()), int(L.vx_sizeof(0)), int(L.vx_sizeof(1)), int(L.vx_sizeof(2)))
        want = (ABI_VERSION, C.sizeof(vx_config), C.sizeof(vx_result), C.sizeof(vx_nccl_id))
        if got != want:
            raise RuntimeError(f"libvaxmc.so ABI mismatch: library (version, sizeof vx_config, vx_result, "
                               f"vx_nccl_id) = {got}, this binding expects {want}; rebuild the library")
        _LIB = L
    return _LIB


def weeks(T: int) -> int:
    return int(load().vx_host_weeks(int(T)))


def rows(T: int) -> int:
    return int(load().vx_host_rows(int(T)))


def _as_dp(a: np.ndarray):
    return a.ctypes.data_as(_dp)


class ResultBuffers:
    """Caller-owned host buffers for vx_result (lengths T, W, T, T)."""

    def __init__(self, T: int):
        W = weeks(T)
        self.lens = (T, W, T, T)
        self.mean = [np.zeros(n, dtype=np.float64) for n in self.lens]
        self.lower = [np.zeros(n, dtype=np.float64) for n in self.lens]
        self.upper = [np.zeros(n, dtype=np.float64) for n in self.lens]
        self.c = vx_result()
        for s in range(4):
            self.c.mean[s] = _as_dp(self.mean[s])
            self.c.lower[s] = _as_dp(self.lower[s])
            self.c.upper[s] = _as_dp(self.upper[s])


def make_config(bounds, N, T, n_samples, seed, q_lo, q_hi, mode=MODE_STOCHASTIC, direct_max=0,
                pilot_samples=0, max_running_time=None) -> vx_config:
    cfg = vx_config()
    b = np.asarray(bounds, dtype=np.float64).reshape(12)
    for i in range(12):
        cfg.bounds[i] = float(b[i])
    cfg.N = int(N)
    cfg.T = int(T)
    cfg.mode = int(mode)
    cfg.n_samples = int(n_samples)
    cfg.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    cfg.q_lo = float(q_lo)
    cfg.q_hi = float(q_hi)
    cfg.direct_max = int(direct_max)
    cfg.pilot_samples = int(pilot_samples)
    # vx_config: <= 0 means "no limit".  The reference checks `elapsed > max_running_time` after every
    # sample (model.py:187-189), so a budget of 0 is a REAL limit there (one sample, then stop): it
    # is passed on as the smallest positive budget, which stops the job after its first chunk.
    if max_running_time is None:
        cfg.max_running_time = 0.0
    else:
        cfg.max_running_time = max(float(max_running_time), 1e-12)
    return cfg


class Context:
    """One libvaxmc context = one CUDA device of this process."""

    def __init__(self, device: int = 0):
        self._lib = load()
        h = _vp()
        rc = self._lib.vx_create(C.byref(h), int(device))
        if rc != VX_OK:
            raise VaxmcError(rc, self._lib.vx_last_error(None).decode())
        self._h = h
        self.device = int(device)
        self.comm_world = 1

    def close(self):
        if getattr(self, "_h", None):
            self._lib.vx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc: int) -> int:
        if rc < 0:
            raise VaxmcError(rc, self._lib.vx_last_error(self._h).decode())
        return rc

    def set_option(self, name: str, value: float):
        self._ck(self._lib.vx_set_option(self._h, name.encode(), float(value)))

    def trim(self):
        self._ck(self._lib.vx_trim(self._h))

    @property
    def launch_count(self) -> int:
        return int(self._lib.vx_launch_count(self._h))

    # ---- whole jobs
    def run_bounds(self, cfg: vx_config) -> ResultBuffers:
        out = ResultBuffers(cfg.T)
        self._ck(self._lib.vx_run_bounds(self._h, C.byref(cfg), C.byref(out.c)))
        return out

    # ---- sharded over the ranks of an NCCL communicator owned by the library
    def comm_init(self, unique_id: bytes, rank: int, world: int):
        uid = vx_nccl_id()
        C.memmove(C.byref(uid), unique_id, 128)
        self._ck(self._lib.vx_comm_init(self._h, C.byref(uid), int(rank), int(world)))
        self.comm_world = int(world)

    def comm_destroy(self):
        self._ck(self._lib.vx_comm_destroy(self._h))
        self.comm_world = 1

    def comm_info(self):
        r, w, v = C.c_int(), C.c_int(), C.c_int()
        self._ck(self._lib.vx_comm_info(self._h, C.byref(r), C.byref(w), C.byref(v)))
        return int(r.value), int(w.value), int(v.value)

    def run_bounds_comm(self, cfg: vx_config) -> ResultBuffers:
        out = ResultBuffers(cfg.T)
        self._ck(self._lib.vx_run_bounds_comm(self._h, C.byref(cfg), C.byref(out.c)))
        return out

    def run_grid(self, cfgs, wave: int = 0) -> list:
        """A batch of independent jobs (bound boxes) pipelined on this GPU: list of ResultBuffers."""
        n = len(cfgs)
        outs = [ResultBuffers(c.T) for c in cfgs]
        carr = (vx_config * n)(*cfgs)
        rarr = (vx_result * n)()
        for i, o in enumerate(outs):
            C.memmove(C.byref(rarr[i]), C.byref(o.c), C.sizeof(vx_result))
        self._ck(self._lib.vx_run_grid(self._h, carr, n, rarr, int(wave)))
        for i, o in enumerate(outs):
            C.memmove(C.byref(o.c), C.byref(rarr[i]), C.sizeof(vx_result))
        return outs

    def run_params(self, cfg: vx_config, params: np.ndarray, want_counts=False):
        params = np.ascontiguousarray(params, dtype=np.float64).reshape(-1, 6)
        cfg.n_samples = params.shape[0]
        out = ResultBuffers(cfg.T)
        counts = None
        cptr = None
        if want_counts:
            counts = np.empty((rows(cfg.T), params.shape[0]), dtype=np.int32)
            cptr = counts.ctypes.data_as(_i32p)
        self._ck(self._lib.vx_run_params(self._h, C.byref(cfg), _as_dp(params), C.byref(out.c), cptr))
        return (out, counts) if want_counts else out

    # ---- pieces
    def sample_params(self, cfg: vx_config, first: int, count: int):
        params = np.empty((count, 6), dtype=np.float64)
        soft = np.empty(count, dtype=np.float64)
        self._ck(self._lib.vx_sample_params(self._h, C.byref(cfg), int(first), int(count),
                                            _as_dp(params), _as_dp(soft)))
        return params, soft

    def dump_counts(self, cfg: vx_config, first: int, count: int) -> np.ndarray:
        out = np.empty((rows(cfg.T), count), dtype=np.int32)
        self._ck(self._lib.vx_dump_counts(self._h, C.byref(cfg), int(first), int(count),
                                          out.ctypes.data_as(_i32p)))
        return out

    def expected_series(self, cfg: vx_config, params: np.ndarray) -> np.ndarray:
        params = np.ascontiguousarray(params, dtype=np.float64).reshape(-1, 6)
        n = params.shape[0]
        out = np.empty((4, n, cfg.T), dtype=np.float64)
        self._ck(self._lib.vx_expected_series(self._h, C.byref(cfg), _as_dp(params), n, _as_dp(out)))
        return out

    # ---- staged job (multi-GPU path)
    def job_begin(self, cfg: vx_config, shard_first: int, shard_count: int):
        self._ck(self._lib.vx_job_begin(self._h, C.byref(cfg), int(shard_first), int(shard_count)))

    def job_param_stats(self) -> np.ndarray:
        st = np.zeros(4, dtype=np.float64)
        self._ck(self._lib.vx_job_param_stats(self._h, _as_dp(st)))
        return st

    def job_set_param_stats(self, st: np.ndarray) -> bool:
        st = np.ascontiguousarray(st, dtype=np.float64)
        return self._ck(self._lib.vx_job_set_param_stats(self._h, _as_dp(st))) == 1

    def job_pilot(self):
        self._ck(self._lib.vx_job_pilot(self._h))

    def job_is_direct(self) -> bool:
        return bool(self._lib.vx_job_is_direct(self._h))

    def job_accumulate(self):
        self._ck(self._lib.vx_job_accumulate(self._h))

    def job_acc(self):
        p = _vp()
        n = C.c_int64()
        self._ck(self._lib.vx_job_acc(self._h, C.byref(p), C.byref(n)))
        return (p.value or 0), int(n.value)

    def job_finalize(self, out: ResultBuffers) -> int:
        return self._ck(self._lib.vx_job_finalize(self._h, C.byref(out.c)))

    def job_end(self):
        self._ck(self._lib.vx_job_end(self._h))

    # ---- device self-tests
    def test_philox(self, ctr, key):
        c = (C.c_uint32 * 4)(*[int(x) for x in ctr])
        k = (C.c_uint32 * 2)(*[int(x) for x in key])
        o = (C.c_uint32 * 4)()
        self._ck(self._lib.vx_test_philox(self._h, c, k, o))
        return [int(x) for x in o]

    def test_poisson(self, lam: float, n: int, seed: int) -> np.ndarray:
        out = np.empty(n, dtype=np.int32)
        self._ck(self._lib.vx_test_poisson(self._h, float(lam), int(n), int(seed),
                                           out.ctypes.data_as(_i32p)))
        return out

    def test_normal(self, n: int, seed: int) -> np.ndarray:
        out = np.empty(n, dtype=np.float32)
        self._ck(self._lib.vx_test_normal(self._h, int(n), int(seed),
                                          out.ctypes.data_as(C.POINTER(C.c_float))))
        return out

    def test_select(self, data: np.ndarray, ranks):
        data = np.ascontiguousarray(data, dtype=np.int32)
        r, n = data.shape
        rk = np.ascontiguousarray(ranks, dtype=np.int64)
        out = np.empty((r, len(rk)), dtype=np.int32)
        sums = np.empty(r, dtype=np.int64)
        self._ck(self._lib.vx_test_select(self._h, data.ctypes.data_as(_i32p), r, n,
                                          rk.ctypes.data_as(_i64p), len(rk),
                                          out.ctypes.data_as(_i32p), sums.ctypes.data_as(_i64p)))
        return out, sums


def comm_unique_id() -> bytes:
    """ncclGetUniqueId through the library (one rank makes it, the rendezvous ships it)."""
    L = load()
    uid = vx_nccl_id()
    rc = L.vx_comm_unique_id(C.byref(uid))
    if rc != VX_OK:
        raise VaxmcError(rc, L.vx_last_error(None).decode())
    return bytes(C.string_at(C.byref(uid), 128))


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    f, c = C.c_int64(), C.c_int64()
    load().vx_host_shard_range(int(n), int(rank), int(world), C.byref(f), C.byref(c))
    return int(f.value), int(c.value)


_CONTEXTS: dict[int, Context] = {}


def default_context(device: int = 0) -> Context:
    """Process-wide context cache (one per device)."""
    ctx = _CONTEXTS.get(device)
    if ctx is None:
        ctx = _CONTEXTS[device] = Context(device)
    return ctx
