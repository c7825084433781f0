"""oracle/ref_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes front-end of oracle/vaxmodel_oracle.c plus the numpy/pandas half of the
reference's reduction (np.quantile / ndarray.mean / pd.date_range are the very
third-party calls the reference makes at model.py:171, 211, 220-223, so the oracle
calls them directly instead of restating them).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  Parity pin: see the header of vaxmodel_oracle.c and
tests/test_oracle_pin.py (fixtures generated from the reference itself by
oracle/make_golden.py).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from collections import defaultdict

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

SERIES = (
    "people_vaccinated_per_hundred",
    "daily_vaccinations_per_million",
    "cum_number_vac_received_per_hundred",
    "vaccines_in_stock_per_hundred",
)


def build(force: bool = False) -> str:
    """Compile libvaxoracle.so with the committed Makefile (gcc only)."""
    so = os.path.join(_HERE, "libvaxoracle.so")
    src = os.path.join(_HERE, "vaxmodel_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return so


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        c_d = ctypes.POINTER(ctypes.c_double)
        c_i32 = ctypes.POINTER(ctypes.c_int32)
        L.vo_sizeof_mt.restype = ctypes.c_int
        L.vo_mt_seed.argtypes = [ctypes.c_void_p, ctypes.c_uint32]
        L.vo_mt_double.argtypes = [ctypes.c_void_p]
        L.vo_mt_double.restype = ctypes.c_double
        L.vo_uniform.argtypes = [ctypes.c_void_p, ctypes.c_double, ctypes.c_double]
        L.vo_uniform.restype = ctypes.c_double
        L.vo_poisson.argtypes = [ctypes.c_void_p, ctypes.c_double]
        L.vo_poisson.restype = ctypes.c_int64
        L.vo_sample_param_combinations.argtypes = [ctypes.c_void_p, c_d, ctypes.c_int64, c_d, c_d]
        L.vo_sample_param_combinations.restype = ctypes.c_int64
        L.vo_run_single_realization.argtypes = [
            ctypes.c_void_p, ctypes.c_int, c_d, ctypes.c_int64, ctypes.c_int64, c_d, c_i32]
        L.vo_run_single_realization.restype = ctypes.c_int
        L.vo_weekly_series.argtypes = [c_d, ctypes.c_int64, c_d]
        L.vo_run_sampling_sequential.argtypes = [
            ctypes.c_void_p, ctypes.c_int, c_d, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, c_d]
        L.vo_run_sampling_sequential.restype = ctypes.c_int
        L.vo_run_bulk_counts.argtypes = [
            ctypes.c_uint32, c_d, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, c_i32, c_i32,
            ctypes.c_int]
        L.vo_run_bulk_counts.restype = ctypes.c_int
        _LIB = L
    return _LIB


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _ip(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))


class MT:
    """numpy's legacy global RandomState (MT19937), restated.  MT(seed) == np.random.seed(seed)."""

    def __init__(self, seed: int):
        self._buf = ctypes.create_string_buffer(lib().vo_sizeof_mt())
        self.seed(seed)

    def seed(self, seed: int):
        lib().vo_mt_seed(self._buf, ctypes.c_uint32(seed & 0xFFFFFFFF))

    @property
    def ptr(self):
        return ctypes.cast(self._buf, ctypes.c_void_p)

    def random_sample(self) -> float:
        return lib().vo_mt_double(self.ptr)

    def uniform(self, lo: float, hi: float) -> float:
        return lib().vo_uniform(self.ptr, lo, hi)

    def poisson(self, lam: float) -> int:
        return lib().vo_poisson(self.ptr, lam)


# --------------------------------------------------------------------------- A1
def sample_param_combinations(mt: MT, p_pro_bounds, p_anti_bounds, pressure_bounds, tau_bounds,
                              nv_0_bounds, nv_max_bounds, n_rep: int):
    """model.py:231-303.  Returns (params [n,6] float64, p_soft_no [n]) or (None, None)."""
    bounds = np.ascontiguousarray(
        np.concatenate([np.asarray(b, dtype=np.float64).reshape(2) for b in (
            p_pro_bounds, p_anti_bounds, pressure_bounds, tau_bounds, nv_0_bounds, nv_max_bounds)]))
    params = np.empty((n_rep, 6), dtype=np.float64)
    soft = np.empty(n_rep, dtype=np.float64)
    rej = lib().vo_sample_param_combinations(mt.ptr, _dp(bounds), n_rep, _dp(params), _dp(soft))
    if rej < 0:
        return None, None
    return params, soft


# ------------------------------------------------------------------------ A2-A6
def run_single_realization(mt, mode: int, p, max_day_number: int, N: int, want_counts=False):
    """model.py:29-136.  mode 0 stochastic (needs mt), mode 1 expected value."""
    p = np.ascontiguousarray(np.asarray(p, dtype=np.float64).reshape(6))
    T = int(max_day_number)
    out = np.empty((4, T), dtype=np.float64)
    cnt = np.empty((4, T), dtype=np.int32) if want_counts else None
    rc = lib().vo_run_single_realization(
        mt.ptr if mt is not None else None, mode, _dp(p), T, int(N), _dp(out),
        _ip(cnt) if want_counts else None)
    if rc:
        raise AssertionError("p_pro + p_anti <= 1.0  (model.py:63)")
    data = {k: out[i] for i, k in enumerate(SERIES)}
    if want_counts:
        return data, cnt
    return data


def samples_sequential(mt, mode: int, params, T: int, N: int) -> np.ndarray:
    """model.py:176-193 -> [4][n][T] float64, one sequential random stream."""
    params = np.ascontiguousarray(params, dtype=np.float64)
    n = params.shape[0]
    out = np.empty((4, n, T), dtype=np.float64)
    rc = lib().vo_run_sampling_sequential(
        mt.ptr if mt is not None else None, mode, _dp(params), n, T, int(N), _dp(out))
    if rc:
        raise AssertionError("p_pro + p_anti <= 1.0  (model.py:63)")
    return out


def weekly_from_daily(daily_samples: np.ndarray) -> np.ndarray:
    """model.py:196-214 on an [n,T] array: out[:,k] = x[:,7k] + (x[:,7k-30] if 7k>=30 else 0)."""
    n, T = daily_samples.shape
    W = (T + 6) // 7
    out = np.empty((n, W), dtype=np.float64)
    row = np.empty(W, dtype=np.float64)
    for i in range(n):
        src = np.ascontiguousarray(daily_samples[i])
        lib().vo_weekly_series(_dp(src), T, _dp(row))
        out[i] = row
    return out


def reduce_samples(samples4: np.ndarray, start_date, end_date, CI: float):
    """model.py:193-228 given the stacked [4][n][T] samples.  CI is the fraction run_sampling sees."""
    import pandas as pd

    dates = pd.date_range(start_date, end_date, freq="1D")
    T = len(dates)
    assert samples4.shape[2] == T
    data = {k: {"dates": dates, "samples": samples4[i]} for i, k in enumerate(SERIES)}
    k = "daily_vaccinations_per_million"
    data[k]["samples"] = weekly_from_daily(data[k]["samples"])
    data[k]["dates"] = pd.date_range(start=start_date, end=end_date, freq="7D")
    data_CI = defaultdict(dict)
    for k in SERIES:
        s = data[k]["samples"]
        q = np.quantile(s, [(1 - CI) / 2.0, (1 + CI) / 2.0], axis=0)
        data_CI[k]["upper"] = q[1]
        data_CI[k]["lower"] = q[0]
        data_CI[k]["mean"] = s.mean(axis=0)
        data_CI[k]["dates"] = data[k]["dates"]
    data_CI["number_finished_samples"] = samples4.shape[1]
    return data_CI


def run_sampling(mt, mode, params, start_date, end_date, CI, N):
    """model.py:139-228 (no timeout: the oracle always finishes)."""
    import pandas as pd

    T = len(pd.date_range(start_date, end_date, freq="1D"))
    s4 = samples_sequential(mt, mode, params, T, N)
    return reduce_samples(s4, start_date, end_date, CI)


def agnostics_message(p_soft_no_values) -> str:
    """model.py:339-349."""
    msg = "Agnosticts: "
    v = 100 * np.array(p_soft_no_values)
    a = max(np.mean(v) - np.std(v), 0)
    b = np.mean(v) + np.std(v)
    a_str = "{0:.0f}".format(a)
    b_str = "{0:.0f}".format(b)
    if abs(a - b) < 1:
        msg += a_str + "%"
    else:
        msg += a_str + " - " + b_str + "%"
    return msg


ERR_BOUNDS = ("ERROR: The pertentages of pro- and anti-vaccines are simultaneously too high. "
              "Please reduce them.")


def run_model(mt, p_pro_bounds, p_anti_bounds, pressure_bounds, tau_bounds, nv_0_bounds,
              nv_max_bounds, CI, n_rep, N, date_range, mode=0):
    """model.py:306-368."""
    params, soft = sample_param_combinations(
        mt, np.array(p_pro_bounds) / 100, np.array(p_anti_bounds) / 100,
        np.array(pressure_bounds), np.array(tau_bounds), np.array(nv_0_bounds) / 100,
        np.array(nv_max_bounds) / 100, n_rep)
    if params is None:
        return None, ERR_BOUNDS, "Agnosticts: "
    msg = agnostics_message(soft)
    res = run_sampling(mt, mode, params, date_range["start_date"], date_range["end_date"],
                       CI / 100, N)
    return res, "", msg


# -------------------------------------------------------------- bulk (KS / timing)
def bulk_counts(seed: int, params, T: int, N: int, threads: int | None = None,
                want_daily=True, want_weekly=True):
    """Independent-stream fan-out of model.py:29-136 over host threads.

    Returns (daily [3][T][n] int32 or None, weekly [W][n] int32 or None); daily rows are
    n_vaccinated, cum_number_vac_received, vaccines_stock."""
    params = np.ascontiguousarray(params, dtype=np.float64)
    n = params.shape[0]
    W = (T + 6) // 7
    threads = threads or (os.cpu_count() or 1)
    daily = np.empty((3, T, n), dtype=np.int32) if want_daily else None
    weekly = np.empty((W, n), dtype=np.int32) if want_weekly else None
    rc = lib().vo_run_bulk_counts(
        ctypes.c_uint32(seed & 0xFFFFFFFF), _dp(params), n, T, int(N),
        _ip(daily) if want_daily else None, _ip(weekly) if want_weekly else None, int(threads))
    if rc:
        raise RuntimeError(f"vo_run_bulk_counts rc={rc}")
    return daily, weekly
