"""Host-side mirror of the reference's model.py interface for the Monte Carlo hot path.

Same names, argument meaning, return shapes and error strings as
/root/reference/model.py, with the arithmetic done by libvaxmc.so (hand-written sm_100a
CUDA behind the C ABI of include/vaxmc.h):

    run_model                  model.py:306-368   (the drop-in boundary)
    run_sampling               model.py:139-228   (explicit parameter tuples)
    sample_param_combinations  model.py:231-303
    run_single_realization     model.py:29-136
    main / SplitArgsStr / SplitArgsFloat   model.py:371-546  (same flags and parsing)

What differs, on purpose:
  * randomness is Philox4x32-10 keyed by (seed, global sample id) instead of numpy's global
    MT19937 stream; `seed(12345)` plays the role of `np.random.seed(12345)` (app.py:442).
  * parameter tuples are drawn on the device and never materialised for run_model (48 B per
    sample would be 48 GB at 1e9 samples).
  * extra OPTIONAL keyword arguments (seed=, device=, mode=) and CLI flags (--seed, --device,
    --expected, --json_out); the reference's own flags and positional arguments are unchanged.

pandas is used for dates only (model.py:171, 211), exactly as in the reference.
"""
from __future__ import annotations

import argparse
import functools
import json
import os
import time
from collections import OrderedDict

import numpy as np
import pandas as pd

from . import _native as nat

SERIES = nat.SERIES

MSG_ERR_BOUNDS = ("ERROR: The pertentages of pro- and anti-vaccines are simultaneously too high. "
                  "Please reduce them.")  # model.py:351 (typo kept: the UI shows it verbatim)

# ------------------------------------------------------------------------- seeding
_seed_state = {"base": None, "calls": 0}


def seed(value: int | None):
    """Counterpart of np.random.seed(value) for this engine (app.py:442 seeds per callback)."""
    _seed_state["base"] = None if value is None else int(value)
    _seed_state["calls"] = 0


def _next_seed(explicit: int | None, collective: bool = False) -> tuple[int, bool]:
    """(Philox key, deterministic?) for the next job.  `collective`: the caller is an entry point
    that every rank of a torchrun job enters (run_model, run_model_quantiles); an entropy key is
    then rank 0's on every rank (distributed.broadcast_seed) -- one job has one key."""
    if explicit is not None:
        return int(explicit) & 0xFFFFFFFFFFFFFFFF, True
    base = _seed_state["base"]
    if base is None:
        key = int.from_bytes(os.urandom(8), "little")
        if collective:
            from . import distributed as dist_

            key = dist_.broadcast_seed(key)
        return key, False
    k = _seed_state["calls"]
    _seed_state["calls"] = k + 1
    # first job after seed(v) uses v itself; later jobs a fixed odd-multiplier walk from it
    return (base + k * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF, True


# --------------------------------------------------------------------------- dates
@functools.lru_cache(maxsize=64)
def _dates(start_date, end_date):
    daily = pd.date_range(start_date, end_date, freq="1D")            # model.py:171
    weekly = pd.date_range(start=start_date, end=end_date, freq="7D")  # model.py:211
    return daily, weekly


def _pack(out: nat.ResultBuffers, daily, weekly) -> dict:
    """vx_result -> the mapping run_sampling returns (model.py:216-228)."""
    data_CI = {}
    for s, k in enumerate(SERIES):
        data_CI[k] = {
            "upper": out.upper[s],
            "lower": out.lower[s],
            "mean": out.mean[s],
            "dates": weekly if s == 1 else daily,
        }
    data_CI["number_finished_samples"] = int(out.c.n_finished)
    return data_CI


# ------------------------------------------------------------- result cache (N1)
# The reference memoises run_sampling on the whole parameter tuple (model.py:139), which is
# what lets app.py re-plot without recomputing (app.py:435-447).  The tuple does not exist
# here, so the key is everything that determines it: bounds, n_rep, N, dates, CI, seed, mode.
_CACHE: "OrderedDict[tuple, tuple]" = OrderedDict()
_CACHE_MAX = 10


def cache_clear():
    _CACHE.clear()


def _bounds12(p_pro_bounds, p_anti_bounds, pressure_bounds, tau_bounds, nv_0_bounds, nv_max_bounds):
    return np.concatenate([np.asarray(b, dtype=np.float64).reshape(2) for b in (
        p_pro_bounds, p_anti_bounds, pressure_bounds, tau_bounds, nv_0_bounds, nv_max_bounds)])


def _agnostics_message(soft_mean_frac: float, soft_std_frac: float) -> str:
    """model.py:339-349 from the device-side moments of p_soft_no."""
    msg = "Agnosticts: "
    m, s = 100 * soft_mean_frac, 100 * soft_std_frac
    a = max(m - s, 0)
    b = m + s
    a_str = "{0:.0f}".format(a)
    b_str = "{0:.0f}".format(b)
    if abs(a - b) < 1:
        msg += a_str + "%"
    else:
        msg += a_str + " - " + b_str + "%"
    return msg


# --------------------------------------------------------------------- the boundary
def run_model(
    p_pro_bounds, p_anti_bounds, pressure_bounds,      # population: % pro, % anti, pressure
    tau_bounds, nv_0_bounds, nv_max_bounds,            # supply: doubling time, first and max delivery (%)
    CI, n_rep, N, date_range,                          # CI in percent, samples, population, dates
    max_running_time=None,
    *,
    seed=None,
    device=None,
    mode="stochastic",
    direct_max=0,
    pilot_samples=0,
):
    """Drop-in for the reference's run_model (model.py:306-368).

    Returns (model_results | None, msg_error, msg_agnostics_pct) with the reference's keys,
    shapes and strings.  CI is in PERCENT and is divided by 100 here, as the reference does
    (model.py:358) -- so the CLI default --CI 0.95 gives the 0.49525/0.50475 quantiles, and
    the Dash slider's 95 the 0.025/0.975 ones.
    """
    msg_agnostics_pct = "Agnosticts: "
    msg_error = ""

    # some sliders use values 0-100 (model.py:328-336)
    bounds = _bounds12(
        np.array(p_pro_bounds) / 100,
        np.array(p_anti_bounds) / 100,
        np.array(pressure_bounds),
        np.array(tau_bounds),
        np.array(nv_0_bounds) / 100,
        np.array(nv_max_bounds) / 100,
    )
    ci_frac = CI / 100
    daily, weekly = _dates(date_range["start_date"], date_range["end_date"])
    T = len(daily)
    key_seed, deterministic = _next_seed(seed, collective=True)
    vx_mode = nat.MODE_EXPECTED if mode == "expected" else nat.MODE_STOCHASTIC

    cache_key = None
    if deterministic and max_running_time is None:
        cache_key = (tuple(bounds.tolist()), int(n_rep), int(N), str(daily[0]), str(daily[-1]),
                     float(CI), key_seed, vx_mode, int(direct_max), int(pilot_samples))
        hit = _CACHE.get(cache_key)
        if hit is not None:
            _CACHE.move_to_end(cache_key)
            return hit

    cfg = nat.make_config(bounds, N, T, n_rep, key_seed, (1 - ci_frac) / 2.0, (1 + ci_frac) / 2.0,
                          mode=vx_mode, direct_max=direct_max, pilot_samples=pilot_samples,
                          max_running_time=max_running_time)

    from . import distributed as dist_

    out = dist_.run_job(cfg, device=device)

    if out.c.aborted:
        msg_error = MSG_ERR_BOUNDS
        ret = (None, msg_error, msg_agnostics_pct)
    else:
        msg_agnostics_pct = _agnostics_message(out.c.soft_no_mean, out.c.soft_no_std)
        model_results = _pack(out, daily, weekly)
        if max_running_time is not None:
            number_finished_samples = model_results["number_finished_samples"]
            if number_finished_samples < n_rep:
                msg_error = (f"ERROR: Maximum computation time of {max_running_time}s exceeded. "
                             f"Only {number_finished_samples} of the desired {n_rep} Monte Carlo "
                             f"runs were performed.")
        ret = (model_results, msg_error, msg_agnostics_pct)

    if cache_key is not None:
        _CACHE[cache_key] = ret
        while len(_CACHE) > _CACHE_MAX:
            _CACHE.popitem(last=False)
    return ret


def run_model_grid(boxes, CI, n_rep, N, date_range, *, seeds=None, device=None, workers=None,
                   mode="stochastic", direct_max=0, pilot_samples=0, wave=0):
    """Batched multi-scenario form of run_model (BASELINE.json configs[4]).

    `boxes` is a sequence of (p_pro_bounds, p_anti_bounds, pressure_bounds, tau_bounds,
    nv_0_bounds, nv_max_bounds) in run_model's user units; returns a list of run_model triples in
    the same order, each identical to what run_model(box, ..., seed=seeds[i]) returns.  The boxes
    go to the library in ONE call (vx_run_grid): it pipelines them over sub-contexts so that the
    latency-bound pilots of the next boxes run in the shadow of the current boxes' folds and the
    host never waits on a kernel it has just queued.  Under torchrun the caller partitions the
    boxes over ranks (no collective is needed).  `workers` is accepted for compatibility with the
    first version (host threads) and ignored."""
    daily, weekly = _dates(date_range["start_date"], date_range["end_date"])
    T = len(daily)
    ci_frac = CI / 100
    vx_mode = nat.MODE_EXPECTED if mode == "expected" else nat.MODE_STOCHASTIC
    boxes = list(boxes)
    if seeds is None:
        seeds = [_next_seed(None)[0] for _ in boxes]
    cfgs = []
    for b, sd in zip(boxes, seeds):
        bounds = _bounds12(np.array(b[0]) / 100, np.array(b[1]) / 100, np.array(b[2]), np.array(b[3]),
                           np.array(b[4]) / 100, np.array(b[5]) / 100)
        cfgs.append(nat.make_config(bounds, N, T, n_rep, sd, (1 - ci_frac) / 2.0, (1 + ci_frac) / 2.0,
                                    mode=vx_mode, direct_max=direct_max, pilot_samples=pilot_samples))
    outs = nat.default_context(_device_index(device)).run_grid(cfgs, wave=wave) if cfgs else []
    res = []
    for out in outs:
        if out.c.aborted:
            res.append((None, MSG_ERR_BOUNDS, "Agnosticts: "))
        else:
            res.append((_pack(out, daily, weekly), "", _agnostics_message(out.c.soft_no_mean, out.c.soft_no_std)))
    return res


def run_model_quantiles(p_pro_bounds, p_anti_bounds, pressure_bounds, tau_bounds, nv_0_bounds,
                        nv_max_bounds, quantiles, n_rep, N, date_range, *, seed=None, device=None,
                        direct_max=0, pilot_samples=0):
    """Fan chart: arbitrary quantile FRACTIONS of the same Monte Carlo sample.

    The engine extracts two exact quantiles per job; because a sample's trajectory depends only
    on (seed, sample id), jobs that share the seed see the very same trajectories, so the
    fractions are served pairwise by ceil(len/2) jobs and are mutually consistent (monotone in q,
    equal to np.quantile over one common sample).  Returns (model_results, msg_error,
    msg_agnostics_pct) where each series carries "mean", "dates" and "quantiles": {q: array}.
    """
    qs = sorted(float(q) for q in quantiles)
    if not qs or qs[0] < 0.0 or qs[-1] > 1.0:
        raise ValueError("quantiles must be fractions in [0, 1]")
    key_seed, _ = _next_seed(seed, collective=True)
    bounds = _bounds12(np.array(p_pro_bounds) / 100, np.array(p_anti_bounds) / 100,
                       np.array(pressure_bounds), np.array(tau_bounds), np.array(nv_0_bounds) / 100,
                       np.array(nv_max_bounds) / 100)
    daily, weekly = _dates(date_range["start_date"], date_range["end_date"])
    from . import distributed as dist_

    pairs = [(qs[i], qs[-1 - i]) for i in range((len(qs) + 1) // 2)]     # outermost pair first
    results = None
    msg = "Agnosticts: "
    for q_lo, q_hi in pairs:
        cfg = nat.make_config(bounds, N, len(daily), n_rep, key_seed, q_lo, q_hi, direct_max=direct_max,
                              pilot_samples=pilot_samples)
        out = dist_.run_job(cfg, device=device)
        if out.c.aborted:
            return None, MSG_ERR_BOUNDS, "Agnosticts: "
        if results is None:
            msg = _agnostics_message(out.c.soft_no_mean, out.c.soft_no_std)
            results = {k: {"mean": out.mean[s], "dates": weekly if s == 1 else daily, "quantiles": {}}
                       for s, k in enumerate(SERIES)}
            results["number_finished_samples"] = int(out.c.n_finished)
        for s, k in enumerate(SERIES):
            results[k]["quantiles"][q_lo] = out.lower[s]
            results[k]["quantiles"][q_hi] = out.upper[s]
    return results, "", msg


def run_sampling(params, start_date, end_date, CI, N, max_running_time=None, *, seed=None,
                 device=None, mode="stochastic", direct_max=0, pilot_samples=0):
    """model.py:139-228 over EXPLICIT parameter tuples (CI is the fraction, as there)."""
    p = np.asarray(params, dtype=np.float64).reshape(-1, 6)
    daily, weekly = _dates(start_date, end_date)
    key_seed, _ = _next_seed(seed)
    vx_mode = nat.MODE_EXPECTED if mode == "expected" else nat.MODE_STOCHASTIC
    cfg = nat.make_config(np.zeros(12), N, len(daily), p.shape[0], key_seed, (1 - CI) / 2.0,
                          (1 + CI) / 2.0, mode=vx_mode, direct_max=direct_max,
                          pilot_samples=pilot_samples, max_running_time=max_running_time)
    ctx = nat.default_context(_device_index(device))
    out = ctx.run_params(cfg, p)
    return _pack(out, daily, weekly)


def sample_param_combinations(p_pro_bounds, p_anti_bounds, pressure_bounds, tau_bounds,
                              nv_0_bounds, nv_max_bounds, n_rep, *, seed=None, device=None):
    """model.py:231-303: (tuple of n_rep 6-tuples, tuple of p_soft_no) or (None, None)."""
    bounds = _bounds12(p_pro_bounds, p_anti_bounds, pressure_bounds, tau_bounds, nv_0_bounds,
                       nv_max_bounds)
    key_seed, _ = _next_seed(seed)
    cfg = nat.make_config(bounds, 1, 1, n_rep, key_seed, 0.5, 0.5)
    ctx = nat.default_context(_device_index(device))
    ctx.job_begin(cfg, 0, n_rep)
    try:
        aborted = ctx.job_set_param_stats(ctx.job_param_stats())
    finally:
        ctx.job_end()
    if aborted:
        return None, None
    params, soft = ctx.sample_params(cfg, 0, n_rep)
    return tuple(tuple(float(x) for x in row) for row in params), tuple(float(x) for x in soft)


def run_single_realization(p_pro, p_anti, pressure, tau, nv_0, nv_max, max_day_number, N, *,
                           seed=None, device=None, mode="stochastic"):
    """model.py:29-136 for one parameter tuple: dict of four per-day lists."""
    assert p_pro + p_anti <= 1.0
    T = int(max_day_number)
    p = np.array([[p_pro, p_anti, pressure, tau, nv_0, nv_max]], dtype=np.float64)
    key_seed, _ = _next_seed(seed)
    ctx = nat.default_context(_device_index(device))
    if mode == "expected":
        cfg = nat.make_config(np.zeros(12), N, T, 1, key_seed, 0.5, 0.5, mode=nat.MODE_EXPECTED)
        s = ctx.expected_series(cfg, p)
        return {k: list(s[i, 0]) for i, k in enumerate(SERIES)}
    cfg = nat.make_config(np.zeros(12), N, T, 1, key_seed, 0.5, 0.5)
    _, counts = ctx.run_params(cfg, p, want_counts=True)
    W = nat.weeks(T)
    n_vacc = counts[0:T, 0].astype(np.int64)
    received = np.repeat(counts[T + W:T + 2 * W, 0].astype(np.int64), 7)[:T]
    stock = counts[T + 2 * W:, 0].astype(np.int64)
    dv = np.diff(n_vacc, prepend=0)
    return {
        SERIES[0]: [(int(c) / N) * 100 for c in n_vacc],
        SERIES[1]: [int(c) * 1e6 / N for c in dv],
        SERIES[2]: [int(c) * 100 / N for c in received],
        SERIES[3]: [int(c) * 100 / N for c in stock],
    }


def _device_index(device) -> int:
    if device is not None:
        return int(device)
    return int(os.environ.get("LOCAL_RANK", "0")) if "LOCAL_RANK" in os.environ else 0


# ------------------------------------------------------------------------------ CLI
class SplitArgsStr(argparse.Action):
    def __call__(self, parser, namespace, values, option_string=None):
        values = values.split(",")
        if len(values) == 1:
            values = [values[0], values[0]]
        setattr(namespace, self.dest, values)


class SplitArgsFloat(argparse.Action):
    def __call__(self, parser, namespace, values, option_string=None):
        values = [float(x) for x in values.split(",")]
        if len(values) == 1:
            values = [values[0], values[0]]
        setattr(namespace, self.dest, values)


def build_parser() -> argparse.ArgumentParser:
    """The reference's flags (model.py:419-500): same names, same split-on-comma actions, same types
    and required-ness, the same defaults for --mc_samples (100) and --CI (0.95).  The defaults of
    the seven `required=True` flags can never take effect in either program and the help texts
    are worded independently.  Optional additions follow (none changes the reference's flags)."""
    parser = argparse.ArgumentParser(
        description="Model of the COVID-19 vaccination campaign (B200-native Monte Carlo path)")
    parser.add_argument("--pro", type=str, action=SplitArgsFloat, default=[30.0, 40.0], required=True,
                        help="comma-separated upper and lower bounds for the probability that a "
                             "certain person belongs to the pro-vaccines group")
    parser.add_argument("--anti", type=str, action=SplitArgsFloat, default=[30.0, 40.0], required=True,
                        help="comma-separated upper and lower bounds for the probability that a "
                             "specific person belongs to the anti-vaccines group")
    parser.add_argument("--pressure", type=str, action=SplitArgsFloat, default=[0.02, 0.025],
                        required=True,
                        help="comma-separated upper and lower bounds for the strength of the social "
                             "pressure effect")
    parser.add_argument("--dupl_time", type=str, action=SplitArgsFloat, default=[3, 4], required=True,
                        help="comma-separated upper and lower bounds for the duplication time of "
                             "the weekly arriving vaccines")
    parser.add_argument("--init_stock", type=str, action=SplitArgsFloat, default=[0.2, 0.24],
                        required=True,
                        help="comma-separated upper and lower bounds for the initial stock of "
                             "vaccines, measured as a percentege of the population size")
    parser.add_argument("--max_delivery", type=str, action=SplitArgsFloat, default=[10.0, 10.0],
                        required=True,
                        help="comma-separated upper and lower bounds for the maximum weekly "
                             "delivery capacity, measured as a percentage over the population size")
    parser.add_argument("--mc_samples", type=int, default=100,
                        help="number of Monte Carlo samples, i.e. of parameter combinations drawn")
    parser.add_argument("--date_range", type=str, action=SplitArgsStr,
                        default=["2020-12-30", "2021-12-1"], required=True,
                        help="comma-separated starting and ending dates (optional)")
    parser.add_argument("--CI", type=float, default=0.95,
                        help="value of the quantile used for establishing the confidence intervals")
    # optional additions (not in the reference)
    parser.add_argument("--seed", type=int, default=None, help="Philox seed (default: entropy)")
    parser.add_argument("--device", type=int, default=None, help="CUDA device index")
    parser.add_argument("--expected", action="store_true",
                        help="deterministic expected-value mode (Poisson draws replaced by means)")
    parser.add_argument("--json_out", type=str, default=None,
                        help="write model_results as JSON (dates as ISO strings) to this file")
    parser.add_argument("--no_plot", action="store_true", help="do not try to open the plot")
    return parser


def results_to_jsonable(model_results: dict) -> dict:
    """The Dash-store wire format of model_results (app.py:240, 383, 483): lists + strings."""
    out = {}
    for k, v in model_results.items():
        if isinstance(v, dict):
            out[k] = {f: ([str(d) for d in x] if f == "dates" else np.asarray(x).tolist())
                      for f, x in v.items()}
        else:
            out[k] = int(v)
    return out


def main(argv=None):
    args = build_parser().parse_args(argv)
    p_pro_bounds = args.pro
    p_anti_bounds = args.anti
    pressure_bounds = args.pressure
    tau_bounds = args.dupl_time
    nv_0_bounds = args.init_stock
    nv_max_bounds = args.max_delivery
    n_rep = args.mc_samples
    N = 50000  # model.py:512
    start_date, end_date = args.date_range
    CI = args.CI
    date_range = dict(start_date=start_date, end_date=end_date)

    t0 = time.perf_counter()
    model_results, msg_error, msg_agnostics_pct = run_model(
        p_pro_bounds, p_anti_bounds, pressure_bounds, tau_bounds, nv_0_bounds, nv_max_bounds,
        CI, n_rep, N, date_range, seed=args.seed, device=args.device,
        mode="expected" if args.expected else "stochastic")
    dt = time.perf_counter() - t0

    if msg_error != "":
        print(msg_error)
    if model_results is None:
        return 1
    print(msg_agnostics_pct)
    T = len(model_results[SERIES[0]]["mean"])
    print(f"{model_results['number_finished_samples']} samples x {T} days in {dt:.3f} s")
    for k in SERIES:
        r = model_results[k]
        print(f"{k}: last date {r['dates'][-1].date()} mean {r['mean'][-1]:.6g} "
              f"[{r['lower'][-1]:.6g}, {r['upper'][-1]:.6g}]")
    if args.json_out:
        with open(args.json_out, "w") as f:
            json.dump(results_to_jsonable(model_results), f)
    if not args.no_plot:
        try:  # the reference's plot.py, if it is importable (needs plotly)
            from plot import plot_model_results  # type: ignore

            fig = plot_model_results(model_results, CI)
            fig.show(renderer="browser")
        except Exception:
            pass
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
