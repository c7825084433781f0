"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle.

Bars (BASELINE.json north_star):
  (1) expected-value mode vs the oracle/reference          <= 1e-5 relative (we see ~1e-15)
  (2) stochastic mode on identical parameter tuples         two-sample KS per week
  (3) integer results                                       bit-exact (order statistics, sums,
                                                            direct vs streamed, 1 vs k shards)
Everything deterministic given the parameters (arrivals -> `received` rows) is bit-exact
against the oracle.
"""
import os

import numpy as np
import pytest
from scipy import stats

import philox_ref as pr
import ref_oracle as ro
from conftest import USA_DATES, usa_bounds_frac, wide_bounds_frac

pytestmark = pytest.mark.gpu

SERIES = ro.SERIES
USA_T, USA_W = 564, 81


def _cfg(vax, bounds, n, seed=12345, T=USA_T, N=50000, q=(0.025, 0.975), **kw):
    return vax._native.make_config(bounds, N, T, n, seed, q[0], q[1], **kw)


# ------------------------------------------------------------------ building blocks
def test_philox_known_answers(ctx):
    for c, k, e in pr.KAT:
        assert ctx.test_philox(c, k) == list(e)


def test_device_parameter_draw_matches_numpy_restatement(vax, ctx):
    for bounds, seed in ((usa_bounds_frac(), 12345), (wide_bounds_frac(), 7),
                         (np.array([0.5, 0.9, 0.3, 0.6, 0.0, 0.1, 3, 4, 0.002, 0.0024, 0.1, 0.1]), 99)):
        cfg = _cfg(vax, bounds, 5000, seed=seed)
        params, soft = ctx.sample_params(cfg, 1000, 5000)
        ep, es, rej = pr.draw_params(bounds, seed, 1000, 5000)
        assert np.array_equal(params, ep)
        assert np.array_equal(soft, es)
        assert (params[:, 0] + params[:, 1] <= 1.0).all()
        # rejection bookkeeping (model.py:280-287) and p_soft_no moments (model.py:339-342)
        ctx.job_begin(cfg, 1000, 4000)
        st = ctx.job_param_stats()
        ctx.job_end()
        assert st[0] == rej[:4000].sum() and st[3] == 4000
        assert np.isclose(st[1], es[:4000].sum(), rtol=1e-13)
        assert np.isclose(st[2], (es[:4000] ** 2).sum(), rtol=1e-13)


def test_rejection_abort_semantics(vax):
    # model.py:283-287: more than 10*n_rep rejections -> (None, None)
    res, err, msg = vax.run_model([70, 90], [50, 60], [0, 0.1], [3, 4], [0.2, 0.24], [10, 10], 95, 50,
                                  50000, USA_DATES, seed=3)
    assert res is None and msg == "Agnosticts: "
    assert err == ro.ERR_BOUNDS
    p, s = vax.sample_param_combinations([0.7, 0.9], [0.5, 0.6], [0, 0.1], [3, 4], [0.002, 0.0024],
                                         [0.1, 0.1], 5, seed=3)
    assert p is None and s is None
    # feasible but with rejections: accepted tuples all satisfy the constraint
    p, s = vax.sample_param_combinations([0.5, 0.9], [0.3, 0.6], [0, 0.1], [3, 4], [0.002, 0.0024],
                                         [0.1, 0.1], 40, seed=2)
    assert len(p) == 40 and all(a + b <= 1.0 for a, b, *_ in p)
    assert np.allclose(s, [1 - (a + b) for a, b, *_ in p], rtol=0, atol=0)


@pytest.mark.parametrize("lam", [0.0, 1e-3, 0.5, 1.999, 2.0, 2.5, 3.0, 4.0, 9.999, 10.0, 10.5, 16.0, 63.0,
                                 300.0, 2857.0, 1e4, 1e5])
def test_poisson_sampler_law(ctx, lam):
    """Chi-square goodness of fit of the device sampler against the exact Poisson pmf."""
    n = 2_000_000
    x = ctx.test_poisson(lam, n, seed=2024 + int(lam * 1000) % 9973).astype(np.int64)
    if lam == 0.0:
        assert (x == 0).all()
        return
    assert abs(x.mean() - lam) < 5 * np.sqrt(lam / n) + 1e-6 * lam
    assert abs(x.var() - lam) < 6 * lam * np.sqrt(2.0 / n) + 6 * np.sqrt(lam / n) + 2e-6 * lam * lam
    lo = max(int(stats.poisson.ppf(1e-5, lam)), 0)
    hi = int(stats.poisson.ppf(1 - 1e-5, lam)) + 1
    edges = np.arange(lo, hi + 1)
    # merge to at most ~200 cells with expected count >= 50
    cdf = stats.poisson.cdf(edges - 1, lam)
    step = max(1, len(edges) // 200)
    edges = edges[::step]
    cdf = stats.poisson.cdf(edges - 1, lam)
    probs = np.diff(np.concatenate([[0.0], cdf[1:], [1.0]]))
    obs = np.histogram(x, bins=np.concatenate([[-0.5], edges[1:] - 0.5, [np.inf]]))[0]
    keep = probs * n >= 50
    probs_k = np.concatenate([probs[keep], [probs[~keep].sum()]])
    obs_k = np.concatenate([obs[keep], [obs[~keep].sum()]])
    if probs_k[-1] * n < 5:
        probs_k[-2] += probs_k[-1]; obs_k[-2] += obs_k[-1]
        probs_k, obs_k = probs_k[:-1], obs_k[:-1]
    chi2 = ((obs_k - n * probs_k) ** 2 / (n * probs_k)).sum()
    p = stats.chi2.sf(chi2, len(probs_k) - 1)
    assert p > 1e-4, (lam, chi2, len(probs_k), p)


def test_normal_pair_law(ctx):
    z = ctx.test_normal(4_000_000, seed=5).astype(np.float64)
    assert abs(z.mean()) < 5 / np.sqrt(len(z))
    assert abs(z.var() - 1) < 5 * np.sqrt(2 / len(z))
    assert stats.kstest(z[:1_000_000], "norm").pvalue > 1e-3
    # the two normals of one Box-Muller pair are uncorrelated
    assert abs(np.corrcoef(z[0::2], z[1::2])[0, 1]) < 5 / np.sqrt(len(z) / 2)
    assert np.abs(z).max() < 7.0


def test_radix_select_exact(ctx):
    """Order statistics + sums of the pilot/direct matrix against np.sort."""
    rng = np.random.default_rng(1)
    for n in (70001, 16384, 200000):
        data = np.stack([
            rng.integers(0, 350000, n), rng.integers(0, 5, n), np.zeros(n, dtype=np.int64),
            np.full(n, 2**27 - 1), rng.poisson(3000, n), np.arange(n)[::-1],
            (rng.random(n) < 0.3) * rng.integers(0, 1 << 26, n), rng.integers(0, 2, n) * 50000,
            np.where(rng.random(n) < 0.97, 0, rng.integers(1, 1000, n)),       # 97 % ties at 0
            np.sort(rng.integers(0, 1 << 20, n)), rng.poisson(0.02, n), rng.integers(0, 1 << 27, n),
        ]).astype(np.int32)
        ranks = sorted([0, 1, n // 40, n // 40 + 1, n // 2, n - n // 40, n - 2, n - 1])
        got, sums = ctx.test_select(data, ranks)
        assert np.array_equal(got, np.sort(data, axis=1)[:, ranks]), n
        assert np.array_equal(sums, data.astype(np.int64).sum(axis=1))
    d = rng.integers(0, 1000, (3, 5000)).astype(np.int32)
    got, _ = ctx.test_select(d, [4999, 0, 2500])                                # unsorted rank list
    assert np.array_equal(got, np.sort(d, axis=1)[:, [4999, 0, 2500]])
    for m in (1, 2, 31, 33, 513, 1024, 1025):                                    # tiny rows
        d = rng.integers(0, 1000, (3, m)).astype(np.int32)
        rk = sorted({0, m // 2, m - 1})
        got, sums = ctx.test_select(d, rk)
        assert np.array_equal(got, np.sort(d, axis=1)[:, rk]), m
    # the three kernels (histogram refinement with the row staged in shared memory up to ~47k columns
    # or streamed from global memory above; 8-bit radix passes as the cross-check) on the same data:
    # wide ranges (three histogram levels), negatives, int32 extremes, duplicate ranks, odd row
    # lengths (unaligned rows take the scalar load path), rows longer than two load steps
    for n in (16384, 16383, 47000, 47104, 1000, 5000, 200000, 200003):
        data = np.stack([
            rng.integers(-(1 << 31), (1 << 31) - 1, n), rng.integers(0, 1 << 27, n), rng.integers(-5, 5, n),
            np.full(n, -(1 << 31)), np.full(n, (1 << 31) - 1), rng.poisson(40, n), rng.poisson(3e5, n),
            np.where(rng.random(n) < 0.5, -(1 << 31), (1 << 31) - 1), rng.integers(0, 4097, n),
            (rng.integers(0, 3, n) * (1 << 22)) + rng.integers(0, 2, n),
        ]).astype(np.int32)
        ranks = [0, 0, n // 40, n // 40 + 1, n // 2, n // 2, n - 2, n - 1]
        want = np.sort(data, axis=1)[:, ranks]
        try:
            outs = []
            for g in (0, 1, 2):
                ctx.set_option("select_global", g)
                got, sums = ctx.test_select(data, ranks)
                assert np.array_equal(got, want), (n, g)
                assert np.array_equal(sums, data.astype(np.int64).sum(axis=1)), (n, g)
        finally:
            ctx.set_option("select_global", 0)


# ------------------------------------------------------ (1) expected-value parity
def test_expected_single_realizations_vs_golden_and_oracle(vax, ctx, golden_dir):
    g = np.load(os.path.join(golden_dir, "single_realizations.npz"))
    plist = g["params"]
    worst = 0.0
    for T, N in ((564, 50000), (100, 1000), (1826, 1000000)):
        cfg = _cfg(vax, np.zeros(12), len(plist), T=T, N=N)
        got = ctx.expected_series(cfg, plist)            # [4][n][T]
        for i in range(len(plist)):
            want = g[f"exp/{i}/{T}/{N}"]                  # reference itself, [4][T]
            orc = ro.run_single_realization(None, 1, plist[i], T, N)
            for s, k in enumerate(SERIES):
                assert np.array_equal(orc[k], want[s])
                err = np.abs(got[s, i] - want[s]) / np.maximum(np.abs(want[s]), 1e-300)
                err[want[s] == 0] = np.abs(got[s, i])[want[s] == 0]
                worst = max(worst, err.max())
    assert worst <= 1e-5, worst           # north_star tolerance
    assert worst <= 1e-12, worst          # what fp64 state actually delivers


@pytest.mark.parametrize("name", ["usa_expected_seed12345_ci95_n100.npz",
                                  "wide5y_expected_seed7_ci90_n32_N1e6.npz"])
def test_expected_run_sampling_vs_reference_fixture(vax, golden_dir, name):
    """run_sampling over the reference's own parameter tuples, expected-value mode: mean and
    both quantile bands against the reference's output (tolerance 1e-5 relative)."""
    g = np.load(os.path.join(golden_dir, name))
    res = vax.run_sampling(g["params"], str(g["start_date"]), str(g["end_date"]), float(g["CI"]) / 100,
                           int(g["N"]), mode="expected")
    assert res["number_finished_samples"] == int(g["number_finished_samples"])
    for k in SERIES:
        for f in ("mean", "lower", "upper"):
            want = g[f"{k}/{f}"]
            got = res[k][f]
            assert got.shape == want.shape
            assert np.allclose(got, want, rtol=1e-5, atol=1e-9), (k, f)
            assert np.allclose(got, want, rtol=1e-11, atol=1e-12), (k, f)
        assert [str(x) for x in res[k]["dates"]] == list(g[f"{k}/dates"])


# ------------------------------------------- (2)/(3) stochastic mode, same parameters
def _oracle_rows(daily, weekly, T, W):
    """oracle bulk_counts -> the engine's raw-row layout [R][n]."""
    n_vacc, received, stock = daily
    return np.concatenate([n_vacc, weekly, received[0::7][:W], stock], axis=0)


def test_deterministic_rows_bit_exact_and_invariants(vax, ctx):
    bounds = usa_bounds_frac()
    n = 4096
    cfg = _cfg(vax, bounds, n, seed=77)
    params, _ = ctx.sample_params(cfg, 0, n)
    out, counts = ctx.run_params(cfg, params, want_counts=True)
    T, W = USA_T, USA_W
    n_vacc, dose, recv_w, stock = counts[:T], counts[T:T + W], counts[T + W:T + 2 * W], counts[T + 2 * W:]
    # `received` is a deterministic function of the tuple (model.py:69, 85-92): bit-exact
    d, w = ro.bulk_counts(5, params, T, 50000, threads=8)
    assert np.array_equal(recv_w, d[1][0::7][:W])
    recv_d = np.repeat(recv_w, 7, axis=0)[:T]
    assert np.array_equal(n_vacc + stock, recv_d)                 # dose conservation
    assert (np.diff(n_vacc, axis=0) >= 0).all()
    n_pro = (params[:, 0] * 50000).astype(np.int64)
    n_agn = ((1 - (params[:, 0] + params[:, 1])) * 50000).astype(np.int64)
    assert (n_vacc[-1] <= n_pro + n_agn).all()
    dv = np.diff(n_vacc, axis=0, prepend=0)
    wk = dv[0::7][:W].copy()
    wk[5:] += dv[5::7][:W - 5]                                    # dv[7k] + dv[7k-30]
    assert np.array_equal(dose, wk)
    # bounds-mode dump of the same global ids is the same integers
    assert np.array_equal(ctx.dump_counts(cfg, 0, n), counts)
    assert np.array_equal(ctx.dump_counts(cfg, 1000, 500), counts[:, 1000:1500])
    # the latency variant of the stepper (random numbers one day pair ahead) and the throughput
    # variant are the same arithmetic in a different order
    try:
        ctx.set_option("pipe_pilot", 0)
        assert np.array_equal(ctx.dump_counts(cfg, 0, n), counts)
        cfg_odd = _cfg(vax, bounds, 300, seed=78, T=37)             # odd T: a half pair at the end
        a_ = ctx.dump_counts(cfg_odd, 5, 300)
    finally:
        ctx.set_option("pipe_pilot", 1)
    assert np.array_equal(ctx.dump_counts(cfg_odd, 5, 300), a_)


def test_direct_reduction_equals_numpy_on_the_same_integers(vax, ctx):
    """Order statistics, lerp and mean: the engine vs np.quantile/mean (model.py:220-223) on
    the very trajectories it simulated (bit-exact bands, mean to 1e-13)."""
    for bounds, T, N, q, n in ((usa_bounds_frac(), USA_T, 50000, (0.025, 0.975), 1000),
                               (usa_bounds_frac(), USA_T, 50000, (0.49525, 0.50475), 100),
                               (wide_bounds_frac(), 1826, 1000000, (0.005, 0.995), 257),
                               (usa_bounds_frac(), 30, 333, (0.0, 1.0), 7),
                               (usa_bounds_frac(), 8, 50000, (0.25, 0.75), 1)):
        W = (T + 6) // 7
        cfg = _cfg(vax, bounds, n, seed=11, T=T, N=N, q=q)
        params, _ = ctx.sample_params(cfg, 0, n)
        out, counts = ctx.run_params(cfg, params, want_counts=True)
        c = counts.astype(np.int64)
        vals = [(c[:T] / N) * 100, c[T:T + W] * 1e6 / N,
                np.repeat(c[T + W:T + 2 * W], 7, axis=0)[:T] * 100 / N, c[T + 2 * W:] * 100 / N]
        for s in range(4):
            qq = np.quantile(vals[s], [q[0], q[1]], axis=1)
            assert np.array_equal(out.lower[s], qq[0]), (T, N, s)
            assert np.array_equal(out.upper[s], qq[1]), (T, N, s)
            assert np.allclose(out.mean[s], vals[s].mean(axis=1), rtol=1e-13, atol=1e-13)
        # run_bounds on the same seed draws the same tuples itself
        ob = ctx.run_bounds(cfg)
        for s in range(4):
            assert np.array_equal(ob.lower[s], out.lower[s]) and np.array_equal(ob.upper[s], out.upper[s])
            assert np.array_equal(ob.mean[s], out.mean[s])


def test_ks_per_week_against_oracle_on_identical_parameters(vax, ctx):
    """north_star check (2): identical sampled parameter sets, matched sample counts, two-sample
    KS per weekly date and series between the CUDA path and the numpy-law oracle."""
    bounds = usa_bounds_frac()
    n = int(os.environ.get("VX_KS_SAMPLES", "1000000"))   # BASELINE configs[1] as written: 1e6 samples per arm
    cfg = _cfg(vax, bounds, n, seed=4242)
    params, _ = ctx.sample_params(cfg, 0, n)
    counts = ctx.dump_counts(cfg, 0, n)
    d, w = ro.bulk_counts(987, params, USA_T, 50000, threads=os.cpu_count())
    orc = _oracle_rows(d, w, USA_T, USA_W)
    T, W = USA_T, USA_W
    rows = {"n_vacc": np.arange(0, T, 7), "dose": T + np.arange(W), "stock": T + 2 * W + np.arange(0, T, 7)}
    pvals = []
    for name, rr in rows.items():
        for r in rr:
            a, b = counts[r], orc[r]
            if np.array_equal(np.unique(a), np.unique(b)) and len(np.unique(a)) == 1:
                continue
            pvals.append(stats.ks_2samp(a, b, method="asymp").pvalue)
    pvals = np.array(pvals)
    assert len(pvals) > 200
    # north_star words the bar as "p > 0.01 per week".  Read literally over ~240 simultaneous tests
    # that bar fails by chance alone in 9 runs out of 10, so the assertion is its family-wise
    # form: no week rejects after Bonferroni, and the share of weeks below 0.01 stays at chance
    # level.  (Both arms share the parameter tuples, so they are positively correlated and
    # ks_2samp is conservative here; the sharp check is the pure-sampler-noise KS below.)
    assert pvals.min() > 0.01 / len(pvals), pvals.min()           # Bonferroni: no week rejects
    assert (pvals < 0.01).mean() <= 0.03, (pvals < 0.01).mean()   # chance level of p<0.01
    assert stats.kstest(pvals, "uniform").pvalue > 1e-4 or np.median(pvals) > 0.3
    print(f"\nKS per week: n={n} per arm, {len(pvals)} (series, week) tests, min p={pvals.min():.4f}, "
          f"median p={np.median(pvals):.3f}, share p<0.01: {(pvals < 0.01).mean():.4f}")
    # the per-sample difference has zero mean within Monte Carlo error (same tuples -> paired)
    for r in (100, 300, 563, T + 20, T + 60, T + 2 * W + 300):
        diff = counts[r].astype(np.float64) - orc[r]
        assert abs(diff.mean()) < 5 * diff.std() / np.sqrt(n) + 1e-9, r


@pytest.mark.parametrize("tuple_", [
    (0.4, 0.2, 0.01, 6.0, 0.011, 0.07),        # SURVEY 8c tuple
    (0.36, 0.24, 0.0005, 5.2, 0.0101, 0.052),  # tiny pressure: lam2 in the inversion regime
    (0.30, 0.055, 0.1, 4.0, 0.009, 0.125),     # wide-sweep corner, strong pressure
])
def test_ks_per_week_pure_sampler_noise(vax, ctx, tuple_):
    """Sharper than the mixed-parameter KS: EVERY sample has the same tuple, so the spread across
    samples is nothing but the Poisson draws.  CUDA sampler (inversion < 4, CF expansion above)
    vs numpy's algorithms (multiplication method < 10, PTRS above) in the oracle."""
    n = 200_000
    params = np.tile(np.array(tuple_, dtype=np.float64), (n, 1))
    cfg = _cfg(vax, np.zeros(12), n, seed=99)
    out, counts = ctx.run_params(vax._native.make_config(np.zeros(12), 50000, USA_T, n, 99, 0.025, 0.975,
                                                         direct_max=1 << 20), params, want_counts=True)
    d, w = ro.bulk_counts(4321, params, USA_T, 50000, threads=os.cpu_count())
    orc = _oracle_rows(d, w, USA_T, USA_W)
    T, W = USA_T, USA_W
    assert np.array_equal(counts[T + W:T + 2 * W], orc[T + W:T + 2 * W])      # deliveries: bit-exact
    rows = np.concatenate([np.arange(0, T, 7), T + np.arange(W), T + 2 * W + np.arange(0, T, 7)])
    pvals = []
    for r in rows:
        a, b = counts[r], orc[r]
        if a.min() == a.max() and b.min() == b.max():
            assert a[0] == b[0]
            continue
        pvals.append(stats.ks_2samp(a, b, method="asymp").pvalue)
        # first two moments agree within Monte Carlo error
        sd = max(a.std(), b.std(), 1e-9)
        assert abs(a.mean() - b.mean()) < 6 * sd * np.sqrt(2 / n) + 1e-9, r
        m4 = max(((a - a.mean()) ** 4).mean(), ((b - b.mean()) ** 4).mean())
        se_sd = np.sqrt(max(m4 - sd ** 4, 0.0) / (4 * n * sd * sd)) * np.sqrt(2)   # delta method
        assert abs(a.std() - b.std()) < 6 * se_sd + 1e-9, r
    pvals = np.array(pvals)
    assert len(pvals) > 100      # rows that are constant in both arms carry no test
    assert pvals.min() > 0.01 / len(pvals), pvals.min()
    assert (pvals < 0.01).mean() <= 0.04, (pvals < 0.01).mean()
    print(f"\npure-noise KS {tuple_}: {len(pvals)} tests, min p={pvals.min():.4f}, median p={np.median(pvals):.3f}, "
          f"share p<0.01: {(pvals < 0.01).mean():.4f}")


@pytest.mark.parametrize("tuple_,N", [
    ((0.4, 0.2, 0.0006, 6.0, 0.011, 0.07), 50000),     # conversion rate lam2 sits in [2, 10] for most of the run
    ((0.4, 0.2, 0.01, 6.0, 0.011, 0.07), 1000000),     # the Dash app's maximum population (app.py:157-166)
])
def test_ks_pure_sampler_noise_full_size(vax, ctx, tuple_, N):
    """The sensitive form of north_star check (2) at BASELINE configs[1]'s size: 1e6 samples per arm,
    ONE tuple for every sample, so the spread across samples is nothing but the Poisson draws and
    the two arms are independent (what ks_2samp assumes).  First tuple: lam2 = n_agn*f*pressure stays
    in [2, 10] -- where the device sampler's normal-quantile expansion is least accurate (sup-CDF
    error 1.5e-4 at lam in [2,3]) -- and the late-campaign lam1 ~ 1.7 * (yesterday's conversions) as
    well.  Second tuple: population 1e6, lam up to ~1e5."""
    n = int(os.environ.get("VX_KS_SAMPLES", "1000000"))
    T, W = USA_T, USA_W
    params = np.tile(np.array(tuple_, dtype=np.float64), (n, 1))
    cfg = vax._native.make_config(np.zeros(12), N, T, n, 2026, 0.025, 0.975, direct_max=n + 1)
    _, counts = ctx.run_params(cfg, params, want_counts=True)
    rows = np.concatenate([np.arange(0, T, 7), T + np.arange(W), T + 2 * W + np.arange(0, T, 7)])
    counts = counts[np.concatenate([rows, T + W + np.arange(W)])]          # keep host memory bounded
    d, w = ro.bulk_counts(20261020, params, T, N, threads=os.cpu_count())
    orc = _oracle_rows(d, w, T, W)[np.concatenate([rows, T + W + np.arange(W)])]
    del d, w
    assert np.array_equal(counts[len(rows):], orc[len(rows):])            # deliveries: deterministic, bit-exact
    lam2_mid = None
    if N == 50000:
        n_agn0 = int((1 - (tuple_[0] + tuple_[1])) * N)
        f = counts[:81].astype(np.float64).mean(axis=1) / N                   # mean vaccinated fraction per week
        lam2 = n_agn0 * f * tuple_[2]                                         # upper bound (n_agn only shrinks)
        lam2_mid = float(np.median(lam2))
        assert ((lam2 > 1.5) & (lam2 < 10)).mean() > 0.6, lam2               # the regime this test is about
    pvals, zmean = [], []
    for i in range(len(rows)):
        a, b = counts[i], orc[i]
        if a.min() == a.max() and b.min() == b.max():
            assert a[0] == b[0]
            continue
        pvals.append(stats.ks_2samp(a, b, method="asymp").pvalue)
        sd = max(a.std(), b.std(), 1e-9)
        zmean.append((a.mean() - b.mean()) / (sd * np.sqrt(2 / n)))
    pvals, zmean = np.array(pvals), np.array(zmean)
    assert len(pvals) > 150
    assert pvals.min() > 0.01 / len(pvals), pvals.min()                   # no week rejects (family-wise 1 %)
    assert (pvals < 0.01).mean() <= 0.04, (pvals < 0.01).mean()           # share below 0.01 at chance level
    assert np.abs(zmean).max() < 5.0, np.abs(zmean).max()                 # per-week means: no bias at 1e-3 sd
    print(f"\npure-noise KS at n={n} per arm, tuple {tuple_}, N={N}: {len(pvals)} tests, min p={pvals.min():.4f}, "
          f"median p={np.median(pvals):.3f}, share p<0.01: {(pvals < 0.01).mean():.4f}, max |z| of means "
          f"{np.abs(zmean).max():.2f}, median lam2 {lam2_mid}")


def test_bands_agree_with_reference_fixture_within_mc_error(vax, golden_dir):
    """The reference's own seeded run (n=100, CI=95): our bands at n=20000 must be compatible."""
    g = np.load(os.path.join(golden_dir, "usa_seed12345_ci95_n100.npz"))
    res, err, msg = vax.run_model([35, 45], [12, 25], [0.0, 0.02], [5, 7], [1, 1.2], [5, 10], 95, 20000,
                                  50000, USA_DATES, seed=12345)
    assert err == "" and msg == "Agnosticts: 37 - 46%"
    assert msg == str(g["msg_agnostics_pct"])
    for k in SERIES:
        spread = np.maximum(res[k]["upper"] - res[k]["lower"], 1e-9)
        # mean of 100 samples scatters by ~ spread/4/sqrt(100); allow 5 sigma
        assert (np.abs(res[k]["mean"] - g[f"{k}/mean"]) < 5 * spread / 4 / 10 + 1e-9).mean() > 0.97, k
        assert [str(x) for x in res[k]["dates"]] == list(g[f"{k}/dates"])
        assert res[k]["mean"].shape == g[f"{k}/mean"].shape


# --------------------------------------- (3) streamed == direct, sharding invariance
def _run(vax, ctx, bounds, n, seed, direct_max, T=USA_T, N=50000, q=(0.025, 0.975), pilot=0):
    cfg = _cfg(vax, bounds, n, seed=seed, T=T, N=N, q=q, direct_max=direct_max, pilot_samples=pilot)
    return ctx.run_bounds(cfg)


def _same(a, b):
    for s in range(4):
        if not (np.array_equal(a.lower[s], b.lower[s]) and np.array_equal(a.upper[s], b.upper[s])
                and np.array_equal(a.mean[s], b.mean[s])):
            return False
    return a.c.n_finished == b.c.n_finished


@pytest.mark.parametrize("q", [(0.025, 0.975), (0.49525, 0.50475), (0.0, 1.0), (0.001, 0.9999)])
def test_streamed_equals_direct(vax, ctx, q):
    bounds = usa_bounds_frac()
    n = 50000
    direct = _run(vax, ctx, bounds, n, 31, direct_max=1 << 20, q=q)
    assert direct.c.n_passes == 0
    streamed = _run(vax, ctx, bounds, n, 31, direct_max=1, q=q, pilot=8192)
    assert streamed.c.n_passes >= 1
    assert _same(direct, streamed)


@pytest.mark.parametrize("case", [
    dict(T=1, n=4097, N=50000, bounds="usa"),                      # a single day
    dict(T=13, n=5000, N=50000, bounds="usa"),                     # less than one 14-day tile
    dict(T=15, n=33, N=50000, bounds="usa"),                       # fewer samples than the pilot
    dict(T=100, n=40001, N=1000, bounds="usa"),                    # ragged warp, tiny population
    dict(T=371, n=20000, N=1000000, bounds="wide"),                # lam up to ~1e5, deliveries ~1e6
    dict(T=200, n=30000, N=50000, bounds="reject"),                # rejection-heavy parameter draw
    dict(T=564, n=65537, N=50000, bounds="usa", q=(0.1, 0.9)),     # just above the default direct_max
])
def test_streamed_equals_direct_edge_shapes(vax, ctx, case):
    b = {"usa": usa_bounds_frac(), "wide": wide_bounds_frac(),
         "reject": np.array([0.5, 0.9, 0.3, 0.6, 0.0, 0.1, 3, 4, 0.002, 0.0024, 0.1, 0.1])}[case["bounds"]]
    q = case.get("q", (0.025, 0.975))
    kw = dict(T=case["T"], N=case["N"], q=q)
    direct = _run(vax, ctx, b, case["n"], 17, direct_max=1 << 22, **kw)
    streamed = _run(vax, ctx, b, case["n"], 17, direct_max=1, pilot=2048, **kw)
    assert direct.c.n_passes == 0 and streamed.c.n_passes >= 1
    assert _same(direct, streamed)
    assert streamed.c.n_rejected == direct.c.n_rejected
    if case["bounds"] == "reject":
        assert direct.c.n_rejected > 0
    default = _run(vax, ctx, b, case["n"], 17, direct_max=0, **kw)      # library's own choice of path
    assert _same(direct, default)


def test_explicit_tuples_streamed_equals_direct(vax, ctx):
    """run_sampling's tuple-driven form (vx_run_params) through the pilot + streaming fold."""
    nat = vax._native
    bounds = usa_bounds_frac()
    n = 30011
    params, _ = ctx.sample_params(_cfg(vax, bounds, n, seed=8), 0, n)
    a = ctx.run_params(nat.make_config(np.zeros(12), 50000, 150, n, 8, 0.025, 0.975, direct_max=1 << 20), params)
    b = ctx.run_params(nat.make_config(np.zeros(12), 50000, 150, n, 8, 0.025, 0.975, direct_max=1,
                                       pilot_samples=2048), params)
    assert a.c.n_passes == 0 and b.c.n_passes >= 1 and _same(a, b)
    # and the bounds-driven job on the same seed draws the very same tuples
    c = ctx.run_bounds(_cfg(vax, bounds, n, seed=8, T=150, direct_max=1, pilot_samples=2048))
    assert _same(a, c)
    with pytest.raises(nat.VaxmcError):                       # assert p_pro + p_anti <= 1 (model.py:63)
        bad = params.copy(); bad[5, 0] = 0.9; bad[5, 1] = 0.2
        ctx.run_params(nat.make_config(np.zeros(12), 50000, 150, n, 8, 0.025, 0.975), bad)


def test_streamed_refinement_paths(vax, ctx):
    """Tiny windows / no pilot / degenerate bounds force the VX_REFINE loop; still exact."""
    bounds = usa_bounds_frac()
    n = 20000
    direct = _run(vax, ctx, bounds, n, 5, direct_max=1 << 20)
    try:
        ctx.set_option("max_bins", 16)
        a = _run(vax, ctx, bounds, n, 5, direct_max=1, pilot=4096)
        assert a.c.n_passes >= 2 and _same(direct, a)
        ctx.set_option("use_pilot", 0)
        b = _run(vax, ctx, bounds, n, 5, direct_max=1)
        assert b.c.n_passes >= 3 and _same(direct, b)
        ctx.set_option("max_bins", 262144)
        c = _run(vax, ctx, bounds, n, 5, direct_max=1)
        assert c.c.n_passes >= 2 and _same(direct, c)
        ctx.set_option("use_pilot", 1)
        ctx.set_option("pilot_sigmas", 0.0)     # windows too narrow: ranks fall outside
        d = _run(vax, ctx, bounds, n, 5, direct_max=1, pilot=2048)
        assert d.c.n_passes >= 2 and _same(direct, d)
    finally:
        ctx.set_option("max_bins", 262144); ctx.set_option("use_pilot", 1); ctx.set_option("pilot_sigmas", 6.0)
    try:                                      # regime sort on/off: same integers either way
        ctx.set_option("sort_samples", 0)
        e = _run(vax, ctx, bounds, n, 5, direct_max=1, pilot=4096)
        assert _same(direct, e)
    finally:
        ctx.set_option("sort_samples", 1)
    try:                                      # samples stepped beside the pilot: none / some / the default
        for side in (0, 1024, 3000):
            ctx.set_option("side_samples", side)
            f = _run(vax, ctx, bounds, n, 5, direct_max=1, pilot=4096)
            assert _same(direct, f) and f.c.side_samples == (side & ~127), side
        ctx.set_option("max_bins", 16)        # ... and through the VX_REFINE passes
        g = _run(vax, ctx, bounds, n, 5, direct_max=1, pilot=4096)
        assert g.c.n_passes >= 2 and _same(direct, g)
    finally:
        ctx.set_option("side_samples", -1); ctx.set_option("max_bins", 262144)
    dflt = _run(vax, ctx, bounds, 200_000, 6, direct_max=1)
    assert dflt.c.side_samples == 0 and dflt.c.n_pilot == 16384     # default: nothing stepped beside the pilot
    assert dflt.c.fold_samples == 200_000 - dflt.c.n_pilot - dflt.c.side_samples
    assert _same(dflt, _run(vax, ctx, bounds, 200_000, 6, direct_max=1 << 20))
    # single-value bounds ([v, v], model.py:376-377): every sample has the same tuple
    deg = np.array([0.4, 0.4, 0.2, 0.2, 0.01, 0.01, 6, 6, 0.011, 0.011, 0.07, 0.07])
    d0 = _run(vax, ctx, deg, 30000, 9, direct_max=1 << 20)
    d1 = _run(vax, ctx, deg, 30000, 9, direct_max=1, pilot=4096)
    assert _same(d0, d1)
    assert np.array_equal(d0.lower[2], d0.upper[2])       # deliveries are deterministic here


def _sharded_equals_single(vax, ctx, nat, dist, torch, bounds, n, pilot, worlds):
    cfg = _cfg(vax, bounds, n, seed=2025, direct_max=1, pilot_samples=pilot)
    single = ctx.run_bounds(cfg)
    for world, stats_in_slab in worlds:
        ctxs = [nat.Context(0) for _ in range(world)]
        engines = [dist.ContextEngine(c) for c in ctxs]
        for r, e in enumerate(engines):
            e.begin(cfg, *dist.shard_range(n, r, world))
        sts = [e.param_stats() for e in engines]
        assert sum(sts)[3] == n
        for e in engines:
            if not stats_in_slab:                      # caller-side reduction of the statistics
                assert not e.ctx.job_set_param_stats(sum(sts))
            e.pilot()
        outs = [nat.ResultBuffers(cfg.T) for _ in engines]
        for _ in range(8):
            for e in engines:
                e.accumulate()
            accs = [e.acc_tensor() for e in engines]
            total = torch.stack(accs).sum(dim=0)       # what the NCCL all-reduce(sum) does
            for a in accs:
                a.copy_(total)
            torch.cuda.synchronize()
            rcs = [e.finalize(o) for e, o in zip(engines, outs)]
            assert len(set(rcs)) == 1
            if rcs[0] != nat.VX_REFINE:
                break
        for o in outs:
            assert _same(single, o)
            assert o.c.n_rejected == single.c.n_rejected and not o.c.aborted
            assert o.c.soft_no_mean == pytest.approx(single.c.soft_no_mean, rel=1e-8)
            assert o.c.soft_no_std == pytest.approx(single.c.soft_no_std, rel=1e-6)
        for e in engines:
            e.end()
        for c in ctxs:
            c.close()


def test_sharding_is_bit_exact(vax, ctx):
    """1 vs 2 vs 3 vs 8 shards of the same global sample range, merged by integer addition
    (what the NCCL all-reduce does across GPUs): identical bytes out."""
    import torch

    nat = vax._native
    from importlib import import_module
    dist = import_module("covid19-vaccination-model_b200.distributed")
    bounds = usa_bounds_frac()
    # (n, pilot, worlds): the second and third shape put shard boundaries INSIDE the pilot matrix, whose
    # share of a shard is folded from the matrix (full 8192-column chunks with 16-byte loads, partial
    # chunks, and an odd leading dimension that takes the scalar path)
    for n, pilot, worlds in ((40000, 4096, ((2, False), (3, True), (8, True))),
                             (40000, 24576, ((2, True), (3, True))),
                             (40001, 20001, ((3, True),))):
        _sharded_equals_single(vax, ctx, nat, dist, torch, bounds, n, pilot, worlds)
    n = 40000
    # infeasible bounds discovered through the slab header: every rank reports `aborted`
    bad = _cfg(vax, np.array([0.7, 0.9, 0.5, 0.6, 0, 0.1, 3, 4, 0.002, 0.0024, 0.1, 0.1]), n, seed=3,
               direct_max=1, pilot_samples=4096)
    ctxs = [nat.Context(0) for _ in range(2)]
    engines = [dist.ContextEngine(c) for c in ctxs]
    outs = [nat.ResultBuffers(bad.T) for _ in engines]
    for r, e in enumerate(engines):
        e.begin(bad, *dist.shard_range(n, r, 2))
        e.param_stats()
        e.pilot()
        e.accumulate()
    accs = [e.acc_tensor() for e in engines]
    total = torch.stack(accs).sum(dim=0)
    for a in accs:
        a.copy_(total)
    torch.cuda.synchronize()
    for e, o in zip(engines, outs):
        assert e.finalize(o) == nat.VX_OK and o.c.aborted == 1 and o.c.n_rejected > 10 * n
        e.end()
    for c in ctxs:
        c.close()


def test_max_running_time_partial_results(vax, ctx):
    """model.py:187-189, 363-366: a time budget yields partial results plus the error string."""
    ctx.set_option("chunk_samples", 4096)
    try:
        res, err, msg = vax.run_model([35, 45], [12, 25], [0.0, 0.02], [5, 7], [1, 1.2], [5, 10], 95,
                                      3_000_000, 50000, USA_DATES, max_running_time=1e-4, seed=1,
                                      direct_max=1, pilot_samples=4096)
    finally:
        ctx.set_option("chunk_samples", 1 << 20)
    k = res["number_finished_samples"]
    assert 0 < k < 3_000_000
    assert err == (f"ERROR: Maximum computation time of 0.0001s exceeded. Only {k} of the desired "
                   f"3000000 Monte Carlo runs were performed.")
    assert np.isfinite(res[SERIES[0]]["mean"]).all()
    # a budget of 0 is a real limit in the reference (one sample, then stop): here one chunk
    ctx.set_option("chunk_samples", 4096)
    try:
        res0, err0, _ = vax.run_model([35, 45], [12, 25], [0.0, 0.02], [5, 7], [1, 1.2], [5, 10], 95,
                                      3_000_000, 50000, USA_DATES, max_running_time=0, seed=1,
                                      direct_max=1, pilot_samples=4096)
    finally:
        ctx.set_option("chunk_samples", 1 << 20)
    assert res0["number_finished_samples"] == 4096 and "Maximum computation time of 0s exceeded" in err0
    # a direct job is one chunk: it always completes, and then there is no error (model.py:364-366)
    res1, err1, _ = vax.run_model([35, 45], [12, 25], [0.0, 0.02], [5, 7], [1, 1.2], [5, 10], 95,
                                  1000, 50000, USA_DATES, max_running_time=0, seed=1)
    assert res1["number_finished_samples"] == 1000 and err1 == ""


# ------------------------------------------------------- full BASELINE sizes: properties
def test_usa_1e6_properties(vax, ctx):
    """configs[1]: USA scenario at 1e6 samples.  Size-independent properties + the streamed
    result equals the direct (materialise + exact select) result at full size."""
    n = 1_000_000
    res, err, msg = vax.run_model([35, 45], [12, 25], [0.0, 0.02], [5, 7], [1, 1.2], [5, 10], 95, n,
                                  50000, USA_DATES, seed=12345)
    assert err == "" and msg == "Agnosticts: 37 - 46%"
    assert res["number_finished_samples"] == n
    v, d, r, s = (res[k] for k in SERIES)
    assert len(v["mean"]) == 564 and len(d["mean"]) == 81
    assert np.allclose(v["mean"] + s["mean"], r["mean"], rtol=1e-12)   # doses are conserved
    assert (np.diff(v["mean"]) >= 0).all() and (np.diff(r["mean"]) >= 0).all()
    for x in (v, d, r, s):
        assert (x["lower"] <= x["upper"]).all()
        assert (x["lower"] <= x["mean"] + 1e-9).mean() > 0.95
    assert 60 < v["mean"][-1] < 85 and v["upper"][-1] <= 88.0 + 1e-9   # pro+agnostics cap
    direct, _, _ = vax.run_model([35, 45], [12, 25], [0.0, 0.02], [5, 7], [1, 1.2], [5, 10], 95, n,
                                 50000, USA_DATES, seed=12345, direct_max=2_000_000)
    for k in SERIES:
        for f in ("mean", "lower", "upper"):
            assert np.array_equal(res[k][f], direct[k][f]), (k, f)


def test_wide_5y_streamed_properties(vax, ctx):
    """configs[3] shape (wide bounds, 5-year range) at a size the box finishes in seconds."""
    b = dict(start_date="2020-12-15", end_date="2025-12-14")
    res, err, msg = vax.run_model([30, 50], [5.5, 31.5], [0.0, 0.1], [4, 8], [0.9, 1.3], [2.5, 12.5], 99,
                                  400_000, 50000, b, seed=7)
    assert err == "" and res["number_finished_samples"] == 400_000
    v, d, r, s = (res[k] for k in SERIES)
    assert len(v["mean"]) == 1826 and len(d["mean"]) == 261
    assert np.allclose(v["mean"] + s["mean"], r["mean"], rtol=1e-12)
    assert (np.diff(v["mean"]) >= 0).all()
    assert (v["upper"] <= 94.5 + 1e-9).all()       # 1 - min(anti) = 94.5 %


def test_run_model_grid_equals_sequential_run_model(vax):
    """configs[4] entry point: the boxes pipelined by vx_run_grid (several waves, a ragged last
    wave, an infeasible and an almost infeasible box in the middle) == run_model one by one."""
    import bench

    boxes = bench.grid_boxes(10, seed=5)
    args = [(b["pro"], b["anti"], b["pressure"], b["tau"], b["nv0"], b["nvmax"]) for b in boxes]
    args.insert(4, ([70, 90], [50, 60], [0, 0.1], [3, 4], [0.2, 0.24], [10, 10]))      # infeasible box
    args.insert(8, ([60, 99], [38, 99], [0, 0.1], [3, 4], [0.2, 0.24], [10, 10]))      # aborts on the rejection count
    seeds = [100 + i for i in range(len(args))]
    for n_rep, wave in ((200_000, 3), (200_000, 0), (2_000, 5)):                         # streamed jobs, direct jobs
        grid = vax.run_model_grid(args, 99, n_rep, 50000, USA_DATES, seeds=seeds, wave=wave)
        assert len(grid) == len(args)
        n_none = 0
        for a, s_, g in zip(args, seeds, grid):
            res, err, msg = vax.run_model(*a, 99, n_rep, 50000, USA_DATES, seed=s_)
            assert (err, msg) == (g[1], g[2])
            if res is None:
                assert g[0] is None and err == ro.ERR_BOUNDS
                n_none += 1
                continue
            assert g[0]["number_finished_samples"] == n_rep
            for k in SERIES:
                for f in ("mean", "lower", "upper"):
                    assert np.array_equal(res[k][f], g[0][k][f]), (k, f)
        assert n_none == 2
    assert vax.run_model_grid([], 99, 1000, 50000, USA_DATES) == []


def test_integration_md_stub_runs_and_matches(vax):
    """The ctypes stub printed in INTEGRATION.md (what a reference maintainer would paste into
    model.py) is executed verbatim and must reproduce run_model's arrays."""
    import re

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    md = open(os.path.join(root, "INTEGRATION.md")).read()
    code = [b for b in re.findall(r"```python\n(.*?)```", md, flags=re.S) if "run_sampling_bounds" in b][0]
    cwd = os.getcwd()
    os.chdir(root)
    try:
        ns = {}
        exec(compile(code, "INTEGRATION.md", "exec"), ns)
        data, res = ns["run_sampling_bounds"](usa_bounds_frac(), "2020-12-15", "2022-07-1", 0.95, 50000,
                                              20000, 4711)
    finally:
        os.chdir(cwd)
    ours, err, msg = vax.run_model([35, 45], [12, 25], [0.0, 0.02], [5, 7], [1, 1.2], [5, 10], 95, 20000,
                                   50000, USA_DATES, seed=4711)
    assert data["number_finished_samples"] == 20000 and err == ""
    for k in SERIES:
        for f in ("mean", "lower", "upper"):
            assert np.array_equal(data[k][f], ours[k][f]), (k, f)
        assert list(data[k]["dates"]) == list(ours[k]["dates"])


def test_run_model_quantiles_fan_is_exact(vax, ctx):
    """N3: any set of quantile fractions of ONE sample (jobs sharing the seed see the same
    trajectories): equal to np.quantile over the dumped integers, streamed and direct."""
    qs = [0.05, 0.25, 0.5, 0.75, 0.95, 0.99, 0.01]
    n, N, T, W = 30000, 50000, USA_T, USA_W
    counts = ctx.dump_counts(_cfg(vax, usa_bounds_frac(), n, seed=77), 0, n).astype(np.int64)
    vals = [(counts[:T] / N) * 100, counts[T:T + W] * 1e6 / N,
            np.repeat(counts[T + W:T + 2 * W], 7, axis=0)[:T] * 100 / N, counts[T + 2 * W:] * 100 / N]
    for dm in (1 << 20, 1):
        res, err, msg = vax.run_model_quantiles([35, 45], [12, 25], [0.0, 0.02], [5, 7], [1, 1.2], [5, 10],
                                                qs, n, N, USA_DATES, seed=77, direct_max=dm, pilot_samples=4096)
        assert err == "" and res["number_finished_samples"] == n
        for s, k in enumerate(SERIES):
            assert sorted(res[k]["quantiles"]) == sorted(qs)
            want = np.quantile(vals[s], sorted(qs), axis=1)
            for i, q in enumerate(sorted(qs)):
                assert np.array_equal(res[k]["quantiles"][q], want[i]), (k, q, dm)
            assert np.allclose(res[k]["mean"], vals[s].mean(axis=1), rtol=1e-13, atol=1e-13)


def _chi2_p(x, lam, n):
    lo = max(int(stats.poisson.ppf(1e-5, lam)), 0)
    hi = int(stats.poisson.ppf(1 - 1e-5, lam)) + 1
    edges = np.arange(lo, hi + 1)[::max(1, (hi + 1 - lo) // 200)]
    cdf = stats.poisson.cdf(edges - 1, lam)
    probs = np.diff(np.concatenate([[0.0], cdf[1:], [1.0]]))
    obs = np.histogram(x, bins=np.concatenate([[-0.5], edges[1:] - 0.5, [np.inf]]))[0]
    keep = probs * n >= 50
    pk = np.concatenate([probs[keep], [probs[~keep].sum()]])
    ok = np.concatenate([obs[keep], [obs[~keep].sum()]])
    if pk[-1] * n < 5:
        pk[-2] += pk[-1]; ok[-2] += ok[-1]; pk, ok = pk[:-1], ok[:-1]
    return stats.chi2.sf(((ok - n * pk) ** 2 / (n * pk)).sum(), len(pk) - 1)


def test_exact_poisson_mode(vax, ctx):
    """The validation sampler (inversion < 10, PTRS >= 10 with an fp64 acceptance test): exact law
    per lambda, and GPU-vs-GPU agreement of whole trajectories with the default sampler on one
    tuple (pure sampler noise) at 1e6 samples per arm."""
    n = 2_000_000
    ctx.set_option("exact_poisson", 1)
    try:
        for lam in (0.5, 9.999, 10.0, 12.0, 63.0, 2857.0, 1e5):
            x = ctx.test_poisson(lam, n, seed=77 + int(lam)).astype(np.int64)
            assert abs(x.mean() - lam) < 5 * np.sqrt(lam / n) + 1e-6 * lam
            assert _chi2_p(x, lam, n) > 1e-4, lam
        m = 1_000_000
        params = np.tile(np.array([0.4, 0.2, 0.01, 6.0, 0.011, 0.07]), (m, 1))
        cfgx = vax._native.make_config(np.zeros(12), 50000, 200, m, 5, 0.025, 0.975, direct_max=1 << 21)
        _, exact = ctx.run_params(cfgx, params, want_counts=True)
    finally:
        ctx.set_option("exact_poisson", 0)
    _, fast = ctx.run_params(cfgx, params, want_counts=True)
    assert not np.array_equal(exact, fast)                      # different samplers, different draws
    T, W = 200, 29
    assert np.array_equal(exact[T + W:T + 2 * W], fast[T + W:T + 2 * W])     # deliveries: deterministic
    pv = []
    for r in list(range(0, T, 7)) + list(range(T, T + W)) + list(range(T + 2 * W, 2 * T + 2 * W, 7)):
        a, b = exact[r], fast[r]
        if a.min() == a.max() and b.min() == b.max():
            continue
        pv.append(stats.ks_2samp(a, b, method="asymp").pvalue)
        sd = max(a.std(), b.std())
        assert abs(a.mean() - b.mean()) < 6 * sd * np.sqrt(2 / m) + 1e-9, r
    pv = np.array(pv)
    assert pv.min() > 0.01 / len(pv) and (pv < 0.01).mean() <= 0.04, (pv.min(), (pv < 0.01).mean())
    print(f"\nexact vs default sampler, {m} samples per arm: {len(pv)} KS tests, min p={pv.min():.4f}, "
          f"median p={np.median(pv):.3f}")


def test_cli_configs0_readme_scenario(vax, tmp_path):
    """BASELINE configs[0]: the README command line (default mc_samples=100, --CI default 0.95 with
    the /100 quirk of model.py:358) through main(); JSON == run_model's arrays; bands against the
    reference's own seeded run of the same command within Monte Carlo error of n=100."""
    import json

    out = tmp_path / "res.json"
    argv = ("--pro=35,45 --anti=12,25 --pressure=0.,0.02 --dupl_time=5,7 --init_stock=1,1.2 "
            "--max_delivery=5,10 --date_range=2020-12-15,2022-07-1 --seed 5 --no_plot").split()
    assert vax.main(argv + ["--json_out", str(out)]) == 0
    js = json.loads(out.read_text())
    res, err, msg = vax.run_model([35.0, 45.0], [12.0, 25.0], [0.0, 0.02], [5.0, 7.0], [1.0, 1.2], [5.0, 10.0],
                                  0.95, 100, 50000, dict(start_date="2020-12-15", end_date="2022-07-1"), seed=5)
    assert err == "" and js["number_finished_samples"] == 100
    for k in SERIES:
        assert len(js[k]["mean"]) == (81 if k == SERIES[1] else 564)
        for f in ("mean", "lower", "upper"):
            assert js[k][f] == res[k][f].tolist()
        assert js[k]["dates"][0].startswith("2020-12-15")
        # CI = 0.95 % -> quantiles 0.49525 / 0.50475: the band hugs the median
        assert (res[k]["upper"] >= res[k]["lower"]).all()
    width = res[SERIES[0]]["upper"][-1] - res[SERIES[0]]["lower"][-1]
    assert width < 2.0            # reference measured 0.149 at CI=0.95 vs 46.0 at CI=95


def test_multigpu_nccl_bit_exact_under_torchrun(vax):
    """north_star check (3) on real GPUs: tests/multigpu_bitexact.py under torch.distributed.run with
    min(2, device_count) ranks -- libvaxmc's own ncclAllReduce of the slab; every scenario (streamed,
    the CLI CI quirk, the 5-year sweep, a direct job, an entropy seed, an infeasible box) must be
    bit-identical to the single-GPU result and identical on all ranks."""
    import json
    import subprocess
    import sys

    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (the driver's scaling run prints the same digest for N = 1, 2, 4, 8)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    proc = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                           "--master-addr", "127.0.0.1", "--master-port", "29731",
                           os.path.join(root, "tests", "multigpu_bitexact.py")],
                          capture_output=True, text=True, timeout=900)
    assert proc.returncode == 0, proc.stdout[-2000:] + proc.stderr[-2000:]
    lines = [json.loads(x) for x in proc.stdout.splitlines() if x.startswith("{")]
    assert len(lines) >= 6 and all(x["ok"] for x in lines), lines
