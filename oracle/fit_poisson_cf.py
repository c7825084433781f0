"""oracle/fit_poisson_cf.py -- derivation of the Poisson sampler's normal-quantile expansion.

TEST/DESIGN INFRASTRUCTURE (not product code).  The device sampler for lam >= 2
(csrc/vaxmc_kernels.cuh: poisson_draw) returns floor(Q(z)), z ~ N(0,1), with

    Q = lam + s z + (1/3 + z^2/6) + (-z/36 - z^3/72)/s + (-8/405 + 7 z^2/810 + z^4/270)/lam
        + (k3 + a1 z + a3 z^3 + a5 z^5)/lam^1.5 + (b0 + b2 z^2 + b4 z^4 + b6 z^6)/lam^2

The first three orders are the known expansion of the continuous inverse of the Poisson CDF
(Giles, "Algorithm 955", ACM TOMS 42(1) 2016, Q_N1).  This script
  1. recovers the next two orders asymptotically (large lam) from the exact jump points
     z_k = Phi^-1(F(k; lam)) -> Q = k+1, as a check of the expansion's form, and
  2. fits a/b by weighted least squares on the jump points for lam in [2, 32] (an asymptotic
     series is not optimal at lam ~ 2-4),
     which is what the kernel uses (KERNEL_A / KERNEL_B below),
  3. prints the sup-CDF error, the chi-square non-centrality per draw and the bias of the first
     two moments of floor(Q(z)) against the exact Poisson law, and
  4. (--moments) shows the alternative that was NOT adopted: a Gauss-Newton refinement of a/b and
     a constant k3 with the mean and variance of floor(Q(z)) as extra residuals.  A trajectory is a
     sum of ~1e3 draws, so what a 1e9-sample job can see of the sampler is its moment bias (up to
     1.2e-4 counts per draw at lam 2-10, visible at 1e9 samples per arm against the exact PTRS
     mode: profiles/r02_sampler_bias_1e9.jsonl).  The refit lowers that bias ~5x, but only by
     inflating the far tail at lam ~ 2 (pmf errors of +20 % at k = 10 to +130 % at k = 12, chi-square
     non-centrality 4e-7 -> 2e-5 per draw), which a goodness-of-fit test at one lam sees with 1e6
     draws.  Adding the lam^-5/2 and lam^-3 orders to the fit (tried) lowers the sup-CDF error to
     5e-5 but only halves the mean bias: its pattern in lam (+1.1e-4 at 2.5, -1e-4 at 6) is a
     lattice effect of the floor, not a missing power of lam^-1/2.

Run:  python oracle/fit_poisson_cf.py
"""
from __future__ import annotations

import numpy as np
from scipy import stats


def q3(z, lam):
    s = np.sqrt(lam)
    return (lam + s * z + (1 / 3 + z * z / 6) + (-z / 36 - z ** 3 / 72) / s
            + (-8 / 405 + 7 * z * z / 810 + z ** 4 / 270) / lam)


def q5(z, lam, a, b, zclamp=None, k3=0.0):
    # the kernel does not clamp z in the two fitted orders (an earlier version did, at +-5: the
    # law's error is the same either way, see main())
    zc = z if zclamp is None else np.clip(z, -zclamp, zclamp)
    return (q3(z, lam) + (k3 + a[0] * zc + a[1] * zc ** 3 + a[2] * zc ** 5) / lam ** 1.5
            + (b[0] + b[1] * zc ** 2 + b[2] * zc ** 4 + b[3] * zc ** 6) / lam ** 2)


def jump_points(lam, zmax=5.5):
    ks = np.arange(0, int(lam + 7 * np.sqrt(lam)) + 8)
    F = stats.poisson.cdf(ks, lam)
    S = stats.poisson.sf(ks, lam)
    z = np.where(F < 0.5, stats.norm.ppf(F), -stats.norm.ppf(S))
    m = np.isfinite(z) & (np.abs(z) < zmax)
    return z[m], (ks + 1)[m].astype(float)


def fit(lams=(2, 2.25, 2.5, 2.75, 3, 3.5, 4, 4.5, 5, 5.5, 6, 7, 8, 9, 10, 12, 14, 16, 20, 25, 32)):
    Z, R, L, Wt = [], [], [], []
    for lam in lams:
        z, q = jump_points(lam)
        Z.append(z); R.append(q - q3(z, lam)); L.append(np.full(len(z), float(lam)))
        Wt.append(stats.norm.pdf(z) ** 0.5 / lam ** 0.25)
    Z, R, L, Wt = map(np.concatenate, (Z, R, L, Wt))
    A = np.stack([Z / L ** 1.5, Z ** 3 / L ** 1.5, Z ** 5 / L ** 1.5,
                  1 / L ** 2, Z ** 2 / L ** 2, Z ** 4 / L ** 2, Z ** 6 / L ** 2], 1)
    c = np.linalg.lstsq(A * Wt[:, None], R * Wt, rcond=None)[0]
    return c[:3], c[3:]


def law_error(lam, qfun, zlim=8.0, M=4_000_001):
    """(sup |F_approx - F_exact|, chi-square non-centrality per draw) of X = max(floor(Q(z)), 0)."""
    zs = np.linspace(-zlim, zlim, M)
    edges = np.concatenate([[-np.inf], (zs[1:] + zs[:-1]) / 2, [np.inf]])
    w = np.diff(stats.norm.cdf(edges))
    K = int(lam + 12 * np.sqrt(lam)) + 30
    X = np.minimum(np.maximum(np.floor(qfun(zs, lam)), 0).astype(np.int64), K)
    pm = np.bincount(X, weights=w, minlength=K + 1)
    p = stats.poisson.pmf(np.arange(K + 1), lam)
    m = p > 1e-12
    return (float(np.max(np.abs(np.cumsum(pm) - stats.poisson.cdf(np.arange(K + 1), lam)))),
            float(np.sum((pm[m] - p[m]) ** 2 / p[m])))


class _Grid:
    """Midpoint quadrature of a function of z ~ N(0,1) on a fine uniform grid."""

    def __init__(self, M=1_000_001, zlim=8.5):
        self.z = np.linspace(-zlim, zlim, M)
        edges = np.concatenate([[-np.inf], (self.z[1:] + self.z[:-1]) / 2, [np.inf]])
        self.w = np.diff(stats.norm.cdf(edges))


_GRID = None


def moment_bias(lam, qfun, grid=None):
    """(E[X] - lam, Var[X] - lam) of X = max(floor(Q(z)), 0)."""
    global _GRID
    if grid is None:
        if _GRID is None:
            _GRID = _Grid()
        grid = _GRID
    X = np.maximum(np.floor(qfun(grid.z, lam)), 0)
    m = float((grid.w * X).sum())
    v = float((grid.w * (X - m) ** 2).sum())
    return m - lam, v - lam


MOMENT_LAMS = (2, 2.1, 2.25, 2.5, 2.75, 3, 3.25, 3.5, 4, 4.5, 5, 5.5, 6, 7, 8, 9, 10, 12, 14, 16, 20, 25, 32, 45, 64)


def refit_with_moments(a, b, weight=1000.0, iters=5,
                       lams=(2, 2.25, 2.5, 2.75, 3, 3.5, 4, 4.5, 5, 5.5, 6, 7, 8, 9, 10, 12, 14, 16, 20, 25, 32)):
    """Gauss-Newton on p = (a1, a3, a5, b0, b2, b4, b6, k3): the weighted jump-point residuals of
    fit() plus, per lam in MOMENT_LAMS, weight * (mean bias / sd, variance bias / (2 lam))."""
    jp = [(lam,) + jump_points(lam) for lam in lams]

    def resid(p):
        r = [(q5(z, lam, p[0:3], p[3:7], k3=p[7]) - q) * (stats.norm.pdf(z) ** 0.5 / lam ** 0.25) for lam, z, q in jp]
        mm = []
        for lam in MOMENT_LAMS:
            dm, dv = moment_bias(lam, lambda z, l: q5(z, l, p[0:3], p[3:7], k3=p[7]))
            mm += [weight * dm / np.sqrt(lam), weight * 0.5 * dv / lam]
        return np.concatenate(r + [np.array(mm)])

    p = np.concatenate([a, b, [0.0]])
    h = np.array([2e-4, 1e-4, 5e-5, 2e-4, 2e-4, 1e-4, 2e-5, 2e-4])
    for _ in range(iters):
        r = resid(p)
        J = np.stack([(resid(p + h[k] * np.eye(8)[k]) - r) / h[k] for k in range(8)], 1)
        p = p + 0.7 * np.linalg.lstsq(J, -r, rcond=1e-8)[0]
    return p[0:3], p[3:7], float(p[7])


# the constants in csrc/vaxmc_kernels.cuh (CF_A*, CF_B*): fit() to 8 digits
KERNEL_A = (0.01865529, 0.00004626, -0.00184102)
KERNEL_B = (0.00100592, -0.01903893, 0.00285482, 0.00029421)
KERNEL_K3 = 0.0


def main():
    import sys

    a, b = fit()
    print("fitted a1,a3,a5 =", ", ".join(f"{x:.8f}" for x in a))
    print("fitted b0,b2,b4,b6 =", ", ".join(f"{x:.8f}" for x in b))
    alt = None
    if "--moments" in sys.argv:
        alt = refit_with_moments(a, b)
        print("NOT adopted, with moments: a =", ", ".join(f"{x:.8f}" for x in alt[0]), " b =",
              ", ".join(f"{x:.8f}" for x in alt[1]), f" k3 = {alt[2]:.8f}")
    print(f"{'lam':>7} {'sup-CDF 3 ord.':>15} {'5 orders':>10} {'chi2/draw':>10} {'E X - lam':>11} {'Var X - lam':>12}"
          + ("   | with moments: sup-CDF, chi2/draw, E X - lam, Var X - lam" if alt else ""))
    for lam in (2, 2.5, 3, 3.5, 4, 5, 6, 8, 10, 16, 32, 100, 1000, 10000):
        e3, _ = law_error(lam, q3)
        qk = lambda z, l: q5(z, l, a, b)
        ek, ck = law_error(lam, qk)
        mk, vk = moment_bias(lam, qk)
        line = f"{lam:7g} {e3:15.2e} {ek:10.2e} {ck:10.1e} {mk:+11.2e} {vk:+12.2e}"
        if alt:
            qa = lambda z, l: q5(z, l, alt[0], alt[1], k3=alt[2])
            ea, ca = law_error(lam, qa)
            ma, va = moment_bias(lam, qa)
            line += f"   | {ea:.2e} {ca:.1e} {ma:+.2e} {va:+.2e}"
        print(line)


if __name__ == "__main__":
    main()
