"""oracle/fit_poisson_cf.py -- derivation of the Poisson sampler's normal-quantile expansion.

TEST/DESIGN INFRASTRUCTURE (not product code).  The device sampler for lam >= 2
(csrc/vaxmc_kernels.cuh: poisson_draw) returns floor(Q(z)), z ~ N(0,1), with

    Q = lam + s z + (1/3 + z^2/6) + (-z/36 - z^3/72)/s + (-8/405 + 7 z^2/810 + z^4/270)/lam
        + (a1 z + a3 z^3 + a5 z^5)/lam^1.5 + (b0 + b2 z^2 + b4 z^4 + b6 z^6)/lam^2

The first three orders are the known expansion of the continuous inverse of the Poisson CDF
(Giles, "Algorithm 955", ACM TOMS 42(1) 2016, Q_N1).  This script
  1. recovers the next two orders asymptotically (large lam) from the exact jump points
     z_k = Phi^-1(F(k; lam)) -> Q = k+1, as a check of the expansion's form, and
  2. fits a/b by weighted least squares on the jump points for lam in [2, 32], which is what
     the kernel uses (an asymptotic series is not optimal at lam ~ 2-4), and
  3. prints the sup-CDF error and the chi-square non-centrality per draw of floor(Q(z))
     against the exact Poisson law.

Run:  python oracle/fit_poisson_cf.py
"""
from __future__ import annotations

import numpy as np
from scipy import stats


def q3(z, lam):
    s = np.sqrt(lam)
    return (lam + s * z + (1 / 3 + z * z / 6) + (-z / 36 - z ** 3 / 72) / s
            + (-8 / 405 + 7 * z * z / 810 + z ** 4 / 270) / lam)


def q5(z, lam, a, b, zclamp=None):
    # the kernel does not clamp z in the two fitted orders (an earlier version did, at +-5: the
    # law's error is the same either way, see main())
    zc = z if zclamp is None else np.clip(z, -zclamp, zclamp)
    return (q3(z, lam) + (a[0] * zc + a[1] * zc ** 3 + a[2] * zc ** 5) / lam ** 1.5
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


def main():
    a, b = fit()
    print("fitted a1,a3,a5 =", ", ".join(f"{x:.8f}" for x in a))
    print("fitted b0,b2,b4,b6 =", ", ".join(f"{x:.8f}" for x in b))
    print(f"{'lam':>7} {'sup-CDF 3 orders':>18} {'sup-CDF 5 orders':>18} {'chi2/draw (5)':>15}")
    for lam in (2, 2.5, 3, 4, 5, 6, 8, 10, 16, 32, 100, 1000, 10000):
        e3, _ = law_error(lam, q3)
        e5, c5 = law_error(lam, lambda z, l: q5(z, l, a, b))
        e5c, _ = law_error(lam, lambda z, l: q5(z, l, a, b, zclamp=5.0))
        print(f"{lam:7g} {e3:18.2e} {e5:18.2e} {c5:15.1e}   (z clamped at +-5: {e5c:.2e})")


if __name__ == "__main__":
    main()
