"""CPU: the mathematics behind the device RNG/sampler (no GPU needed).

* Philox4x32-10 restatement against the Random123 known-answer vectors, and the counter layout.
* The Poisson sampler's normal-quantile expansion (constants parsed out of the CUDA source)
  against the exact Poisson law: sup-CDF error bounds quoted in DESIGN.md.
"""
import os
import re

import numpy as np
import pytest

import philox_ref as pr
from fit_poisson_cf import law_error, moment_bias, q5

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUH = os.path.join(ROOT, "covid19-vaccination-model_b200", "csrc", "vaxmc_kernels.cuh")


def _const(name):
    src = open(CUH).read()
    m = re.search(rf"\b{name}\s*=\s*(-?[0-9.eE+-]+)f", src)
    assert m, name
    return float(m.group(1))


def test_philox_known_answers():
    for c, k, e in pr.KAT:
        r = pr.philox4x32_10(*[np.uint32(x) for x in c], k[0], k[1])
        assert [int(x) for x in r] == list(e)


def test_parameter_draw_restatement_properties():
    b = [0.5, 0.9, 0.3, 0.6, 0.0, 0.1, 3, 4, 0.002, 0.0024, 0.1, 0.1]
    p, s, rej = pr.draw_params(b, 99, 0, 4000)
    assert (p[:, 0] + p[:, 1] <= 1.0).all() and rej.sum() > 0
    assert ((p[:, 0] >= 0.5) & (p[:, 0] <= 0.9) & (p[:, 1] >= 0.3) & (p[:, 1] <= 0.6)).all()
    assert (p[:, 5] == 0.1).all()                         # single-value bound -> constant
    # a sample's tuple depends only on (seed, global id): sharding cannot change it
    p2, _, _ = pr.draw_params(b, 99, 1000, 500)
    assert np.array_equal(p2, p[1000:1500])
    p3, _, _ = pr.draw_params(b, 100, 0, 100)
    assert not np.array_equal(p3, p[:100])


def test_poisson_expansion_error_bounds():
    a = [_const("CF_A1"), _const("CF_A3"), _const("CF_A5")]
    b = [_const("CF_B0"), _const("CF_B2"), _const("CF_B4"), _const("CF_B6")]
    thr = _const("POISSON_INV_MAX")
    assert thr == 2.0
    bounds = {2: 1.6e-4, 2.5: 1.7e-4, 3: 1.3e-4, 3.5: 8e-5, 4: 6e-5, 5: 6e-5, 6: 5e-5, 8: 3e-5, 10: 2e-5, 16: 1e-5,
              32: 5e-6, 100: 5e-6, 2857: 5e-6}  # >=32: grid resolution of the check
    for lam, bound in bounds.items():
        sup, chi = law_error(lam, lambda z, l: q5(z, l, a, b), M=1_000_001)
        assert sup <= bound, (lam, sup)
        assert chi <= 8e-7 or lam > 1000, (lam, chi)        # needs > 2e7 draws at ONE lam to show at 3 sigma


def test_poisson_expansion_moment_bias():
    """What a 1e9-sample job can see of the approximate sampler: the bias of its first two moments
    (DESIGN.md 3.1 quotes these; profiles/r02_sampler_bias_1e9.jsonl is the measurement)."""
    a = [_const("CF_A1"), _const("CF_A3"), _const("CF_A5")]
    b = [_const("CF_B0"), _const("CF_B2"), _const("CF_B4"), _const("CF_B6")]
    for lam in (2, 2.5, 3, 4, 6, 8, 10, 16, 32, 100):
        dm, dv = moment_bias(lam, lambda z, l: q5(z, l, a, b))
        assert abs(dm) <= 1.25e-4 and abs(dm) / np.sqrt(lam) <= 8e-5, (lam, dm)
        assert abs(dv) / lam <= 3.2e-4, (lam, dv)
        if lam >= 32:
            assert abs(dm) <= 1.2e-5, (lam, dm)


def test_inversion_regime_tail_is_negligible():
    # the sequential search is capped at 40 steps; for lam < 2 that tail is < 1e-30
    from scipy import stats

    assert stats.poisson.sf(40, 2.0) < 1e-30
