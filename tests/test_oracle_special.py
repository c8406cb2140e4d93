"""Pin the Matern cores of the oracle (reference tests/test_special.py:68-97 and
tests/kernels/test_kernels.py:500-512 restated): kvmodx2 against mpmath, kvmodx2_hi against kvmodx2."""
import mpmath
import numpy as np
import pytest

from oracle import iso


@pytest.mark.parametrize('nu', [0.5, 1.5, 2.5, 0.3, 1.7, 3.2])
def test_kvmodx2_vs_mpmath(nu):
    x2 = np.array([1e-8, 1e-3, 0.1, 1.0, 4.0, 30.0, 200.0])
    got = iso.kvmodx2(nu, x2)
    mpmath.mp.dps = 40
    ref = np.array([float(2 / mpmath.gamma(nu) * (mpmath.sqrt(v) / 2) ** nu * mpmath.besselk(nu, mpmath.sqrt(v)))
                    for v in x2])
    np.testing.assert_allclose(got, ref, rtol=1e-13, atol=1e-15)
    assert iso.kvmodx2(nu, np.array(0.0)) == 1


@pytest.mark.parametrize('p', [0, 1, 2, 3, 5])
def test_kvmodx2_hi_vs_kvmodx2(p):
    x2 = np.logspace(-6, 3, 200)
    np.testing.assert_allclose(iso.kvmodx2_hi(x2, p), iso.kvmodx2(p + 0.5, x2), rtol=1e-13, atol=1e-300)


@pytest.mark.parametrize('p', [0, 1, 2, 3])
def test_matern_half_integer_equals_maternp(p, rng):
    r2 = rng.uniform(0, 40, 1000)
    r2[:3] = 0
    np.testing.assert_allclose(iso.matern_core(r2, p + 0.5), iso.maternp_core(r2, p), rtol=1e-9)


@pytest.mark.parametrize('p', [1, 2, 3])
def test_dr2_finite_difference(p):
    r2 = np.array([0.01, 0.5, 2.0, 9.0])
    h = 1e-6
    for core, d in [(lambda r: iso.maternp_core(r, p), lambda r: iso.maternp_dr2(r, p)),
                    (lambda r: iso.matern_core(r, p + 0.5), lambda r: iso.matern_dr2(r, p + 0.5)),
                    (iso.expquad_core, iso.expquad_dr2), (iso.cauchy_core, iso.cauchy_dr2)]:
        fd = (core(r2 + h) - core(r2 - h)) / (2 * h)
        np.testing.assert_allclose(d(r2), fd, rtol=1e-6)
