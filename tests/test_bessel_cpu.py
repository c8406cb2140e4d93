"""General-order Matern core of the Gram kernels (lsqfitgp_b200/csrc/bessel_k.cuh: Temme series / Steed CF2 for K_nu),
built for the host by oracle/Makefile and checked against mpmath (40 digits) and against the oracle's restatement of the
reference formula with scipy.special.kv (src/lsqfitgp/_special/_bessel.py:70-99, _kernels/_matern.py:74-76).
Gram tolerance of BASELINE.json: 1e-13 relative."""

import ctypes
import pathlib

import numpy as np
import pytest

from oracle import iso as oiso

ROOT = pathlib.Path(__file__).resolve().parents[1]
NUS = [0.0, 0.1, 0.25, 0.49, 0.7, 1.0, 1.3, 2.0, 2.7, 3.49, 5.2, 10.3, 30.7]


@pytest.fixture(scope='module')
def matern():
    path = ROOT / 'oracle' / 'libfastmath_host.so'
    if not path.exists():
        import subprocess
        subprocess.run(['make', '-C', str(ROOT / 'oracle')], check=True)
    lib = ctypes.CDLL(str(path))
    lib.lgp_host_matern_nu.argtypes = [ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long]

    def call(nu, r2):
        r2 = np.ascontiguousarray(r2, dtype=np.float64)
        val, dr2 = np.empty_like(r2), np.empty_like(r2)
        rc = lib.lgp_host_matern_nu(nu, r2.ctypes.data, val.ctypes.data, dr2.ctypes.data, r2.size)
        return rc, val, dr2
    return call


def sample_r2(nu, rng, k):
    tn = 2 * (nu or 1)
    return np.concatenate([10.0 ** rng.uniform(-12, 2.3, k), [2.0 / tn, 4.0 / tn, 4.0000001 / tn, 1e-300]])


@pytest.mark.parametrize('nu', NUS)
def test_value_and_derivative_vs_mpmath(matern, nu):
    import mpmath as mp
    mp.mp.dps = 40
    rng = np.random.default_rng(int(nu * 100) + 1)
    r2 = sample_r2(nu, rng, 150)
    rc, val, dr2 = matern(nu, r2)
    assert rc == 0
    tn = 2 * (nu or 1)
    for a, v, d in zip(r2, val, dr2):
        z = tn * mp.mpf(float(a))
        x = mp.sqrt(z)
        if nu == 0:
            assert v == 0.0 and d == 0.0
            continue
        rv = 2 / mp.gamma(nu) * (x / 2) ** nu * mp.besselk(nu, x)
        rd = -mp.mpf(tn) / 4 * 2 / mp.gamma(nu) * (x / 2) ** (nu - 1) * mp.besselk(nu - 1, x)
        if abs(rv) > 1e-280:
            assert abs(v - rv) / abs(rv) < 5e-14, (nu, a, v, rv)
        if abs(rd) > 1e-280 and abs(rd) < 1e280 and a > 1e-200:
            assert abs(d - rd) / abs(rd) < 5e-14, (nu, a, d, rd)


@pytest.mark.parametrize('nu', NUS)
def test_value_vs_oracle_formula(matern, nu):
    """against the reference's own evaluation route (scipy kv); AMOS itself is ~1e-13 from the exact value for some orders"""
    rng = np.random.default_rng(int(nu * 100) + 2)
    r2 = sample_r2(nu, rng, 2000)[:-1]
    rc, val, dr2 = matern(nu, r2)
    ref = oiso.matern_core(r2, nu)
    if nu == 0:
        assert np.all(val == 0) and np.all(ref == 0)
        return
    ok = np.abs(ref) > 1e-280
    assert np.max(np.abs(val[ok] - ref[ok]) / np.abs(ref[ok])) < 2e-13
    with np.errstate(all='ignore'):
        dref = oiso.matern_dr2(r2, nu)
    ok = (np.abs(dref) > 1e-280) & np.isfinite(dref)
    assert np.max(np.abs(dr2[ok] - dref[ok]) / np.abs(dref[ok])) < 2e-13


def test_limits_and_rejections(matern):
    rc, val, dr2 = matern(1.7, np.array([0.0]))
    assert rc == 0 and val[0] == 1.0 and dr2[0] == 0.0
    rc, val, _ = matern(0.0, np.array([0.0, 1e-3]))
    assert list(val) == [1.0, 0.0]                      # nu = 0: white noise (_matern.py:74)
    assert matern(-0.1, np.array([1.0]))[0] == 1
    assert matern(100.5, np.array([1.0]))[0] == 1
    assert matern(float('nan'), np.array([1.0]))[0] == 1
    # half-integer orders agree with the closed form (tests/test_special.py:84-97 of the reference: 1e-14 on the core)
    r2 = 10.0 ** np.random.default_rng(5).uniform(-8, 2, 500)
    for p in range(4):
        _, val, _ = matern(p + 0.5, r2)
        ref = oiso.maternp_core(r2 - 1e-30 / (2 * p + 1), p)
        assert np.max(np.abs(val - ref) / ref) < 5e-14
