"""Accuracy of the short in-kernel exp/sqrt of the Gram kernels (lsqfitgp_b200/csrc/fastmath.cuh), measured on the
host against glibc: the header is plain C++ and oracle/fastmath_host.c builds it with g++ (oracle/Makefile).  The
Gram parity bar is 1e-13 relative (BASELINE.json north_star); the functions must stay below 1 ulp = 2.2e-16."""

import ctypes
import pathlib

import numpy as np
import pytest

ROOT = pathlib.Path(__file__).resolve().parents[1]


@pytest.fixture(scope='module')
def host():
    path = ROOT / 'oracle' / 'libfastmath_host.so'
    if not path.exists():
        import subprocess
        subprocess.run(['make', '-C', str(ROOT / 'oracle')], check=True)
    lib = ctypes.CDLL(str(path))
    lib.lgp_host_div_recip.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_long]
    lib.lgp_host_div_recip.restype = None
    call_div = lib.lgp_host_div_recip
    for name in ('lgp_host_exp_neg', 'lgp_host_exp_neg_fast', 'lgp_host_sqrt'):
        getattr(lib, name).argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long]
        getattr(lib, name).restype = None

    def call(name, a):
        if name == 'lgp_host_div_recip':
            a, b = (np.ascontiguousarray(v, dtype=np.float64) for v in a)
            out = np.empty_like(a)
            call_div(a.ctypes.data, b.ctypes.data, out.ctypes.data, a.size)
            return out
        a = np.ascontiguousarray(a, dtype=np.float64)
        out = np.empty_like(a)
        getattr(lib, name)(a.ctypes.data, out.ctypes.data, a.size)
        return out
    return call


def ulps(got, ref):
    return np.abs(got - ref) / np.spacing(np.abs(ref))


def test_exp_neg_full_range(host):
    rng = np.random.default_rng(11)
    a = -np.concatenate([rng.uniform(0, 40, 400000), rng.uniform(0, 708, 400000), 10.0 ** rng.uniform(-320, 2.8, 200000),
                         [0.0, 5e-324, 1e-300, 0.5, 1.0, 707.9, 708.0]])
    got = host('lgp_host_exp_neg', a)
    ref = np.exp(a)
    assert np.max(ulps(got, ref)) <= 1.0  # glibc itself is within 1 ulp, not correctly rounded
    assert got[-7] == 1.0


def test_exp_neg_underflow_and_nan(host):
    a = np.array([-708.5, -720.0, -744.0, -745.0, -745.13, -745.2, -746.0, -1e3, -np.inf, np.nan])
    got = host('lgp_host_exp_neg', a)
    ref = np.exp(a)
    sub = 5e-324
    assert np.all(np.abs(got[:-1] - ref[:-1]) <= sub)  # subnormal results: within one subnormal step
    assert np.all(got[6:9] == 0.0)
    assert np.isnan(got[-1])


def test_exp_neg_fast_in_range(host):
    rng = np.random.default_rng(12)
    a = -np.concatenate([rng.uniform(0, 50, 500000), rng.uniform(0, 708, 500000), [0.0, 708.0]])
    got = host('lgp_host_exp_neg_fast', a)
    assert np.max(ulps(got, np.exp(a))) <= 1.0
    # same bits as the guarded version wherever both apply
    assert np.array_equal(got, host('lgp_host_exp_neg', a))


def test_sqrt_with_emulated_estimate(host):
    rng = np.random.default_rng(13)
    z = np.concatenate([rng.uniform(0, 100, 500000), 10.0 ** rng.uniform(-289, 290, 500000), [1.0, 4.0, 2.0, 1e-289, 1e290]])
    got = host('lgp_host_sqrt', z)
    ref = np.sqrt(z)
    assert np.max(ulps(got, ref)) <= 1.0   # faithfully rounded
    assert np.mean(got != ref) < 0.2       # and correctly rounded most of the time
    assert got[-5] == 1.0 and got[-4] == 2.0


def test_true_error_against_mpmath(host):
    """error against the exact value (mpmath, 40 digits) on a sample: below 1 ulp for both functions"""
    import mpmath
    mpmath.mp.dps = 40
    rng = np.random.default_rng(14)
    a = -rng.uniform(0, 700, 1500)
    got = host('lgp_host_exp_neg_fast', a)
    err = [abs(mpmath.mpf(float(g)) - mpmath.exp(mpmath.mpf(float(x)))) / mpmath.mpf(float(np.spacing(g))) for g, x in zip(got, a)]
    assert max(err) < 1.0
    z = 10.0 ** rng.uniform(-200, 200, 1500)
    got = host('lgp_host_sqrt', z)
    err = [abs(mpmath.mpf(float(g)) - mpmath.sqrt(mpmath.mpf(float(x)))) / mpmath.mpf(float(np.spacing(g))) for g, x in zip(got, z)]
    assert max(err) < 1.0


def test_division_by_reciprocal_is_correctly_rounded(host):
    """ (x - loc) / scale of the Gram kernels: reciprocal + two exact-residual corrections == IEEE division, bit for bit
    (random and adversarial divisors: significands 1, 1 + ulp, 2 - ulp, ...), with the library division outside the
    proven range (tiny / huge / non-finite numerators) """
    rng = np.random.default_rng(15)
    N = 3_000_000
    a = rng.standard_normal(N) * 10.0 ** rng.uniform(-30, 30, N)
    b = (1 + rng.random(N)) * 2.0 ** rng.integers(-60, 60, N) * rng.choice([-1, 1], N)
    assert np.array_equal(host('lgp_host_div_recip', (a, b)), a / b)
    for sig in [1.0, np.nextafter(1.0, 2), np.nextafter(2.0, 1), 1.5, 1 + 2.0 ** -26, 2 - 2.0 ** -26, 4 / 3, 5 / 3, 1.9999999]:
        a = (1 + rng.random(300000)) * 2.0 ** rng.integers(-5, 5, 300000)
        a[:1000] = np.nextafter(np.linspace(1, 2, 1000), 3)
        bb = np.full_like(a, sig * 8)
        assert np.array_equal(host('lgp_host_div_recip', (a, bb)), a / bb)
    a = np.array([0.0, 6.0, 1e-310, 1e305, np.inf, -np.inf, np.nan, 2.0 ** -600, 2.0 ** 600])
    out = host('lgp_host_div_recip', (a, np.full_like(a, 3.0)))
    ref = a / 3.0
    assert all((o == r) or (o != o and r != r) for o, r in zip(out, ref))


def test_log_and_rational_quadratic_core():
    """ short log of the rational-quadratic Gram path: absolute error ~1 ulp of the result (what exp(c log x) needs);
    the core (1 + r2/beta)^(-beta/2) within 2e-14 of numpy's pow where the kernel uses it (|exponent| <= 200) """
    import ctypes as ct
    lib = ct.CDLL(str(ROOT / 'oracle' / 'libfastmath_host.so'))
    lib.lgp_host_log_ge1.argtypes = [ct.c_void_p, ct.c_void_p, ct.c_long]
    lib.lgp_host_ratquad.argtypes = [ct.c_double, ct.c_void_p, ct.c_void_p, ct.c_long]
    rng = np.random.default_rng(16)
    x = np.concatenate([1 + 10.0 ** rng.uniform(-17, 0, 200000), 10.0 ** rng.uniform(0, 150, 200000),
                        [1.0, 2.0, 1.0078125, np.nextafter(2.0, 1), 2.0 ** 500]])
    out = np.empty_like(x)
    lib.lgp_host_log_ge1(x.ctypes.data, out.ctypes.data, x.size)
    ref = np.log(x)
    assert np.all(np.abs(out - ref) <= np.maximum(1.5 * np.spacing(ref), 4e-18))   # absolute: ~1 ulp, 4e-18 near x = 1
    assert abs(out[-5]) < 1e-17      # log(1)
    for beta in [0.3, 1.0, 3.0, 10.0, 100.0]:
        r2max = min(beta * np.expm1(400.0 / beta), 2.0 ** 500) if 400.0 / beta < 700 else 2.0 ** 500
        r2 = np.concatenate([10.0 ** rng.uniform(-12, min(np.log10(r2max), 12), 200000), [0.0, r2max]])
        o = np.empty_like(r2)
        lib.lgp_host_ratquad(beta, r2.ctypes.data, o.ctypes.data, r2.size)
        refv = (1 + r2 / beta) ** (-beta / 2)
        assert np.max(np.abs(o - refv) / refv) < 6e-14, beta
        ok = refv > 1e-20
        assert np.max(np.abs(o[ok] - refv[ok]) / refv[ok]) < 2e-14, beta
