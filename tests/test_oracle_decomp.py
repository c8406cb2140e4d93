"""Pin the Chol oracle with the reference's own test battery (tests/linalg/test_decomp.py:42-261 of the
reference, restated): every method against scipy.linalg.solve on K + eps I, eigvalsh, and finite differences."""
import numpy as np
import pytest
from scipy import linalg, stats

from oracle.decomp import Chol


def mat(n, s, rng):
    # reference tests/linalg/test_decomp.py:42-81: random orthogonal x eigenvalues 1 + eps + cos(s + i)
    O = stats.ortho_group.rvs(n, random_state=rng) if n > 1 else np.atleast_2d(1.0)
    eigvals = 1 + 1e-3 + np.cos(s + np.arange(n))
    K = (O * eigvals) @ O.T
    return (K + K.T) / 2


@pytest.mark.parametrize('n', [1, 2, 10, 64])
def test_solves(n, rng):
    K = mat(n, 0.3, rng)
    dec = Chol(K)
    Kreg = K + dec.eps * np.eye(n)
    A = rng.standard_normal((n, 3))
    r = rng.standard_normal(n)
    sol = lambda B: linalg.solve(Kreg, B, assume_a='pos')
    np.testing.assert_allclose(dec.ginv_linear(A), sol(A), rtol=1e-9)
    np.testing.assert_allclose(dec.pinv_bilinear(A, r), A.T @ sol(r), rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(dec.ginv_quad(A), A.T @ sol(A), rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(dec.ginv_diagquad(A), np.diag(A.T @ sol(A)), rtol=1e-9)
    np.testing.assert_allclose(dec.ginv(), np.linalg.inv(Kreg), rtol=1e-8, atol=1e-10)
    Z = dec.correlate(np.eye(n))
    np.testing.assert_allclose(Z @ Z.T, Kreg, rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(dec.back_correlate(A), Z.T @ A, rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(dec.pinv_correlate(r), np.linalg.solve(Z, r), rtol=1e-9)
    assert dec.n == n and dec.m == n
    assert dec.matrix() is not None


@pytest.mark.parametrize('n', [1, 2, 10])
def test_value_and_derivatives(n, rng):
    k = 2
    A0 = mat(n, 0.1, rng)
    As = [mat(n, 1.0 + i, rng) * 0.1 for i in range(k)]
    r0 = rng.standard_normal(n)
    rs = rng.standard_normal((n, k))
    Kfun = lambda p: A0 + sum(pi * Ai for pi, Ai in zip(p, As))
    rfun = lambda p: r0 + rs @ p
    p0 = np.array([0.3, -0.2])

    def direct(p):
        K = Kfun(p)
        dec = Chol(K)
        Kreg = K + dec.eps * np.eye(n)
        w = linalg.eigvalsh(Kreg)
        r = rfun(p)
        return 1 / 2 * (n * np.log(2 * np.pi) + np.sum(np.log(w)) + r @ linalg.solve(Kreg, r, assume_a='pos'))

    dec = Chol(Kfun(p0))
    dK = np.stack(As, axis=2)
    dr = rs
    vj = lambda G: np.einsum('ij,ijk->k', G, dK)
    rj = lambda g: g @ dr
    vec = rng.standard_normal(k)
    val, grev, gfwd, fish, fvec = dec.minus_log_normal_density(
        rfun(p0), dK_vjp=vj, dr_vjp=rj, dK=dK, dr=dr, dK_jvp_vec=dK @ vec, dr_jvp_vec=dr @ vec,
        value=True, gradrev=True, gradfwd=True, fisher=True, fishvec=True)
    np.testing.assert_allclose(val, direct(p0), atol=1e-9)
    h = 1e-6
    fd = np.array([(direct(p0 + h * e) - direct(p0 - h * e)) / (2 * h) for e in np.eye(k)])
    np.testing.assert_allclose(grev, fd, rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(gfwd, grev, rtol=1e-10, atol=1e-12)
    # fisher = 1/2 tr(K^-1 dK_i K^-1 dK_j) + dr_i' K^-1 dr_j (reference test :241-261)
    Kreg = Kfun(p0) + dec.eps * np.eye(n)
    Ki = np.linalg.inv(Kreg)
    F = np.array([[0.5 * np.trace(Ki @ As[i] @ Ki @ As[j]) + rs[:, i] @ Ki @ rs[:, j] for j in range(k)]
                  for i in range(k)])
    np.testing.assert_allclose(fish, F, rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(fvec, F @ vec, rtol=1e-8, atol=1e-10)


def test_scaling_and_eps():
    # powers-of-two equilibration and Gershgorin jitter (reference _decomp.py:349-361,384-393)
    K = np.array([[4.0, 1.0], [1.0, 0.25]]) * 1e6
    dec = Chol(K + np.diag([0, 1e6]))
    s = np.array([2.0 ** 11, 2.0 ** 10])
    Kt = (K + np.diag([0, 1e6])) / s / s[:, None]
    eps = 2 * np.finfo(float).eps * np.max(np.sum(np.abs(Kt), axis=1))
    assert dec.eps == pytest.approx(eps * np.min(s * s), rel=1e-15)


def test_not_posdef():
    with pytest.raises(np.linalg.LinAlgError):
        Chol(-np.eye(3))
