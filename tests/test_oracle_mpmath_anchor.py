"""Implementation-independent anchor for the numbers the reference holds no golden values for (SURVEY.md section 8c:
"logML / posterior-mean values: parity unpinned"): log marginal likelihood, its gradient and the posterior mean of a
small instance of config C2 (Matern-5/2 + noise, 3-D) computed with mpmath at 50 digits straight from the formulas
(reference _linalg/_decomp.py:380-393,484-488,505-512; _GP/_compute.py:255-260): closed-form kernel, exact
equilibration (powers of two), jitter eps = n 2^-52 max_i sum_j |K~_ij| added on the scaled matrix, LU-free Cholesky in
multiprecision.  The oracle must agree to 1e-11 (it is float64 LAPACK); the GPU test tier checks the CUDA path
against the same numbers (tests/test_gpu_api.py::test_mpmath_anchor)."""
import json
import pathlib

import numpy as np
import pytest

from oracle import gp as ogp

GOLD = pathlib.Path(__file__).resolve().parent / 'golden' / 'mpmath_anchor_c2_n48.json'


def problem():
    rng = np.random.default_rng(4242)
    n, m = 48, 7
    X = rng.uniform(0, 10, (n, 3))
    y = np.sin(X[:, 0]) + np.cos(X[:, 1]) * X[:, 2] / 10 + 0.1 * rng.standard_normal(n)
    Xs = rng.uniform(0, 10, (m, 3))
    return X, y, Xs, (1.5, 1.0, 0.1)


def mp_reference():
    """ the 50-digit computation (used to write the fixture; re-run by the test when mpmath is present) """
    import mpmath as mp
    mp.mp.dps = 50
    X, y, Xs, (ell, sf, sn) = problem()
    n, m = len(X), len(Xs)
    f = lambda v: mp.mpf(float(v))
    ell_, sf_, sn_ = f(ell), f(sf), f(sn)

    def kern(a, b, same):
        # operation order of the reference: scale each argument, then difference (float64 roundings reproduced by
        # doing the division in float64: the inputs of the multiprecision formula are the float64 scaled points)
        r2 = mp.mpf(0)
        for q in range(3):
            d = f(float(a[q]) / ell) - f(float(b[q]) / ell)
            r2 += d * d
        x = mp.sqrt(5 * r2)
        core = mp.exp(-x) * (1 + x + x * x / 3)
        dcore = -mp.exp(-x) * (1 + x) / 6 * 5          # d core / d r2
        return sf_ ** 2 * core + (sn_ ** 2 if same else 0), core, dcore * r2 * (-2)

    K = mp.matrix(n, n)
    dK = [mp.matrix(n, n) for _ in range(3)]  # d/d log ell, d/d log sf, d/d log sn
    for i in range(n):
        for j in range(n):
            k, core, dlogell = kern(X[i], X[j], i == j)
            K[i, j] = k
            dK[0][i, j] = sf_ ** 2 * dlogell
            dK[1][i, j] = 2 * sf_ ** 2 * core
            dK[2][i, j] = 2 * sn_ ** 2 if i == j else 0
    # Chol.__init__: equilibrate, jitter
    s = [mp.mpf(2) ** int(mp.nint(mp.log(K[i, i], 2) / 2)) for i in range(n)]
    Kt = mp.matrix(n, n)
    for i in range(n):
        for j in range(n):
            Kt[i, j] = K[i, j] / s[i] / s[j]
    eps = n * mp.mpf(2) ** -52 * max(sum(abs(Kt[i, j]) for j in range(n)) for i in range(n))
    for i in range(n):
        Kt[i, i] += eps
    Kreg = mp.matrix(n, n)
    for i in range(n):
        for j in range(n):
            Kreg[i, j] = Kt[i, j] * s[i] * s[j]
    L = mp.cholesky(Kreg)
    yv = mp.matrix([f(v) for v in y])
    a = mp.lu_solve(L, yv)
    b = mp.lu_solve(Kreg, yv)
    logdet = 2 * sum(mp.log(L[i, i]) for i in range(n))
    value = (n * mp.log(2 * mp.pi) + logdet + (a.T * a)[0]) / 2
    invK = Kreg ** -1
    grad = []
    for D in dK:
        t = sum(invK[i, j] * D[j, i] for i in range(n) for j in range(n))
        q = (b.T * D * b)[0]
        grad.append((t - q) / 2)
    Ks = mp.matrix(n, m)
    for i in range(n):
        for j in range(m):
            Ks[i, j] = kern(X[i], Xs[j], False)[0]
    mean = Ks.T * b
    return dict(logml=float(-value), grad_minus_logml=[float(g) for g in grad], mean=[float(v) for v in mean],
                eps_scaled=float(eps))


def oracle_values():
    X, y, Xs, (ell, sf, sn) = problem()
    terms = [(sf ** 2, [dict(kind='matern', nu=2.5, scale=ell)]), (sn ** 2, [dict(kind='white')])]
    xt, xs = X.T.copy(), Xs.T.copy()
    val, g = ogp.logml_and_grad(terms, xt, y, [('logscale', 0, 0), ('amp', 0), ('amp', 1)])
    grad = [g[0], g[1] * 2 * sf ** 2, g[2] * 2 * sn ** 2]
    K = ogp.gram(terms, xt, xt)
    Kxs = ogp.gram(terms[:1], xt, xs)
    Kss = ogp.gram(terms[:1], xs, xs)
    mean, _ = ogp.pred(K, Kxs, Kss, y)
    return -val, np.array(grad), mean


def test_fixture_matches_mpmath():
    mp = pytest.importorskip('mpmath')
    ref = mp_reference()
    gold = json.loads(GOLD.read_text())
    assert abs(ref['logml'] - gold['logml']) <= 1e-15 * abs(gold['logml'])
    np.testing.assert_allclose(ref['grad_minus_logml'], gold['grad_minus_logml'], rtol=1e-14)
    np.testing.assert_allclose(ref['mean'], gold['mean'], rtol=1e-14)


def test_oracle_against_anchor():
    gold = json.loads(GOLD.read_text())
    logml, grad, mean = oracle_values()
    assert abs(logml - gold['logml']) <= 1e-11 * abs(gold['logml'])
    np.testing.assert_allclose(grad, gold['grad_minus_logml'], rtol=1e-9, atol=1e-9 * np.max(np.abs(gold['grad_minus_logml'])))
    np.testing.assert_allclose(mean, gold['mean'], rtol=1e-9, atol=1e-11)


if __name__ == '__main__':  # writes the fixture
    GOLD.write_text(json.dumps(mp_reference(), indent=1) + '\n')
    print(GOLD.read_text())
