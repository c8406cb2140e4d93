"""GPU parity at BASELINE.json's full size (configs[1]: Matern-5/2, 3-D, n = 20000), through size-independent properties
(the CPU oracle needs minutes at this size): sampled Gram rows against the oracle, exact symmetry, L L^T = K + jitter,
solve residuals, K^-1 K = I on random vectors, log-determinant against an independent factorisation, and the gradient of
logML against central finite differences of the value.  Tolerances: Gram 1e-13 relative (north_star), factor products
1e-12, residuals 1e-10, logML-level quantities 1e-9."""
import numpy as np
import pytest
import torch

from lsqfitgp_b200 import _lib, _ops
from oracle import gp as ogp

pytestmark = pytest.mark.gpu
N = 20000


def _data(n=N, seed=2002):
    rng = np.random.default_rng(seed)
    X = rng.uniform(0, 10, (n, 3))
    y = np.sin(X[:, 0]) + np.cos(X[:, 1]) * X[:, 2] / 10 + 0.1 * rng.standard_normal(n)
    return X, y


def _descs(theta):
    ell, sf, sn = np.exp(theta)
    return [dict(kind=_lib.K_MATERNP, term=0, dimmask=7, ipar=2, par0=0.0, scale_x=ell, scale_y=ell, amp=sf ** 2),
            dict(kind=_lib.K_WHITE, term=1, dimmask=7, amp=sn ** 2)]


THETA0 = np.array([np.log(1.5), 0.0, np.log(0.1)])


@pytest.fixture(scope='module')
def state():
    dev = torch.device('cuda:0')
    X, y = _data()
    xd = torch.tensor(np.ascontiguousarray(X.T)).to(dev)
    K = _ops.aligned_empty(N, N, dev)
    _ops.gram_iso(_descs(THETA0), xd, xd, out=K, symmetric=True)
    st = _ops.chol_factor(K)
    assert int(st.info.item()) == 0
    return dict(dev=dev, X=X, y=y, xd=xd, K=K, st=st)


def test_gram_sampled_rows_and_symmetry(state):
    X, K = state['X'], state['K']
    rows = np.random.default_rng(1).choice(N, 48, replace=False)
    rows[0], rows[1] = 0, N - 1
    terms = [(1.0, [dict(kind='matern', nu=2.5, scale=1.5)]), (0.01, [dict(kind='white')])]
    Ko = ogp.gram(terms, X[rows].T.copy(), X.T.copy())          # 48 x 20000 entries through the oracle
    Kg = K[torch.as_tensor(rows, device=K.device)].cpu().numpy()
    assert np.max(np.abs(Kg - Ko) / np.abs(Ko)) < 1e-13
    # the symmetric build mirrors tiles: exactly symmetric, and equal to the non-symmetric build bit for bit
    assert torch.equal(K[:4096, :4096], K[:4096, :4096].T)
    blk = _ops.gram_iso(_descs(THETA0), state['xd'][:, 7000:9000].contiguous(), state['xd'][:, 100:1300].contiguous())
    assert torch.equal(blk, K[7000:9000, 100:1300])


def test_factor_reproduces_matrix(state):
    K, st, dev = state['K'], state['st'], state['dev']
    g = torch.Generator(device='cpu').manual_seed(3)
    V = torch.randn(N, 4, generator=g, dtype=torch.float64).to(dev)
    LtV = _ops.chol_mult(st, V, True)
    LLtV = _ops.chol_mult(st, LtV, False)
    sc = st.scalars()
    s = st.aux[:N]
    KV = K @ V + (sc[1] * s * s)[:, None] * V               # K + eps diag(s^2): what Chol factors (_decomp.py:384-387)
    err = float((LLtV - KV).norm() / KV.norm())
    assert err < 1e-12, err


def test_solve_residual_and_inverse(state):
    K, st, dev, y = state['K'], state['st'], state['dev'], state['y']
    b = torch.tensor(y).to(dev)[:, None]
    a = _ops.chol_solve(st, b, False)
    x = _ops.chol_solve(st, a, True)
    sc = st.scalars()
    s = st.aux[:N]
    r = K @ x + (sc[1] * s * s)[:, None] * x - b
    assert float(r.norm() / b.norm()) < 1e-10
    # inverse from the factor: Kinv (K v) = v on random vectors (lower triangle is what the library defines)
    Kinv = _ops.chol_inverse(st)
    Kl = torch.tril(Kinv)
    g = torch.Generator(device='cpu').manual_seed(4)
    v = torch.randn(N, 2, generator=g, dtype=torch.float64).to(dev)
    w = K @ v + (sc[1] * s * s)[:, None] * v
    u = Kl @ w + torch.tril(Kinv, -1).T @ w
    assert float((u - v).norm() / v.norm()) < 1e-9
    del Kinv, Kl


def test_logdet_against_independent_factorisation(state):
    K, st = state['K'], state['st']
    sc = st.scalars()
    s = st.aux[:N]
    Kj = K.clone()
    Kj.diagonal().add_(sc[1] * s * s)
    ld_ref = float(torch.log(torch.diagonal(torch.linalg.cholesky(Kj))).sum())   # cuSOLVER as the independent checker
    ld = float(sc[4])
    assert abs(ld - ld_ref) <= 1e-12 * abs(ld_ref)
    del Kj


def _value_and_grad(state, theta, grad=True):
    dev, xd, y = state['dev'], state['xd'], state['y']
    yd = torch.tensor(y).to(dev)
    descs = _descs(theta)
    K = _ops.aligned_empty(N, N, dev)
    _ops.gram_iso(descs, xd, xd, out=K, symmetric=True)
    st = _ops.chol_factor(K)
    a = _ops.chol_solve(st, yd[:, None], False)
    ld, q = _ops.chol_logdet_quad(st, a[:, 0].contiguous()).cpu().numpy()
    val = 0.5 * (N * np.log(2 * np.pi) + 2 * ld + q)
    if not grad:
        return val
    b = _ops.chol_solve(st, a, True, inplace=True)
    Kinv = _ops.chol_inverse(st)
    v = _ops.gram_iso_vjp(descs, xd, Kinv, b[:, 0].contiguous()).cpu().numpy()
    ell, sf, sn = np.exp(theta)
    return val, 0.5 * np.array([v[0, 1], v[0, 0] * 2 * sf ** 2, v[1, 0] * 2 * sn ** 2])


def test_gradient_against_finite_differences(state):
    val, grad = _value_and_grad(state, THETA0)
    h = 1e-5
    for k in range(3):
        e = np.zeros(3)
        e[k] = h
        fd = (_value_and_grad(state, THETA0 + e, grad=False) - _value_and_grad(state, THETA0 - e, grad=False)) / (2 * h)
        assert abs(fd - grad[k]) <= 1e-5 * np.max(np.abs(grad)), (k, fd, grad)


def test_matern_real_order_full_size_sampled_rows():
    """ Matern(nu = 1.3) at n = 20000 (K_nu evaluated in the kernel): sampled rows against the oracle's scipy.special.kv route
    (2e-13: AMOS itself is ~1e-13 from exact, tests/test_bessel_cpu.py), exact symmetry of the general kernel's output """
    import time
    dev = torch.device('cuda:0')
    X, _ = _data()
    xd = torch.tensor(np.ascontiguousarray(X.T)).to(dev)
    descs = [dict(kind=_lib.K_MATERN, term=0, dimmask=7, par0=1.3, scale_x=1.5, scale_y=1.5, amp=1.0),
             dict(kind=_lib.K_WHITE, term=1, dimmask=7, amp=0.01)]
    K = _ops.aligned_empty(N, N, dev)
    _ops.gram_iso(descs, xd, xd, out=K, symmetric=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    _ops.gram_iso(descs, xd, xd, out=K, symmetric=True)
    torch.cuda.synchronize()
    print(f'Matern(nu=1.3) Gram n={N}: {(time.perf_counter() - t0) * 1e3:.1f} ms')
    rows = np.random.default_rng(2).choice(N, 24, replace=False)
    terms = [(1.0, [dict(kind='matern', nu=1.3, scale=1.5)]), (0.01, [dict(kind='white')])]
    Ko = ogp.gram(terms, X[rows].T.copy(), X.T.copy())
    Kg = K[torch.as_tensor(rows, device=dev)].cpu().numpy()
    assert np.max(np.abs(Kg - Ko) / np.abs(Ko)) < 2e-13
    assert torch.equal(K[:2048, :2048], K[:2048, :2048].T)


def test_logml_value_against_oracle_at_full_size(state):
    """ the VALUE of logML at n = 20000 against the CPU oracle (round-1 verdict: only logdet / residual / finite-difference
    checks existed at this size).  The oracle evaluates the Gram matrix with the closed half-integer form Maternp(p=2)
    (equal to the reference's kv-based Matern(nu=2.5) to 1e-15, tests/test_oracle_special.py, and 20x cheaper than 4e8 calls
    of scipy.special.kv) into ONE host buffer and factors it in place with LAPACK (oracle.gp.logml_value_lean); the sampled
    rows of test_gram_sampled_rows_and_symmetry pin the device Gram to the kv form.  Tolerance 1e-9 (north_star). """
    import psutil
    if psutil.virtual_memory().available < 2.5 * 8 * N * N:
        pytest.skip('host RAM')
    X, y, st, dev = state['X'], state['y'], state['st'], state['dev']
    terms = [(1.0, [dict(kind='maternp', p=2, scale=1.5)]), (0.01, [dict(kind='white')])]
    val_o, L, eps_o = ogp.logml_value_lean(terms, X.T.copy(), y)
    ld_o = float(np.sum(np.log(np.diagonal(L))))
    del L
    a = _ops.chol_solve(st, torch.tensor(y, device=dev)[:, None], False)
    ld, q = _ops.chol_logdet_quad(st, a[:, 0].contiguous()).cpu().numpy()
    val = 0.5 * (N * np.log(2 * np.pi) + 2 * ld + q)
    assert abs(ld - ld_o) <= 1e-11 * abs(ld_o)
    assert abs(val - val_o) <= 1e-9 * abs(val_o)
    s = st.scalars().cpu().numpy()
    assert abs(s[1] * s[3] - eps_o) <= 1e-12 * eps_o
