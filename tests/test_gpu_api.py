"""GPU parity through the public API (GP / kernels / Chol / empbayes_fit) against the oracle and the golden
fixtures.  Tolerances from BASELINE.json north_star: Gram 1e-13 relative; logML, gradient, posterior mean 1e-9."""
import pathlib

import numpy as np
import pytest
import torch
from scipy import optimize, stats

import lsqfitgp_b200 as lgp
from oracle import gp as ogp, bart as obart, decomp as odecomp

pytestmark = pytest.mark.gpu
GOLD = pathlib.Path(__file__).resolve().parent / 'golden'


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


def test_config1_full_size():
    """ C1: lgp.GP(lgp.ExpQuad()) 1-D regression n=1000: marginal_likelihood + predfromdata """
    rng = np.random.default_rng(1001)
    x = np.sort(rng.uniform(0, 100, 1000))
    y = np.sin(x / 3) + 0.1 * rng.standard_normal(1000)
    xpred = np.linspace(-5, 105, 500)
    gp = lgp.GP(lgp.ExpQuad(scale=3)).addx(x, 'data').addx(xpred, 'pred')
    ycov = 0.01 * np.eye(1000)
    ml = gp.marginal_likelihood({'data': y}, {('data', 'data'): ycov})
    terms = [(1.0, [dict(kind='expquad', scale=3)])]
    Kxx = ogp.gram(terms, x[None], x[None])
    Kxs = ogp.gram(terms, x[None], xpred[None])
    Kss = ogp.gram(terms, xpred[None], xpred[None])
    ml_o = ogp.logml(Kxx, y, ycov)
    assert abs(ml - ml_o) / abs(ml_o) < 1e-9
    m, c = gp.predfromdata({'data': y}, 'pred', {('data', 'data'): ycov}, raw=True)
    m_o, c_o = ogp.pred(Kxx, Kxs, Kss, y, ycov)
    assert rel(m, m_o) < 1e-9
    assert np.max(np.abs(c - c_o)) < 1e-9
    prior = gp.prior('data', raw=True)
    assert np.max(np.abs(prior - Kxx) / np.abs(Kxx).clip(1e-300)) < 1e-13
    md, cd = gp.predfromdata({'data': y}, ['pred', 'data'], {('data', 'data'): ycov}, raw=True)
    assert rel(md['pred'], m_o) < 1e-9 and cd['pred', 'data'].shape == (500, 1000)


def test_predfromfit_vs_oracle():
    """ fromdata=False: cov = Kss - Ksx K^-1 Kxs + A' ycov A, A = K^-1 Kxs (reference _compute.py:262-269) """
    rng = np.random.default_rng(77)
    x = np.sort(rng.uniform(0, 30, 120))
    xp = np.linspace(0, 30, 31)
    y = rng.standard_normal(120)
    k = lgp.ExpQuad(scale=3) + 0.1 * lgp.White()
    gp = lgp.GP(k).addx(x, 'data').addx(xp, 'pred')
    ycov = np.diag(rng.uniform(0.01, 0.1, 120))
    mf, cf = gp.predfromfit({'data': y}, 'pred', {('data', 'data'): ycov}, raw=True)
    terms = [(1.0, [dict(kind='expquad', scale=3)]), (0.1, [dict(kind='white')])]
    Kxx = ogp.gram(terms, x[None], x[None])
    Kxs = ogp.gram(terms, x[None], xp[None])
    Kss = ogp.gram(terms, xp[None], xp[None])
    do = odecomp.Chol(Kxx)
    A = do.ginv_linear(Kxs)
    assert rel(mf, do.pinv_bilinear(Kxs, y)) < 1e-9
    assert rel(cf, Kss - do.ginv_quad(Kxs) + A.T @ ycov @ A) < 1e-9


def test_golden_c1():
    g = np.load(GOLD / 'c1_expquad.npz')
    n = len(g['x'])
    gp = lgp.GP(lgp.ExpQuad(scale=3)).addx(g['x'], 'data').addx(g['xpred'], 'pred')
    ycov = {('data', 'data'): 0.01 * np.eye(n)}
    K = gp.prior('data', raw=True)
    assert np.max(np.abs(K[0] - g['gram_row0']) / g['gram_row0'].clip(1e-300)) < 1e-13
    assert abs(gp.marginal_likelihood({'data': g['y']}, ycov) - g['logml']) / abs(g['logml']) < 1e-9
    m, c = gp.predfromdata({'data': g['y']}, 'pred', ycov, raw=True)
    assert rel(m, g['mean']) < 1e-9
    assert np.max(np.abs(np.diag(c) - g['cov_diag'])) < 1e-9


def _c2_gp(X, theta):
    xs = lgp.unstructured_to_structured(X, names=['f0', 'f1', 'f2'])
    ell, sf, sn = torch.exp(theta[0]), torch.exp(theta[1]), torch.exp(theta[2])
    kern = sf ** 2 * lgp.Matern(nu=2.5, scale=ell) + sn ** 2 * lgp.White()
    return lgp.GP(kern, checkpos=False, checksym=False).addx(xs, 'data')


def test_golden_c2_value_and_gradient():
    g = np.load(GOLD / 'c2_matern.npz')
    theta = torch.tensor(g['theta'], dtype=torch.float64, requires_grad=True)
    gp = _c2_gp(g['X'], theta)
    ml = gp.marginal_likelihood({'data': g['y']})
    grad, = torch.autograd.grad(ml, theta)
    assert abs(float(ml.detach()) + g['minus_logml']) / abs(g['minus_logml']) < 1e-9
    assert rel(-grad.numpy(), g['grad_minus_logml']) < 1e-9
    K = gp.prior('data', raw=True)
    assert np.max(np.abs(K[0] - g['gram_row0']) / g['gram_row0']) < 1e-13
    assert np.max(np.abs(K[:, 5] - g['gram_col5']) / g['gram_col5']) < 1e-13


def test_config2_mid_size_vs_oracle():
    """ C2 at n=1500 (oracle gradient is O(n^3) on the CPU): value, gradient, Gram """
    rng = np.random.default_rng(2002)
    n = 1500
    X = rng.uniform(0, 10, (n, 3))
    y = np.sin(X[:, 0]) + np.cos(X[:, 1]) * X[:, 2] / 10 + 0.1 * rng.standard_normal(n)
    theta = torch.tensor([np.log(1.5), 0.0, np.log(0.1)], dtype=torch.float64, requires_grad=True)
    gp = _c2_gp(X, theta)
    ml = gp.marginal_likelihood({'data': y})
    g, = torch.autograd.grad(ml, theta)
    terms = [(1.0, [dict(kind='matern', nu=2.5, scale=1.5)]), (0.01, [dict(kind='white')])]
    val_o, grad_o = ogp.logml_and_grad(terms, X.T.copy(), y, [('logscale', 0, 0), ('amp', 0), ('amp', 1)])
    grad_o = np.array([grad_o[0], grad_o[1] * 2.0, grad_o[2] * 2 * 0.01])
    assert abs(float(ml.detach()) + val_o) / abs(val_o) < 1e-9
    assert rel(-g.numpy(), grad_o) < 1e-9
    Kg = gp.prior('data', raw=True)
    Ko = ogp.gram(terms, X.T.copy(), X.T.copy())
    assert np.max(np.abs(Kg - Ko) / np.abs(Ko)) < 1e-13
    # no-grad path gives the same value as a float
    with torch.no_grad():
        ml2 = _c2_gp(X, theta.detach()).marginal_likelihood({'data': y})
    assert isinstance(ml2, float) and abs(ml2 - float(ml.detach())) <= 1e-12 * abs(ml2)


def test_config3_empbayes_fit_vs_oracle():
    rng = np.random.default_rng(3003)
    n = 400
    X3 = rng.uniform(0, 100, (n, 2))
    K3 = ogp.gram([(1.3 ** 2, [dict(kind='expquad', scale=8.0)])], X3.T.copy(), X3.T.copy())
    y3 = np.linalg.cholesky(K3 + 1e-10 * np.eye(n)) @ rng.standard_normal(n) + 0.2 * rng.standard_normal(n)
    x3 = lgp.unstructured_to_structured(X3, names=['a', 'b'])
    hyperprior = {'log(ell)': (np.log(3), 1.0), 'log(sf)': (0.0, 1.0), 'log(sn)': (np.log(0.1), 1.0)}

    def gpfactory(hp):
        k = hp['sf'] ** 2 * lgp.ExpQuad(scale=hp['ell']) + hp['sn'] ** 2 * lgp.White()
        return lgp.GP(k, checkpos=False, checksym=False).addx(x3, 'data')
    fit = lgp.empbayes_fit(hyperprior, gpfactory, {'data': y3}, raises=False)

    def obj(p):
        hp = np.array([np.log(3), 0.0, np.log(0.1)]) + p
        terms = [(np.exp(hp[1]) ** 2, [dict(kind='expquad', scale=np.exp(hp[0]))]),
                 (np.exp(hp[2]) ** 2, [dict(kind='white')])]
        v, g = ogp.logml_and_grad(terms, X3.T.copy(), y3, [('logscale', 0, 0), ('amp', 0), ('amp', 1)])
        g = np.array([g[0], g[1] * 2 * np.exp(hp[1]) ** 2, g[2] * 2 * np.exp(hp[2]) ** 2])
        return v + 0.5 * (3 * np.log(2 * np.pi) + p @ p), g + p
    res = optimize.minimize(obj, np.zeros(3), jac=True, method='bfgs')
    assert np.max(np.abs(fit.minresult.x - res.x)) < 1e-5      # reference tests/test_fit.py:142-176 uses atol 1e-5
    assert abs(fit.minresult.fun - res.fun) / abs(res.fun) < 1e-9
    assert set(fit.pmean) == set(hyperprior) and fit.pcov['log(ell)', 'log(sf)'].shape == ()
    # same optimum without gradients
    fit2 = lgp.empbayes_fit(hyperprior, gpfactory, {'data': y3}, raises=False, method='nograd',
                            minkw=dict(options=dict(xatol=1e-7, fatol=1e-10, maxiter=2000)))
    assert np.max(np.abs(fit2.minresult.x - res.x)) < 1e-4
    # fixed parameter and explicit starting point
    fit3 = lgp.empbayes_fit(hyperprior, gpfactory, {'data': y3}, raises=False, fix={'sn': True},
                            initial={'log(ell)': np.log(5.0), 'log(sf)': 0.1, 'log(sn)': np.log(0.2)})
    assert fit3.pmean['log(sn)'] == pytest.approx(np.log(0.2), abs=1e-15)
    # additional_loss offsets the objective exactly (reference tests/test_fit.py:281-310)
    fit4 = lgp.empbayes_fit(hyperprior, gpfactory, {'data': y3}, raises=False, additional_loss=lambda hp: 3.5)
    assert fit4.minresult.fun == pytest.approx(fit.minresult.fun + 3.5, rel=1e-12)


def test_golden_c4_bart():
    g = np.load(GOLD / 'c4_bart.npz')
    splits = lgp.BART.splits_from_coord(g['X'])
    assert np.array_equal(splits[0], g['length'])
    idx = lgp.BART.indices_from_coord(g['X'], splits)
    assert np.array_equal(idx, g['idx'])
    names = [f'c{i}' for i in range(idx.shape[1])]
    xi = lgp.unstructured_to_structured(idx.astype(np.int32), names=names)
    kb = lgp.BART(splits=splits, indices=True, alpha=0.95, beta=2, maxd=10, reset=[2, 4, 6, 8], gamma=1)
    n = len(idx)
    gp = lgp.GP(kb + 0.1 * lgp.White(), checkpos=False, checksym=False, epsrel=0).addx(xi, 'train')
    K = gp.prior('train', raw=True)
    assert np.max(np.abs(K - (g['gram'] + 0.1 * np.eye(n))) / (g['gram'] + 0.1 * np.eye(n))) < 1e-13
    ml = gp.marginal_likelihood({'train': g['y']})
    assert abs(ml - g['logml']) / abs(g['logml']) < 1e-9


def test_bart_variants_vs_oracle():
    rng = np.random.default_rng(4004)
    n = 300
    X4 = np.concatenate([rng.standard_normal((n, 8)), rng.integers(0, 2, (n, 2)).astype(float)], axis=1)
    splits = lgp.BART.splits_from_coord(X4)
    idx = lgp.BART.indices_from_coord(X4, splits)
    xi = lgp.unstructured_to_structured(idx.astype(np.int32), names=[f'c{i}' for i in range(10)])
    for kw in [dict(maxd=2), dict(maxd=1), dict(maxd=0), dict(maxd=4, reset=2), dict(maxd=2, gamma=0.3, intercept=False),
               dict(maxd=10, reset=[2, 4, 6, 8]), dict(maxd=5, reset=[1, 2, 3, 4], gamma=0.0),
               dict(maxd=6, reset=[2, 4], weights=np.r_[np.ones(5), 0., 2., 3., 0.5, 1.]),
               dict(pnt=[0.9, 0.5, 0.2])]:
        kb = lgp.BART(splits=splits, indices=True, **kw)
        Kg = lgp.GP(kb, checkpos=False, checksym=False).addx(xi, 'train').prior('train', raw=True)
        Ko = obart.gram(splits[0], idx, idx, **kw)
        assert np.max(np.abs(Kg - Ko) / np.abs(Ko)) < 1e-13, kw
    # coordinates instead of indices: identical (reference tests/kernels/test_bart.py:355-367, 0 ulp)
    dt = [(f'c{i}', float) for i in range(10)]
    xa = X4[:, None, :].copy().view(dt).squeeze(-1)
    ya = X4[None, :40, :].copy().view(dt).squeeze(-1)
    Kc = lgp.BART(splits=splits, indices=False, maxd=4, reset=2)(xa, ya)
    ia = idx[:, None, :].astype(np.int64).copy().view([(f'c{i}', np.int64) for i in range(10)]).squeeze(-1)
    ib = idx[None, :40, :].astype(np.int64).copy().view([(f'c{i}', np.int64) for i in range(10)]).squeeze(-1)
    Ki = lgp.BART(splits=splits, indices=True, maxd=4, reset=2)(ia, ib)
    assert np.array_equal(Kc, Ki)
    # correlation classmethod: pair form and count form (reference test_altinput)
    nn = splits[0]
    c_alt = lgp.BART.correlation(nn, idx[:50], idx[50:100], maxd=4, reset=2, altinput=True)
    lo, hi = np.minimum(idx[:50], idx[50:100]), np.maximum(idx[:50], idx[50:100])
    c_cnt = lgp.BART.correlation(lo, hi - lo, nn - hi, maxd=4, reset=2)
    np.testing.assert_allclose(c_alt, c_cnt, rtol=1e-15, atol=2e-15)
    np.testing.assert_allclose(c_alt, obart.correlation(nn, idx[:50], idx[50:100], maxd=4, reset=2), rtol=1e-13)
    empty = np.array([], int)
    assert lgp.BART.correlation(empty, empty, empty) == 1


def test_bart_recipe_logml():
    """ bayestree.bart recipe (reference bayestree/_bart.py:187-205): lambda^2 BART + diag noise + constant """
    rng = np.random.default_rng(4005)
    n = 200
    X = rng.standard_normal((n, 4))
    splits = lgp.BART.splits_from_coord(X)
    idx = lgp.BART.indices_from_coord(X, splits)
    xi = lgp.unstructured_to_structured(idx.astype(np.int32), names=list('abcd'))
    lam, sig, kk = 1.3, 0.5, 0.7
    gp = (lgp.GP(lam ** 2 * lgp.BART(splits=splits, indices=True, maxd=10, reset=[2, 4, 6, 8]), checkpos=False,
                 checksym=False, epsrel=0)
          .addx(xi, 'trainmean').addcov(sig ** 2 * np.eye(n), 'trainnoise').addcov(kk ** 2, 'mean')
          .addtransf({'trainmean': 1, 'trainnoise': 1, 'mean': 1}, 'train'))
    y = rng.standard_normal(n)
    ml = gp.marginal_likelihood({'train': y})
    Ko = lam ** 2 * obart.gram(splits[0], idx, idx, maxd=10, reset=[2, 4, 6, 8]) + sig ** 2 * np.eye(n) + kk ** 2
    ml_o = ogp.logml(Ko, y, epsrel=0)
    assert abs(ml - ml_o) / abs(ml_o) < 1e-9
    m, c = gp.predfromdata({'train': y}, 'trainmean', raw=True)
    Kf = lam ** 2 * obart.gram(splits[0], idx, idx, maxd=10, reset=[2, 4, 6, 8])
    m_o, c_o = ogp.pred(Ko, Kf, Kf, y, epsrel=0)
    assert rel(m, m_o) < 1e-9


@pytest.mark.parametrize('n', [1, 2, 10, 300])
def test_chol_class_battery(n):
    """ every Decomposition method against the oracle (reference tests/linalg/test_decomp.py:145-261) """
    rng = np.random.default_rng(n)
    O = stats.ortho_group.rvs(n, random_state=rng) if n > 1 else np.atleast_2d(1.0)
    K = (O * (1 + 1e-3 + np.cos(1 + np.arange(n)))) @ O.T
    K = (K + K.T) / 2
    dec = lgp._linalg.Chol(K)
    do = odecomp.Chol(K)
    B = rng.standard_normal((n, 3))
    r = rng.standard_normal(n)
    assert abs(dec.eps - do.eps) / do.eps < 1e-12 and dec.n == n and dec.m == n
    assert rel(dec.ginv_linear(B), do.ginv_linear(B)) < 1e-9
    assert rel(dec.ginv_linear(r), do.ginv_linear(r)) < 1e-9
    assert rel(dec.pinv_bilinear(B, r), do.pinv_bilinear(B, r)) < 1e-9
    assert rel(dec.pinv_bilinear_robj(B, r), do.pinv_bilinear(B, r)) < 1e-9
    assert rel(dec.ginv_quad(B), do.ginv_quad(B)) < 1e-9
    assert rel(dec.ginv_diagquad(B), do.ginv_diagquad(B)) < 1e-9
    assert rel(dec.correlate(B), do.correlate(B)) < 1e-11
    assert rel(dec.back_correlate(B), do.back_correlate(B)) < 1e-11
    assert rel(dec.pinv_correlate(r), do.pinv_correlate(r)) < 1e-9
    assert rel(dec.ginv(), do.ginv()) < 1e-9
    assert dec.matrix() is K
    v1 = dec.minus_log_normal_density(r, value=True)[0]
    v2 = do.minus_log_normal_density(r, value=True)[0]
    assert abs(v1 - v2) / abs(v2) < 1e-10
    dK = rng.standard_normal((n, n, 2))
    dK = dK + dK.transpose(1, 0, 2)
    dr = rng.standard_normal((n, 2))
    o1 = dec.minus_log_normal_density(r, dK=dK, dr=dr, gradfwd=True, fisher=True)
    o2 = do.minus_log_normal_density(r, dK=dK, dr=dr, gradfwd=True, fisher=True)
    assert o1[0] is None and o1[1] is None and o1[4] is None
    assert rel(o1[2], o2[2]) < 1e-8 and rel(o1[3], o2[3]) < 1e-8
    vj = lambda G: np.einsum('ij,ijk->k', G, dK)
    rj = lambda g: g @ dr
    vec = rng.standard_normal(2)
    kw = dict(dK_vjp=vj, dr_vjp=rj, dK_jvp_vec=dK @ vec, dr_jvp_vec=dr @ vec, gradrev=True, fishvec=True)
    o1 = dec.minus_log_normal_density(r, **kw)
    o2 = do.minus_log_normal_density(r, **kw)
    assert rel(o1[1], o2[1]) < 1e-8 and rel(o1[4], o2[4]) < 1e-8


def test_value_path_fused_gram_matches_block_path():
    """ GP._solver on one set of points with a kernel of the fast family: the Gram build is fused with the equilibration
    pass (no covariance block); same logML, posterior mean and covariance as when the block has been built first (the
    two-pass path), and the finiteness check still reports non-finite points """
    rng = np.random.default_rng(12)
    n = 700
    X = rng.uniform(0, 10, (n, 2))
    y = np.sin(X[:, 0]) + 0.1 * rng.standard_normal(n)
    x = lgp.unstructured_to_structured(X, names=['a', 'b'])
    xp = lgp.unstructured_to_structured(rng.uniform(0, 10, (40, 2)), names=['a', 'b'])
    kern = 1.3 * lgp.Matern(nu=2.5, scale=1.7) + 0.05 * lgp.White()
    fused = lgp.GP(kern, checkpos=False).addx(x, 'd').addx(xp, 'p')
    block = lgp.GP(kern, checkpos=False).addx(x, 'd').addx(xp, 'p')
    block.prior('d', raw=True)                       # builds and caches the block: _solver takes the two-pass path
    assert ('d', 'd') in block._covblocks and ('d', 'd') not in fused._covblocks
    ml_f, ml_b = fused.marginal_likelihood({'d': y}), block.marginal_likelihood({'d': y})
    assert fused._solver(['d'])._K is None and block._solver(['d'])._K is not None
    assert abs(ml_f - ml_b) <= 1e-12 * abs(ml_b)
    m_f, c_f = fused.predfromdata({'d': y}, 'p', raw=True)
    m_b, c_b = block.predfromdata({'d': y}, 'p', raw=True)
    np.testing.assert_allclose(m_f, m_b, rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(c_f, c_b, rtol=1e-9, atol=1e-12)
    Xbad = X.copy()
    Xbad[3, 1] = np.nan
    bad = lgp.GP(kern, checkpos=False).addx(lgp.unstructured_to_structured(Xbad, names=['a', 'b']), 'd')
    with pytest.raises(RuntimeError, match='not finite'):
        bad.marginal_likelihood({'d': y})


def test_error_semantics():
    with pytest.raises(np.linalg.LinAlgError):
        lgp._linalg.Chol(-np.eye(5))
    gp = lgp.GP(lgp.ExpQuad()).addx(np.arange(5.), 'a')
    with pytest.raises(KeyError):
        gp.addx(np.arange(3.), 'a')
    with pytest.raises(ValueError):
        gp.addx(np.arange(3.))
    with pytest.raises(KeyError):
        gp.marginal_likelihood({'zzz': np.zeros(5)})
    with pytest.raises(ValueError):
        gp.marginal_likelihood({'a': np.zeros(4)})
    with pytest.raises(ValueError, match='not finite'):
        gp.marginal_likelihood({'a': np.array([0, 1, np.nan, 0, 0])})
    with pytest.raises(ValueError, match='not symmetric'):
        gp.marginal_likelihood({'a': np.zeros(5)}, {('a', 'a'): np.triu(np.ones((5, 5)))})
    with pytest.raises(TypeError):
        gp.marginal_likelihood([1, 2, 3])
    with pytest.raises(ValueError):
        gp.pred({'a': np.zeros(5)}, 'a', raw=True)
    with pytest.raises(NotImplementedError):
        gp.pred({'a': np.zeros(5)}, 'a', fromdata=True)
    # immutability: addx returns a new object
    gp2 = gp.addx(np.arange(2.), 'b')
    assert 'b' not in gp._elements and 'b' in gp2._elements
    # decomposition cache reused when no ycov (reference tests/GP/test_GP.py:587-596)
    d1 = gp2._solver(['a'])
    assert gp2._solver(['a']) is d1
    dec = lgp.GP.decompose(np.eye(4).reshape(2, 2, 2, 2))
    assert dec.n == 4


def test_kernel_call_broadcasting():
    rng = np.random.default_rng(5)
    x = rng.uniform(0, 5, 7)
    y = rng.uniform(0, 5, 4)
    k = lgp.ExpQuad(scale=2) * 1.5 + lgp.Cauchy(beta=3)
    K = k(x[:, None], y[None, :])
    ref = 1.5 * np.exp(-0.5 * ((x[:, None] / 2 - y[None, :] / 2) ** 2)) + (1 + (x[:, None] - y[None, :]) ** 2 / 3) ** (-3 / 2)
    np.testing.assert_allclose(K, ref, rtol=1e-13)
    np.testing.assert_allclose(k(x, x), np.full(7, 2.5), rtol=1e-14)          # elementwise on equal shapes
    xs = lgp.StructuredArray({'a': x, 'b': x[::-1].copy()})
    ka = lgp.ExpQuad(dim='a')
    np.testing.assert_allclose(ka(xs.reshape(7, 1), xs.reshape(1, 7)), np.exp(-0.5 * (x[:, None] - x[None, :]) ** 2),
                               rtol=1e-13)
    with pytest.raises(ValueError):
        ka(x, x)


@pytest.mark.parametrize('nu', [0.3, 1.3, 4.2])
def test_matern_real_order_vs_oracle(nu):
    """ Matern of non-half-integer order: K_nu evaluated in the Gram kernel (csrc/bessel_k.cuh) where the reference
    calls scipy.special.kv on the host (_special/_bessel.py:35,70-99).  The oracle follows the reference route (scipy kv,
    AMOS), which is itself up to ~1e-13 from the exact value (tests/test_bessel_cpu.py): Gram tolerance 2e-13 here. """
    rng = np.random.default_rng(int(nu * 10))
    n = 700
    X = rng.uniform(0, 10, (n, 3))
    y = np.sin(X[:, 0]) + np.cos(X[:, 1]) * X[:, 2] / 10 + 0.1 * rng.standard_normal(n)
    xs = lgp.unstructured_to_structured(X, names=['f0', 'f1', 'f2'])
    theta = torch.tensor([np.log(1.5), 0.0, np.log(0.1)], dtype=torch.float64, requires_grad=True)
    kern = torch.exp(theta[1]) ** 2 * lgp.Matern(nu=nu, scale=torch.exp(theta[0])) + torch.exp(theta[2]) ** 2 * lgp.White()
    gp = lgp.GP(kern, checkpos=False, checksym=False).addx(xs, 'data')
    ml = gp.marginal_likelihood({'data': y})
    g, = torch.autograd.grad(ml, theta)
    terms = [(1.0, [dict(kind='matern', nu=nu, scale=1.5)]), (0.01, [dict(kind='white')])]
    Kg = gp.prior('data', raw=True)
    Ko = ogp.gram(terms, X.T.copy(), X.T.copy())
    assert np.max(np.abs(Kg - Ko) / np.abs(Ko)) < 2e-13
    val_o, grad_o = ogp.logml_and_grad(terms, X.T.copy(), y, [('logscale', 0, 0), ('amp', 0), ('amp', 1)])
    grad_o = np.array([grad_o[0], grad_o[1] * 2.0, grad_o[2] * 2 * 0.01])
    assert abs(float(ml.detach()) + val_o) / abs(val_o) < 1e-9
    assert rel(-g.numpy(), grad_o) < 1e-9


def test_matern_order_edge_cases():
    x = np.linspace(0, 3, 40)
    # nu = 0: white noise (_matern.py:74); half-integer orders keep the closed form; nu > 100 refused, never a CPU fallback
    K0 = lgp.Matern(nu=0)(x[:, None], x[None, :])
    assert np.array_equal(K0, np.eye(40))
    Kh = lgp.Matern(nu=1.5)(x[:, None], x[None, :])
    Kc = lgp.Matern(nu=1.5 + 1e-9)(x[:, None], x[None, :])
    assert np.max(np.abs(Kh - Kc)) < 1e-8
    with pytest.raises(NotImplementedError):
        lgp.Matern(nu=150.0)


def test_empbayes_fit_fisher_vs_oracle():
    """ method='fisher' (dogleg with Fisher information + prior precision as hessian; reference _fit.py:732-743,
    765-769) reaches the same optimum as BFGS; the Fisher matrix equals the oracle's
    1/2 tr(K^-1 dK_i K^-1 dK_j) + I (reference _decomp.py:535-558) with dK by central differences of the oracle Gram """
    rng = np.random.default_rng(3005)
    n = 300
    X3 = rng.uniform(0, 100, (n, 2))
    K3 = ogp.gram([(1.3 ** 2, [dict(kind='expquad', scale=8.0)])], X3.T.copy(), X3.T.copy())
    y3 = np.linalg.cholesky(K3 + 1e-10 * np.eye(n)) @ rng.standard_normal(n) + 0.2 * rng.standard_normal(n)
    x3 = lgp.unstructured_to_structured(X3, names=['a', 'b'])
    hyperprior = {'log(ell)': (np.log(3), 1.0), 'log(sf)': (0.0, 1.0), 'log(sn)': (np.log(0.1), 1.0)}

    def gpfactory(hp):
        k = hp['sf'] ** 2 * lgp.ExpQuad(scale=hp['ell']) + hp['sn'] ** 2 * lgp.White()
        return lgp.GP(k, checkpos=False, checksym=False).addx(x3, 'data')
    fitg = lgp.empbayes_fit(hyperprior, gpfactory, {'data': y3}, raises=False)
    fitf = lgp.empbayes_fit(hyperprior, gpfactory, {'data': y3}, raises=False, method='fisher', covariance='fisher')
    assert np.max(np.abs(fitf.minresult.x - fitg.minresult.x)) < 1e-4
    assert abs(fitf.minresult.fun - fitg.minresult.fun) / abs(fitg.minresult.fun) < 1e-8

    def Kof(p):
        hp = np.array([np.log(3), 0.0, np.log(0.1)]) + p
        terms = [(np.exp(hp[1]) ** 2, [dict(kind='expquad', scale=np.exp(hp[0]))]),
                 (np.exp(hp[2]) ** 2, [dict(kind='white')])]
        return ogp.gram(terms, X3.T.copy(), X3.T.copy())
    p = fitf.minresult.x
    h = 1e-5
    dK = np.stack([(Kof(p + h * e) - Kof(p - h * e)) / (2 * h) for e in np.eye(3)], axis=2)
    do = odecomp.Chol(Kof(p))
    fo = do.minus_log_normal_density(y3, dK=dK, fisher=True)[3] + np.eye(3)
    assert rel(fitf.minresult.hess, fo) < 1e-6          # limited by the finite differences of the check
    # covariance='fisher': inverse of that matrix, mapped to the hyperparameters (prior sdev 1: identity jacobian)
    cov = np.array([[fitf.pcov[a, b] for b in hyperprior] for a in hyperprior])
    assert rel(cov, np.linalg.inv(fo)) < 1e-5
    # covariance='fisher' after a gradient fit evaluates the Fisher matrix at the optimum
    fitc = lgp.empbayes_fit(hyperprior, gpfactory, {'data': y3}, raises=False, covariance='fisher')
    covc = np.array([[fitc.pcov[a, b] for b in hyperprior] for a in hyperprior])
    assert rel(covc, cov) < 1e-3


def test_mpmath_anchor():
    """ the CUDA path against the 50-digit mpmath computation of logML, its gradient and the posterior mean
    (tests/golden/mpmath_anchor_c2_n48.json, written by tests/test_oracle_mpmath_anchor.py): an anchor that does not
    go through the oracle, LAPACK or float64 at all """
    import json
    from test_oracle_mpmath_anchor import problem
    gold = json.loads((GOLD / 'mpmath_anchor_c2_n48.json').read_text())
    X, y, Xs, (ell, sf, sn) = problem()
    names = ['f0', 'f1', 'f2']
    xs = lgp.unstructured_to_structured(X, names=names)
    xp = lgp.unstructured_to_structured(Xs, names=names)
    theta = torch.tensor([np.log(ell), np.log(sf), np.log(sn)], dtype=torch.float64, requires_grad=True)
    main = torch.exp(theta[1]) ** 2 * lgp.Matern(nu=2.5, scale=torch.exp(theta[0]))
    kern = main + torch.exp(theta[2]) ** 2 * lgp.White()
    gp = lgp.GP(kern, checkpos=False, checksym=False).addx(xs, 'data')
    ml = gp.marginal_likelihood({'data': y})
    g, = torch.autograd.grad(ml, theta)
    assert abs(float(ml.detach()) - gold['logml']) <= 1e-11 * abs(gold['logml'])
    gg = np.array(gold['grad_minus_logml'])
    np.testing.assert_allclose(-g.numpy(), gg, rtol=1e-9, atol=1e-9 * np.max(np.abs(gg)))
    # posterior mean of the latent function at new points: the noise is White, absent between distinct points
    with torch.no_grad():
        gp2 = lgp.GP(sf ** 2 * lgp.Matern(nu=2.5, scale=ell) + sn ** 2 * lgp.White(), checkpos=False, checksym=False)
        gp2 = gp2.addx(xs, 'data').addx(xp, 'pred')
        mean, cov = gp2.predfromdata({'data': y}, 'pred', raw=True)
    np.testing.assert_allclose(mean, gold['mean'], rtol=1e-9, atol=1e-11)


def test_batch_in_flight_matches_sequential():
    """ several evaluations in flight on one GPU (one thread + stream per slot) give the sequential results """
    from lsqfitgp_b200 import _dist
    rng = np.random.default_rng(77)
    n = 900
    X = rng.uniform(0, 10, (n, 3))
    y = np.sin(X[:, 0]) + 0.1 * rng.standard_normal(n)
    xs = lgp.unstructured_to_structured(X, names=['f0', 'f1', 'f2'])

    def fun(theta):
        th = torch.tensor(theta, dtype=torch.float64, requires_grad=True)
        k = torch.exp(th[1]) ** 2 * lgp.Matern(nu=2.5, scale=torch.exp(th[0])) + torch.exp(th[2]) ** 2 * lgp.White()
        ml = lgp.GP(k, checkpos=False, checksym=False).addx(xs, 'data').marginal_likelihood({'data': y})
        g, = torch.autograd.grad(ml, th)
        return np.r_[float(ml.detach()), g.numpy()]
    thetas = np.array([np.log(1.5), 0.0, np.log(0.1)]) + 0.2 * rng.standard_normal((9, 3))
    dev = torch.device('cuda', torch.cuda.current_device())
    seq = _dist.eval_batch_sharded(fun, thetas, device=dev)
    par = _dist.eval_batch_sharded(fun, thetas, device=dev, in_flight=3)
    np.testing.assert_allclose(par[:, 0], seq[:, 0], rtol=1e-13)
    np.testing.assert_allclose(par[:, 1:], seq[:, 1:], rtol=1e-10, atol=1e-10 * np.abs(seq[:, 1:]).max())


def test_reference_gp_shell_semantics():
    """ behaviour pinned by the reference's own tests/GP/test_GP.py (file:line cited per block), raw=True forms """
    rng = np.random.default_rng(20)
    # test_zero_covblock (:509-515): independent addcov keys have a zero cross block
    a = rng.standard_normal((10, 10))
    m = a.T @ a
    gp = lgp.GP().addcov(m, 0).addcov(m, 1)
    prior = gp.prior(raw=True)
    assert np.array_equal(prior[0, 1], np.zeros_like(m)) and np.allclose(prior[0, 0], m, rtol=1e-15)
    # test_addcov_checks (:517-530): non-symmetric / non-finite blocks are refused unless the checks are disabled
    b = a.copy()
    b[0, 0] = np.inf
    with pytest.raises(ValueError):
        lgp.GP().addcov(a, 0)
    with pytest.raises(ValueError):
        lgp.GP().addcov(b.T @ b, 0)
    lgp.GP(checksym=False).addcov(a, 0)
    lgp.GP(checkfinite=False).addcov(b.T @ b, 0)
    # test_zero_givencov (:652-660): a zero data covariance changes nothing
    x, y, z = rng.standard_normal((3, 20))
    gp = lgp.GP(lgp.ExpQuad()).addx(x, 0).addx(y, 1)
    m1, c1 = gp.predfromdata({0: z}, 1, {(0, 0): np.zeros((20, 20))}, raw=True)
    m2, c2 = gp.predfromdata({0: z}, 1, raw=True)
    np.testing.assert_array_equal(m1, m2)
    np.testing.assert_array_equal(c1, c2)
    # test_pred_all (:683-690): no key = all keys
    ma, ca = gp.predfromdata({0: z}, raw=True)
    mb, cb = gp.predfromdata({0: z}, [0, 1], raw=True)
    assert set(ma) == {0, 1} and all(np.array_equal(ma[k], mb[k]) for k in ma)
    assert all(np.array_equal(ca[k], cb[k]) for k in ca)
    # test_pred_checks (:662-681)
    with pytest.raises(ValueError):
        gp.pred({0: z}, 1)
    with pytest.raises(ValueError):
        gp.predfromdata({0: z}, 1, raw=True, keepcorr=True)
    with pytest.raises(ValueError):
        gp.predfromdata({0: np.full_like(z, np.nan)}, 1, raw=True)
    with pytest.raises(ValueError):
        gp.predfromdata({0: z}, 1, {(0, 0): np.full((20, 20), np.nan)}, raw=True)
    with pytest.raises(ValueError):
        gp.predfromdata({0: z}, 1, {(0, 0): rng.standard_normal((20, 20))}, raw=True)
    with pytest.raises(KeyError):
        gp.predfromdata({2: z}, 1, raw=True)
    with pytest.raises(ValueError):
        gp.predfromdata({0: z[:-1]}, 1, raw=True)
    # test_marginal_likelihood_checks (:692-705, without the gvar form)
    with pytest.raises(ValueError):
        gp.marginal_likelihood({0: np.full_like(z, np.nan)})
    with pytest.raises(ValueError):
        gp.marginal_likelihood({0: z}, {(0, 0): np.full((20, 20), np.nan)})
    with pytest.raises(ValueError):
        gp.marginal_likelihood({0: z}, {(0, 0): rng.standard_normal((20, 20))})
    # test_marginal_likelihood_gvar (:707-715), matrix form: equals the oracle
    c = rng.standard_normal((20, 20))
    c = c.T @ c
    ml = gp.marginal_likelihood({0: z}, {(0, 0): c})
    Kxx = ogp.gram([(1.0, [dict(kind='expquad')])], x[None], x[None])
    assert abs(ml - ogp.logml(Kxx, z, c)) <= 1e-12 * abs(ml)


def test_gradfwd_equals_gradrev_and_kernel_contractions():
    """ Chol.minus_log_normal_density: forward-mode gradient (explicit dK, contracted by lgp_symlower_dot over the lower
    triangle of the inverse) == reverse-mode gradient (dK_vjp called on the full matrix written by lgp_sym_expand_sub)
    == oracle (reference _decomp.py:505-531); ginv_diagquad through lgp_colsumsq """
    rng = np.random.default_rng(17)
    n, k = 333, 3
    A = rng.standard_normal((n, n))
    K = A @ A.T / n + np.eye(n)
    r = rng.standard_normal(n)
    dK = rng.standard_normal((n, n, k))
    dK = dK + dK.transpose(1, 0, 2)
    dKn = dK + 0.1 * rng.standard_normal((n, n, k))       # not exactly symmetric: the contraction must not assume it
    dec, ref = lgp._linalg.Chol(K), odecomp.Chol(K)
    for d in (dK, dKn):
        _, _, gf, _, _ = dec.minus_log_normal_density(r, dK=d, gradfwd=True)
        _, _, gf_o, _, _ = ref.minus_log_normal_density(r, dK=d, gradfwd=True)
        np.testing.assert_allclose(gf, gf_o, rtol=1e-10, atol=1e-12)
        _, gr, _, _, _ = dec.minus_log_normal_density(r, dK_vjp=lambda G, d=d: np.einsum('ij,ijk->k', G, d), gradrev=True)
        np.testing.assert_allclose(gr, gf_o, rtol=1e-10, atol=1e-12)
    _, _, gl, _, _ = dec.minus_log_normal_density(r, dK=[dK[:, :, q].copy() for q in range(k)], gradfwd=True)
    np.testing.assert_allclose(gl, gf_o if False else ref.minus_log_normal_density(r, dK=dK, gradfwd=True)[2], rtol=1e-10)
    B = rng.standard_normal((n, 37))
    np.testing.assert_allclose(dec.ginv_diagquad(B), ref.ginv_diagquad(B), rtol=1e-10)
    np.testing.assert_allclose(dec.ginv_diagquad(B[:, 0]), ref.ginv_diagquad(B[:, :1])[0], rtol=1e-10)


def test_raniter_sample_match_oracle_path():
    """ raniter / sample (reference _fastraniter.py:36-121): same seed -> mean + L z with the oracle's Chol factor """
    rng = np.random.default_rng(3)
    n = 50
    A = rng.standard_normal((n, n))
    cov = A @ A.T / n + 0.1 * np.eye(n)
    mean = rng.standard_normal(n)
    L = odecomp.Chol(cov)._L
    zs = np.random.default_rng(99)
    got = list(lgp.raniter(mean, cov, n=3, rng=99))
    for g in got:
        np.testing.assert_allclose(g, mean + L @ zs.standard_normal(n), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(lgp.sample(mean, cov, rng=99), got[0], rtol=1e-13)
    batch = lgp.sample_batch(mean, cov, 3, rng=99)
    np.testing.assert_allclose(batch, np.stack(got), rtol=1e-10, atol=1e-12)
    # dictionary form, scalar form, empirical covariance of a device-generated batch
    md = {'a': mean[:20].reshape(4, 5), 'b': mean[20:]}
    cd = {('a', 'a'): cov[:20, :20].reshape(4, 5, 4, 5), ('a', 'b'): cov[:20, 20:].reshape(4, 5, 30),
          ('b', 'a'): cov[20:, :20].reshape(30, 4, 5), ('b', 'b'): cov[20:, 20:]}
    sd = lgp.sample(md, cd, rng=99)
    np.testing.assert_allclose(np.r_[sd['a'].reshape(-1), sd['b']], got[0], rtol=1e-10, atol=1e-12)
    assert isinstance(lgp.sample(1.0, 4.0, rng=1), float)
    big = lgp.sample_batch(mean, cov, 20000, rng=5, device_rng=True)
    emp = np.cov(big.T)
    assert np.max(np.abs(emp - cov)) < 0.05 * np.max(np.abs(cov))
    with pytest.raises(np.linalg.LinAlgError):
        lgp.sample(np.zeros(2), np.array([[1.0, 2.0], [2.0, 1.0]]), eps=0)


def test_solver_chol_dist_single_process_matches_chol():
    """ GP(..., solver='chol-dist') (reference seam _compute.py:424-428) on a 1 x 1 process grid: same marginal
    likelihood and posterior as solver='chol'; GP.decompose with the distributed class; multi-rank runs of the same code
    are in tests/test_gpu_dist.py """
    rng = np.random.default_rng(8)
    n = 700
    X = rng.uniform(0, 10, (n, 2))
    y = np.sin(X[:, 0]) + 0.1 * rng.standard_normal(n)
    xs = lgp.unstructured_to_structured(X, names=['a', 'b'])
    kern = 1.3 * lgp.ExpQuad(scale=1.7) + 0.01 * lgp.White()
    out = {}
    for solver in ('chol', 'chol-dist'):
        gp = lgp.GP(kern, solver=solver, checkpos=False, checksym=False).addx(xs, 'd').addx(xs[:40], 'p')
        out[solver] = (gp.marginal_likelihood({'d': y}), *gp.predfromdata({'d': y}, 'p', raw=True))
    assert abs(out['chol'][0] - out['chol-dist'][0]) <= 1e-11 * abs(out['chol'][0])
    np.testing.assert_allclose(out['chol-dist'][1], out['chol'][1], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(out['chol-dist'][2], out['chol'][2], rtol=1e-8, atol=1e-10)
    K = ogp.gram([(1.3, [dict(kind='expquad', scale=1.7)]), (0.01, [dict(kind='white')])], X.T.copy(), X.T.copy())
    dd = lgp.GP.decompose(K, solver='chol-dist')
    ref = odecomp.Chol(K)
    b = rng.standard_normal(n)
    np.testing.assert_allclose(dd.ginv_linear(b), ref.ginv_linear(b), rtol=1e-8)
    np.testing.assert_allclose(dd.correlate(b), ref.correlate(b), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(dd.ginv_diagquad(np.stack([b, 2 * b], 1)), ref.ginv_diagquad(np.stack([b, 2 * b], 1)), rtol=1e-9)
    assert dd.eps == pytest.approx(ref.eps, rel=1e-14) and dd.n == n
    with pytest.raises(NotImplementedError):
        th = torch.tensor(1.7, dtype=torch.float64, requires_grad=True)
        lgp.GP(lgp.ExpQuad(scale=th), solver='chol-dist').addx(xs, 'd').marginal_likelihood({'d': y})


def test_empbayes_multistart_sharded_batch():
    """ multistart: the starting points go through eval_batch_sharded as one batch; the best one starts the minimiser """
    rng = np.random.default_rng(12)
    n = 300
    X = rng.uniform(0, 30, (n, 1))
    y = np.sin(X[:, 0]) + 0.1 * rng.standard_normal(n)
    xs = lgp.unstructured_to_structured(X, names=['t'])
    hyperprior = {'log(ell)': (np.log(10.0), 1.5), 'log(sn)': (np.log(0.3), 1.0)}

    def gpfactory(hp):
        return lgp.GP(lgp.ExpQuad(scale=hp['ell']) + hp['sn'] ** 2 * lgp.White(), checkpos=False, checksym=False).addx(xs, 'd')
    fit = lgp.empbayes_fit(hyperprior, gpfactory, {'d': y}, raises=False, multistart=6, in_flight=2)
    ms = fit.multistart
    assert ms['starts'].shape == (7, 2) and np.all(ms['starts'][0] == 0) and ms['values'][ms['best']] == ms['values'].min()
    assert np.array_equal(fit.minargs['x0'], ms['starts'][ms['best']])
    plain = lgp.empbayes_fit(hyperprior, gpfactory, {'d': y}, raises=False)
    assert fit.minresult.fun <= plain.minresult.fun + 1e-6


def test_bart_indices_device_kernel_matches_host():
    """ lgp_searchsorted (BART.indices_from_coord on the device, reference _bart.py:294-299,503-514): bit-equal to the host """
    rng = np.random.default_rng(4)
    X = np.concatenate([rng.standard_normal((500, 3)), rng.integers(0, 4, (500, 2)).astype(float)], axis=1)
    splits = lgp.BART.splits_from_coord(X)
    want = lgp.BART.indices_from_coord(X, splits)
    from lsqfitgp_b200._kernels import _bart_indices_device
    xd = torch.tensor(np.ascontiguousarray(X.T), device='cuda')
    got = _bart_indices_device(xd, splits[1]).cpu().numpy().T
    assert np.array_equal(got, want)
    Xq = rng.standard_normal((77, 5)) * 3          # points outside the training range
    got = _bart_indices_device(torch.tensor(np.ascontiguousarray(Xq.T), device='cuda'), splits[1]).cpu().numpy().T
    assert np.array_equal(got, lgp.BART.indices_from_coord(Xq, splits))


def test_chol_on_two_devices_in_one_process():
    """ stream pools and kernel attributes are per device: factor on cuda:0 then on cuda:1 """
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    rng = np.random.default_rng(2)
    n = 1500
    A = rng.standard_normal((n, n))
    K = A @ A.T / n + np.eye(n)
    b = rng.standard_normal(n)
    want = odecomp.Chol(K).ginv_linear(b)
    for d in (0, 1, 0):
        with torch.cuda.device(d):
            dec = lgp._linalg.Chol(K)
            np.testing.assert_allclose(dec.ginv_linear(b), want, rtol=1e-9)
            low = dec.inverse_lower()
            assert low.device.index == d and torch.isfinite(torch.tril(low)).all()


def test_get_factor_beyond_grid_y_limit():
    """ row counts above 65535 (gridDim.y limit): lgp_chol_get_factor on a synthetic factor state """
    n = 65536 + 300
    st = _ops_state(n)
    L = lgp._ops.chol_get_factor(st)
    rows = torch.tensor([0, 1, 65534, 65535, 65536, n - 1], device='cuda')
    sub = L.index_select(0, rows).cpu().numpy()
    for q, i in enumerate(rows.cpu().numpy()):
        assert np.all(sub[q, :i + 1] == 2.0 * 0.5) and np.all(sub[q, i + 1:] == 0.0), i
    del L, st
    torch.cuda.empty_cache()


def _ops_state(n):
    from lsqfitgp_b200 import _ops, _lib
    lib = _lib.load()
    st = _ops.FactorState()
    st.n, st.npad, st.device = n, int(lib.lgp_chol_npad(n)), torch.device('cuda', torch.cuda.current_device())
    st.W = torch.full((st.npad, st.npad), 0.5, dtype=torch.float64, device=st.device)
    st.aux = torch.full((int(lib.lgp_chol_aux_doubles(n)),), 2.0, dtype=torch.float64, device=st.device)
    st.info = torch.zeros(1, dtype=torch.int32, device=st.device)
    return st
