"""On-GPU check of the public API against the oracle (configs C1, small C2/C3/C4)."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, '.')
import lsqfitgp_b200 as lgp
from oracle import gp as ogp, iso as oiso, bart as obart, decomp as odecomp

ok = True
def report(name, err, tol):
    global ok
    good = bool(err <= tol)
    ok &= good
    print(f'{"PASS" if good else "FAIL"} {name}: err={err:.3e} tol={tol:.1e}', flush=True)
def rel(a, b):
    a = np.asarray(a); b = np.asarray(b)
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))

# ---------------- C1
rng = np.random.default_rng(1001)
x = np.sort(rng.uniform(0, 100, 1000))
y = np.sin(x / 3) + 0.1 * rng.standard_normal(1000)
xpred = np.linspace(-5, 105, 500)
gp = lgp.GP(lgp.ExpQuad(scale=3)).addx(x, 'data').addx(xpred, 'pred')
ycov = 0.01 * np.eye(1000)
ml = gp.marginal_likelihood({'data': y}, {('data', 'data'): ycov})
terms = [(1.0, [dict(kind='expquad', scale=3)])]
Kxx = ogp.gram(terms, x[None], x[None]); Kxs = ogp.gram(terms, x[None], xpred[None]); Kss = ogp.gram(terms, xpred[None], xpred[None])
ml_o = ogp.logml(Kxx, y, ycov)
report('C1 logML', abs(ml - ml_o) / abs(ml_o), 1e-9)
m, c = gp.predfromdata({'data': y}, 'pred', {('data', 'data'): ycov}, raw=True)
m_o, c_o = ogp.pred(Kxx, Kxs, Kss, y, ycov)
report('C1 posterior mean', rel(m, m_o), 1e-9)
report('C1 posterior cov', float(np.max(np.abs(c - c_o))), 1e-9)
prior = gp.prior('data', raw=True)
report('C1 prior gram', float(np.max(np.abs(prior - Kxx) / np.abs(Kxx).clip(1e-300))), 1e-13)
md, cd = gp.predfromdata({'data': y}, ['pred', 'data'], {('data', 'data'): ycov}, raw=True)
report('C1 pred dict mean', rel(md['pred'], m_o), 1e-9)

# ---------------- C2-like (n=1500) value + gradient
rng = np.random.default_rng(2002)
n = 1500
X = rng.uniform(0, 10, (n, 3))
yy = np.sin(X[:, 0]) + np.cos(X[:, 1]) * X[:, 2] / 10 + 0.1 * rng.standard_normal(n)
xs = lgp.unstructured_to_structured(X, names=['f0', 'f1', 'f2'])
theta = torch.tensor([np.log(1.5), 0.0, np.log(0.1)], dtype=torch.float64, requires_grad=True)
ell, sf, sn = torch.exp(theta[0]), torch.exp(theta[1]), torch.exp(theta[2])
kern = sf ** 2 * lgp.Matern(nu=2.5, scale=ell) + sn ** 2 * lgp.White()
gp2 = lgp.GP(kern, checkpos=False, checksym=False).addx(xs, 'data')
ml2 = gp2.marginal_likelihood({'data': yy})
g, = torch.autograd.grad(ml2, theta)
terms = [(1.0, [dict(kind='matern', nu=2.5, scale=1.5)]), (0.01, [dict(kind='white')])]
val_o, grad_o = ogp.logml_and_grad(terms, X.T.copy(), yy, [('logscale', 0, 0), ('amp', 0), ('amp', 1)])
grad_o = np.array([grad_o[0], grad_o[1] * 2 * 1.0, grad_o[2] * 2 * 0.01])  # d/dlog sf = 2 sf^2 d/d amp
report('C2s logML', abs(float(ml2) + val_o) / abs(val_o), 1e-9)
report('C2s grad', rel(-g.numpy(), grad_o), 1e-9)
Kg = gp2.prior('data', raw=True)
Ko = ogp.gram(terms, X.T.copy(), X.T.copy())
report('C2s gram (Matern kv oracle vs closed form)', float(np.max(np.abs(Kg - Ko) / np.abs(Ko))), 1e-13)

# ---------------- C3-like: empbayes_fit on ExpQuad + noise (n=400)
rng = np.random.default_rng(3003)
n = 400
X3 = rng.uniform(0, 100, (n, 2))
truth = dict(ell=8.0, sf=1.3, sn=0.2)
K3 = ogp.gram([(truth['sf'] ** 2, [dict(kind='expquad', scale=truth['ell'])])], X3.T.copy(), X3.T.copy())
y3 = np.linalg.cholesky(K3 + 1e-10 * np.eye(n)) @ rng.standard_normal(n) + truth['sn'] * rng.standard_normal(n)
x3 = lgp.unstructured_to_structured(X3, names=['a', 'b'])
hyperprior = {'log(ell)': (np.log(3), 1.0), 'log(sf)': (0.0, 1.0), 'log(sn)': (np.log(0.1), 1.0)}
def gpfactory(hp):
    k = hp['sf'] ** 2 * lgp.ExpQuad(scale=hp['ell']) + hp['sn'] ** 2 * lgp.White()
    return lgp.GP(k, checkpos=False, checksym=False).addx(x3, 'data')
t0 = time.time()
fit = lgp.empbayes_fit(hyperprior, gpfactory, {'data': y3}, raises=False)
print('fit time', time.time() - t0, 'nfev', fit.minresult.nfev, {k: np.exp(v) for k, v in fit.pmean.items()})
# oracle optimisation with scipy on the same objective
from scipy import optimize
def obj(p):
    hp = np.array([np.log(3), 0.0, np.log(0.1)]) + p
    terms = [(np.exp(hp[1]) ** 2, [dict(kind='expquad', scale=np.exp(hp[0]))]), (np.exp(hp[2]) ** 2, [dict(kind='white')])]
    v, g = ogp.logml_and_grad(terms, X3.T.copy(), y3, [('logscale', 0, 0), ('amp', 0), ('amp', 1)])
    g = np.array([g[0], g[1] * 2 * np.exp(hp[1]) ** 2, g[2] * 2 * np.exp(hp[2]) ** 2])
    return v + 0.5 * (3 * np.log(2 * np.pi) + p @ p), g + p
res = optimize.minimize(obj, np.zeros(3), jac=True, method='bfgs')
report('C3s fit optimum', float(np.max(np.abs(fit.minresult.x - res.x))), 1e-5)
report('C3s fit objective', abs(fit.minresult.fun - res.fun) / abs(res.fun), 1e-9)

# ---------------- C4-like: BART gram n=300 p=10
rng = np.random.default_rng(4004)
n = 300
X4 = np.concatenate([rng.standard_normal((n, 8)), rng.integers(0, 2, (n, 2)).astype(float)], axis=1)
splits = lgp.BART.splits_from_coord(X4)
splits_o = obart.splits_from_coord(X4)
report('C4s splits length', float(np.max(np.abs(splits[0] - splits_o[0]))), 0)
idx = lgp.BART.indices_from_coord(X4, splits)
idx_o = obart.indices_from_coord(X4, splits_o)
report('C4s indices', float(np.max(np.abs(idx - idx_o))), 0)
kb = lgp.BART(splits=splits, indices=True, alpha=0.95, beta=2, maxd=10, reset=[2, 4, 6, 8], gamma=1)
xi = lgp.unstructured_to_structured(idx.astype(np.int32), names=[f'c{i}' for i in range(10)])
gp4 = lgp.GP(kb, checkpos=False, checksym=False).addx(xi, 'train')
K4 = gp4.prior('train', raw=True)
K4o = obart.gram(splits_o[0], idx_o, idx_o, alpha=0.95, beta=2, maxd=10, reset=[2, 4, 6, 8], gamma=1)
report('C4s BART gram maxd=10 reset=[2,4,6,8]', float(np.max(np.abs(K4 - K4o) / np.abs(K4o))), 1e-13)
for kw in [dict(maxd=2), dict(maxd=1), dict(maxd=0), dict(maxd=4, reset=2), dict(maxd=2, gamma=0.3, intercept=False),
           dict(maxd=6, reset=[2, 4], weights=np.r_[np.ones(5), 0., 2., 3., 0.5, 1.])]:
    kb = lgp.BART(splits=splits, indices=True, **kw)
    Kg = lgp.GP(kb, checkpos=False, checksym=False).addx(xi, 'train').prior('train', raw=True)
    Ko = obart.gram(splits_o[0], idx_o, idx_o, **kw)
    report(f'C4s BART gram {kw if "weights" not in kw else "weights"}', float(np.max(np.abs(Kg - Ko) / np.abs(Ko))), 1e-13)
kc = lgp.BART(splits=splits, indices=False, maxd=4, reset=2)
Kc = kc(X4[:, None, :].view([(f'c{i}', float) for i in range(10)]).squeeze(-1), X4[None, :40, :].view([(f'c{i}', float) for i in range(10)]).squeeze(-1))
Kco = obart.gram(splits_o[0], idx_o, idx_o[:40], maxd=4, reset=2)
report('C4s BART from coordinates', float(np.max(np.abs(Kc - Kco) / np.abs(Kco))), 1e-13)
# full recipe: lambda^2 BART + noise via addcov + constant, logML
lam, sig, kk = 1.3, 0.5, 0.7
gp5 = (lgp.GP(lam ** 2 * lgp.BART(splits=splits, indices=True, maxd=10, reset=[2, 4, 6, 8]), checkpos=False, checksym=False, epsrel=0)
       .addx(xi, 'trainmean').addcov(sig ** 2 * np.eye(n), 'trainnoise').addcov(kk ** 2, 'mean')
       .addtransf({'trainmean': 1, 'trainnoise': 1, 'mean': 1}, 'train'))
y4 = rng.standard_normal(n)
ml5 = gp5.marginal_likelihood({'train': y4})
Ko5 = lam ** 2 * K4o + sig ** 2 * np.eye(n) + kk ** 2
ml5o = ogp.logml(Ko5, y4, epsrel=0)
report('C4s bart recipe logML', abs(ml5 - ml5o) / abs(ml5o), 1e-9)

# ---------------- Chol methods battery (reference tests/linalg/test_decomp.py) at n=10, 300
for n in (1, 2, 10, 300):
    rng = np.random.default_rng(n)
    from scipy import stats, linalg
    O = stats.ortho_group.rvs(n, random_state=rng) if n > 1 else np.atleast_2d(1)
    eigvals = 1 + 1e-3 + np.cos(1 + np.arange(n))
    K = (O * eigvals) @ O.T
    K = (K + K.T) / 2
    dec = lgp._linalg.Chol(K)
    do = odecomp.Chol(K)
    B = rng.standard_normal((n, 3)); r = rng.standard_normal(n)
    report(f'Chol n={n} eps', abs(dec.eps - do.eps) / do.eps, 1e-12)
    report(f'Chol n={n} ginv_linear', rel(dec.ginv_linear(B), do.ginv_linear(B)), 1e-9)
    report(f'Chol n={n} pinv_bilinear', rel(dec.pinv_bilinear(B, r), do.pinv_bilinear(B, r)), 1e-9)
    report(f'Chol n={n} ginv_quad', rel(dec.ginv_quad(B), do.ginv_quad(B)), 1e-9)
    report(f'Chol n={n} ginv_diagquad', rel(dec.ginv_diagquad(B), do.ginv_diagquad(B)), 1e-9)
    report(f'Chol n={n} correlate', rel(dec.correlate(B), do.correlate(B)), 1e-11)
    report(f'Chol n={n} back_correlate', rel(dec.back_correlate(B), do.back_correlate(B)), 1e-11)
    report(f'Chol n={n} pinv_correlate', rel(dec.pinv_correlate(r), do.pinv_correlate(r)), 1e-9)
    report(f'Chol n={n} ginv', rel(dec.ginv(), do.ginv()), 1e-9)
    v1 = dec.minus_log_normal_density(r, value=True)[0]; v2 = do.minus_log_normal_density(r, value=True)[0]
    report(f'Chol n={n} value', abs(v1 - v2) / abs(v2), 1e-10)
    dK = rng.standard_normal((n, n, 2)); dK = dK + dK.transpose(1, 0, 2); dr = rng.standard_normal((n, 2))
    o1 = dec.minus_log_normal_density(r, dK=dK, dr=dr, gradfwd=True, fisher=True)
    o2 = do.minus_log_normal_density(r, dK=dK, dr=dr, gradfwd=True, fisher=True)
    report(f'Chol n={n} gradfwd', rel(o1[2], o2[2]), 1e-8)
    report(f'Chol n={n} fisher', rel(o1[3], o2[3]), 1e-8)
    vj = lambda G: np.einsum('ij,ijk->k', G, dK); rj = lambda g: g @ dr
    vec = rng.standard_normal(2)
    o1 = dec.minus_log_normal_density(r, dK_vjp=vj, dr_vjp=rj, dK_jvp_vec=dK @ vec, dr_jvp_vec=dr @ vec, gradrev=True, fishvec=True)
    o2 = do.minus_log_normal_density(r, dK_vjp=vj, dr_vjp=rj, dK_jvp_vec=dK @ vec, dr_jvp_vec=dr @ vec, gradrev=True, fishvec=True)
    report(f'Chol n={n} gradrev', rel(o1[1], o2[1]), 1e-8)
    report(f'Chol n={n} fishvec', rel(o1[4], o2[4]), 1e-8)
try:
    lgp._linalg.Chol(-np.eye(5))
    ok = False; print('FAIL: no LinAlgError')
except np.linalg.LinAlgError:
    print('PASS LinAlgError on non-PD')
print('ALL OK' if ok else 'SOME FAILED')
sys.exit(0 if ok else 1)
