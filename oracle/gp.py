"""Oracle: the GP-level hot path (NumPy/SciPy restatement).

  Gram of points            src/lsqfitgp/_GP/_elements.py:554-579 (kernel(ax[:, None], ay[None, :]))
  Kxx + ycov -> Chol        src/lsqfitgp/_GP/_compute.py:45-94
  marginal_likelihood       src/lsqfitgp/_GP/_compute.py:383-422 -> _linalg/_decomp.py:441-490
  predfromdata (raw)        src/lsqfitgp/_GP/_compute.py:230-260 -> _decomp.py:405-420
  logML gradient            src/lsqfitgp/_fit.py:687-702 + _decomp.py:505-512, with analytic dK/dtheta in place of jax.vjp
The kernel is described by a list of terms; each term = (amp, [factor, ...]); each factor a dict
  {'kind': 'expquad'|'maternp'|'matern'|'cauchy'|'white'|'constant', 'scale': s, 'loc': l, 'dims': [..], 'p':, 'nu':,
   'alpha':, 'beta':}
"""

import numpy as np

from . import iso
from .decomp import Chol


def factor_value(f, x, y):
    kind = f['kind']
    dims = f.get('dims')
    if kind == 'constant':
        return np.ones((x.shape[1], y.shape[1]))
    if kind == 'white':
        xs = x if dims is None else x[list(dims)]
        ys = y if dims is None else y[list(dims)]
        sx = f.get('scale')
        return iso.white(xs if sx is None else xs / sx, ys if sx is None else ys / sx)
    r2 = iso.r2(x, y, scale=f.get('scale'), loc=f.get('loc'), dims=dims)
    if kind == 'expquad':
        return iso.expquad_core(r2)
    if kind == 'maternp':
        return iso.maternp_core(r2, f['p'])
    if kind == 'matern':
        return iso.matern_core(r2, f['nu'])
    if kind == 'cauchy':
        return iso.cauchy_core(r2, f.get('alpha', 2), f.get('beta', 2))
    raise KeyError(kind)


def factor_dlogscale(f, x, y):
    """ d value / d log(scale) = dvalue/dr2 * (-2 r2) """
    kind = f['kind']
    if kind in ('constant', 'white'):
        return np.zeros((x.shape[1], y.shape[1]))
    r2 = iso.r2(x, y, scale=f.get('scale'), loc=f.get('loc'), dims=f.get('dims'))
    if kind == 'expquad':
        d = iso.expquad_dr2(r2)
    elif kind == 'maternp':
        d = iso.maternp_dr2(r2, f['p'])
    elif kind == 'matern':
        d = iso.matern_dr2(r2, f['nu'])
    elif kind == 'cauchy':
        d = iso.cauchy_dr2(r2, f.get('alpha', 2), f.get('beta', 2))
    else:
        raise KeyError(kind)
    return d * (-2 * r2)


def gram(terms, x, y):
    """ sum over terms of amp * prod factors (src/lsqfitgp/_Kernel/_alg.py:48-82) """
    out = None
    for amp, factors in terms:
        val = None
        for i, f in enumerate(factors):
            v = factor_value(f, x, y)
            if i == 0:
                v = amp * v
            val = v if val is None else val * v
        out = val if out is None else out + val
    return out


def gram_chunked(terms, x, y, chunk=1024):
    out = np.empty((x.shape[1], y.shape[1]))
    for s in range(0, x.shape[1], chunk):
        out[s:s + chunk] = gram(terms, x[:, s:s + chunk], y)
    return out


def logml(K, y, ycov=None, **kw):
    """ log marginal likelihood of zero-mean data y with covariance K (+ ycov) """
    if ycov is not None:
        K = K + ycov
    dec = Chol(K, **kw)
    val, _, _, _, _ = dec.minus_log_normal_density(y, value=True)
    return -val


def pred(Kxx, Kxxs, Kxsxs, y, ycov=None, **kw):
    """ posterior mean and covariance, fromdata=True, raw=True (_compute.py:255-260) """
    if ycov is not None:
        Kxx = Kxx + ycov
    solver = Chol(Kxx, **kw)
    mean = solver.pinv_bilinear(Kxxs, y)
    cov = Kxsxs - solver.ginv_quad(Kxxs)
    return mean, cov


def logml_and_grad(terms, x, y, params, timers=None, **kw):
    """ -logML value and gradient w.r.t. a list of hyperparameters, in the reference's formulation
    (L^-1 I, invL' invL, two contractions; _decomp.py:466-472,505-509).
    params: list of ('amp', term_index) | ('logscale', term_index, factor_index): derivative w.r.t. the
    amplitude of a term / the log of the scale of a factor.
    timers: optional dict, filled with the wall-clock seconds of each phase (bench.py's CPU baseline extrapolates every
    phase with its own exponent): 'gram' O(n^2), 'chol' O(n^3), 'solve' O(n^2), 'inverse' O(n^3), 'dgram' O(n^2). """
    import time
    from scipy import linalg

    def lap(name, t0):
        if timers is not None:
            timers[name] = timers.get(name, 0.0) + time.perf_counter() - t0
        return time.perf_counter()
    t = time.perf_counter()
    K = gram_chunked(terms, x, x)
    t = lap('gram', t)
    dec = Chol(K, **kw)
    L = dec._L
    t = lap('chol', t)
    invLr = linalg.solve_triangular(L, y, lower=True)
    invKr = linalg.solve_triangular(L.T, invLr, lower=False)
    t = lap('solve', t)
    invL = linalg.solve_triangular(L, np.eye(len(L)), lower=True)
    invK = invL.T @ invL
    t = lap('inverse', t)
    value = 1 / 2 * (len(L) * np.log(2 * np.pi) + 2 * np.sum(np.log(np.diag(L))) + invLr @ invLr)
    grads = []
    for par in params:
        if par[0] == 'amp':
            amp, factors = terms[par[1]]
            dK = gram([(1.0, factors)], x, x)
        elif par[0] == 'logscale':
            amp, factors = terms[par[1]]
            dK = None
            for i, f in enumerate(factors):
                v = factor_dlogscale(f, x, x) if i == par[2] else factor_value(f, x, x)
                dK = v if dK is None else dK * v
            dK = amp * dK
        else:
            raise KeyError(par)
        tr_invK_dK = np.sum(invK * dK)
        r_invK_dK_invK_r = invKr @ dK @ invKr
        grads.append(1 / 2 * (tr_invK_dK - r_invK_dK_invK_r))
    lap('dgram', t)
    return value, np.array(grads)


def logml_value_lean(terms, x, y, timers=None, chunk=1024, **kw):
    """ -logML value only, with one n x n buffer (Gram written chunk-wise, equilibrated / jittered / factored in place by
    decomp.chol_inplace): the same numbers as `logml`, for sizes where the temporaries of the plain restatement would
    not fit the host (n = 20000 ... 30000 in bench.py). """
    import time
    from scipy import linalg
    from .decomp import chol_inplace
    t0 = time.perf_counter()
    n = x.shape[1]
    K = np.empty((n, n))
    for s in range(0, n, chunk):
        K[s:s + chunk] = gram(terms, x[:, s:s + chunk], x)
    t1 = time.perf_counter()
    L, eps = chol_inplace(K, **kw)
    del K
    t2 = time.perf_counter()
    invLr = linalg.solve_triangular(L, y, lower=True, check_finite=False)
    value = 1 / 2 * (n * np.log(2 * np.pi) + 2 * np.sum(np.log(np.diagonal(L))) + invLr @ invLr)
    t3 = time.perf_counter()
    if timers is not None:
        timers.update(gram=t1 - t0, chol=t2 - t1, solve=t3 - t2)
    return value, L, eps
