"""Generate the golden fixtures in tests/golden/ from the CPU oracle (python tests/golden/make_golden.py).

The reference itself cannot be imported in this image (jax, jaxlib, gvar missing), so these vectors come from the
NumPy/SciPy restatement in oracle/, which is pinned to the reference's own known-answer tests by
tests/test_oracle_*.py.  Each fixture stores inputs (or the seed that makes them) and outputs, small enough to
commit.  The GPU tests compare the CUDA path against them; the CPU tests re-derive them from the oracle.
"""
import pathlib
import sys

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from oracle import gp as ogp, bart as obart  # noqa: E402

OUT = pathlib.Path(__file__).resolve().parent


def c1(n=200, m=50):
    """ config 1 shape (ExpQuad 1-D, data covariance 0.01 I), reduced n """
    rng = np.random.default_rng(1001)
    x = np.sort(rng.uniform(0, 100, n))
    y = np.sin(x / 3) + 0.1 * rng.standard_normal(n)
    xp = np.linspace(-5, 105, m)
    terms = [(1.0, [dict(kind='expquad', scale=3)])]
    Kxx = ogp.gram(terms, x[None], x[None])
    Kxs = ogp.gram(terms, x[None], xp[None])
    Kss = ogp.gram(terms, xp[None], xp[None])
    ycov = 0.01 * np.eye(n)
    mean, cov = ogp.pred(Kxx, Kxs, Kss, y, ycov)
    return dict(x=x, y=y, xpred=xp, gram_row0=Kxx[0], gram_diag1=np.diag(Kxx, 1), logml=ogp.logml(Kxx, y, ycov),
                mean=mean, cov_diag=np.diag(cov), cov_row0=cov[0])


def c2(n=300):
    """ config 2 shape (Matern nu=2.5, 3-D, noise), reduced n: logML and gradient """
    rng = np.random.default_rng(2002)
    X = rng.uniform(0, 10, (n, 3))
    y = np.sin(X[:, 0]) + np.cos(X[:, 1]) * X[:, 2] / 10 + 0.1 * rng.standard_normal(n)
    theta = np.array([np.log(1.5), 0.0, np.log(0.1)])
    terms = [(1.0, [dict(kind='matern', nu=2.5, scale=1.5)]), (0.01, [dict(kind='white')])]
    val, g = ogp.logml_and_grad(terms, X.T.copy(), y, [('logscale', 0, 0), ('amp', 0), ('amp', 1)])
    g = np.array([g[0], g[1] * 2.0, g[2] * 2 * 0.01])
    K = ogp.gram(terms, X.T.copy(), X.T.copy())
    return dict(X=X, y=y, theta=theta, minus_logml=val, grad_minus_logml=g, gram_row0=K[0], gram_col5=K[:, 5])


def c4(n=120, p=10):
    """ config 4 shape (BART maxd=10 reset=[2,4,6,8], 8 continuous + 2 binary covariates), reduced n """
    rng = np.random.default_rng(4004)
    X = np.concatenate([rng.standard_normal((n, 8)), rng.integers(0, 2, (n, 2)).astype(float)], axis=1)
    length, splits = obart.splits_from_coord(X)
    idx = obart.indices_from_coord(X, (length, splits))
    K = obart.gram(length, idx, idx, alpha=0.95, beta=2, maxd=10, reset=[2, 4, 6, 8], gamma=1)
    y = rng.standard_normal(n)
    return dict(X=X, length=length, idx=idx, gram=K, y=y, logml=ogp.logml(K + 0.1 * np.eye(n), y, epsrel=0))


if __name__ == '__main__':
    for name, fn in [('c1_expquad', c1), ('c2_matern', c2), ('c4_bart', c4)]:
        d = fn()
        np.savez_compressed(OUT / f'{name}.npz', **d)
        print(name, {k: np.shape(v) for k, v in d.items()})
