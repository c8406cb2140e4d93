"""bayestree: GP regression with the BART kernel, hyperparameters fitted by marginal MAP.

Mirror of `lsqfitgp.bayestree.bart` (src/lsqfitgp/bayestree/_bart.py:37-404) on top of the B200 path: the tree
splitting grid and the bin indices (host, `BART.splits_from_coord` / `indices_from_coord`), the GP factory

    K_train = (k_sigma_mu / k)^2 BART(alpha, beta, maxd=10, reset=[2, 4, 6, 8]) + diag(sigma2 / w) + k_sigma_mu^2

(:187-205), the copula hyperprior on alpha ~ Beta(2, 1), beta ~ InvGamma(1, 1) and Normal priors on log k,
log sigma2 (:175-184), the fit options `mlkw=dict(epsrel=0)`, l-bfgs-b (:218-227), and the posterior prediction
(:388-404).  Without gvar the fitted hyperparameters are (mean, sdev) pairs and `pred` returns matrices only.
The Gram matrix, its alpha / beta derivatives (lgp_gram_bart_vjp), the Cholesky factorisation and the solves run in
liblgpb200.so.
"""

import numpy
import torch

from . import _array
from ._GP import GP
from ._fit import empbayes_fit
from ._kernels import BART

__all__ = ['bart']


class bart:
    """Nonparametric Bayesian regression with a GP version of BART (reference bayestree/_bart.py:37-146).

    Parameters: x_train (n, p) array / structured array, y_train (n,), weights (n,) (error variance sigma2 / weight),
    fitkw (overrides the arguments of empbayes_fit), kernelkw (overrides the arguments of BART), marginalize_mean.
    Attributes: mean, sigma, alpha, beta, meansdev ((value, sdev) pairs where fitted), fit (the empbayes_fit object).
    Methods: gp(), data(), pred()."""

    def __init__(self, x_train, y_train, *, weights=None, fitkw={}, kernelkw={}, marginalize_mean=True):
        x_train = self._to_structured(x_train)
        if hasattr(y_train, 'to_numpy'):
            y_train = y_train.to_numpy().squeeze()
        y_train = numpy.asarray(y_train, dtype=float)
        assert y_train.shape == x_train.shape

        self._no_weights = weights is None
        if self._no_weights:
            weights = numpy.ones_like(y_train)
        weights = numpy.asarray(weights, dtype=float)
        assert weights.shape == y_train.shape

        # prior mean and variance
        ymin, ymax = float(numpy.min(y_train)), float(numpy.max(y_train))
        mu_mu = (ymax + ymin) / 2
        k_sigma_mu = (ymax - ymin) / 2

        # splitting points and indices
        splits = BART.splits_from_coord(x_train)
        i_train = self._toindices(x_train, splits)

        # prior on hyperparams: the keys copula.makedict gives the reference (:175-184)
        sigma2_priormean = float(numpy.mean((y_train - y_train.mean()) ** 2 * weights))
        hyperprior = {
            '__copula_beta{2, 1}(alpha)': (0.0, 1.0),       # base of tree gen prob
            '__copula_invgamma{1, 1}(beta)': (0.0, 1.0),    # exponent of tree gen prob
            'log(k)': (numpy.log(2), 2.0),                  # denominator of prior sdev
            'log(sigma2)': (numpy.log(sigma2_priormean), 2.0),  # i.i.d. error variance, scaled with weights
            'mean': (mu_mu, k_sigma_mu),                    # mean of the GP
        }
        if marginalize_mean:
            hyperprior.pop('mean')

        y_t = torch.as_tensor(y_train, dtype=torch.float64)
        w_t = torch.as_tensor(weights, dtype=torch.float64)

        def makegp(hp, *, i_train, weights, splits, **_):
            kw = dict(alpha=hp['alpha'], beta=hp['beta'], maxd=10, reset=[2, 4, 6, 8])
            kw.update(kernelkw)
            kernel = BART(splits=splits, indices=True, **kw)
            kernel = kernel * (k_sigma_mu / hp['k']) ** 2
            # the diagonal noise block is built on the device (a host n x n matrix would cross PCIe at every evaluation)
            dev = torch.device('cuda', torch.cuda.current_device())
            noise = (torch.as_tensor(hp['sigma2'], dtype=torch.float64) / torch.as_tensor(weights, dtype=torch.float64)).to(dev)
            gp = (GP(kernel, checkpos=False, checksym=False, checkfinite=False, solver='chol')
                  .addx(i_train, 'trainmean')
                  .addcov(torch.diag(noise), 'trainnoise'))
            pieces = {'trainmean': 1, 'trainnoise': 1}
            if 'mean' not in hp:
                gp = gp.addcov(k_sigma_mu ** 2, 'mean')
                pieces.update({'mean': 1})
            return gp.addtransf(pieces, 'train')

        def info(hp, *, mu_mu, **_):
            m = hp.get('mean', mu_mu) if hasattr(hp, 'get') else mu_mu
            return {'train': y_t - m}

        gpkw = dict(i_train=i_train, weights=w_t, splits=splits, mu_mu=mu_mu)
        options = dict(
            verbosity=0,
            raises=False,
            minkw=dict(method='l-bfgs-b', options=dict(maxls=4, maxiter=100)),
            mlkw=dict(epsrel=0),
            forward=True,
            gpfactorykw=gpkw,
        )
        options.update(fitkw)
        fit = empbayes_fit(hyperprior, makegp, info, **options)

        # extract hyperparameters from the minimization result: value at the MAP and first-order propagated sdev
        self.fit = fit
        self._k_sigma_mu, self._mu_mu = k_sigma_mu, mu_mu
        hpmap = fit.hp_at(fit.minresult.x)
        sd = fit.hp_sdev(fit.minresult.x, ['sigma2', 'alpha', 'beta', 'k'] + ([] if marginalize_mean else ['mean']))
        self.sigma = (float(hpmap['sigma2']) ** 0.5, sd['sigma2'] / (2 * float(hpmap['sigma2']) ** 0.5))
        self.alpha = (float(hpmap['alpha']), sd['alpha'])
        self.beta = (float(hpmap['beta']), sd['beta'])
        self.meansdev = (k_sigma_mu / float(hpmap['k']), k_sigma_mu / float(hpmap['k']) ** 2 * sd['k'])
        self.mean = (float(hpmap['mean']), sd['mean']) if not marginalize_mean else mu_mu
        self._ystd = float(y_train.std())

    def _gethp(self, hp, rng):
        if not isinstance(hp, str):
            return hp
        if hp == 'map':
            return self.fit.hp_at(self.fit.minresult.x)
        if hp == 'sample':
            return self.fit.hp_sample(rng)
        raise KeyError(hp)

    def gp(self, *, hp='map', x_test=None, weights=None, rng=None):
        """ GP object with the fitted hyperparameters; keys 'Xmean', 'Xnoise', 'X' = Xmean + Xnoise for X in
        {train, test} (reference :240-278) """
        hp = self._gethp(hp, rng)
        return self._gp(hp, x_test, weights, self.fit.gpfactorykw)

    def _gp(self, hp, x_test, weights, gpfactorykw):
        with torch.no_grad():
            gp = self.fit.gpfactory(hp, **gpfactorykw)
            if x_test is not None:
                x_test = self._to_structured(x_test)
                i_test = self._toindices(x_test, gpfactorykw['splits'])
                if weights is not None:
                    weights = numpy.asarray(weights, dtype=float)
                    assert weights.shape == i_test.shape
                else:
                    weights = numpy.ones(i_test.shape)
                s2 = float(hp['sigma2'])
                gp = gp.addx(i_test, 'testmean').addcov(numpy.diag(s2 / weights), 'testnoise')
                pieces = {'testmean': 1, 'testnoise': 1}
                if 'mean' not in hp:
                    pieces.update({'mean': 1})
                gp = gp.addtransf(pieces, 'test')
        return gp

    def data(self, *, hp='map', rng=None):
        """ the dictionary representing y_train to be passed to GP.pred (reference :311-333) """
        hp = self._gethp(hp, rng)
        return self.fit.data(hp, **self.fit.gpfactorykw)

    def pred(self, *, hp='map', error=False, format='matrices', x_test=None, weights=None, rng=None):
        """ posterior mean and covariance of the regression function (error=False) or of new outcomes (error=True) at
        x_test (default: the training covariates) (reference :335-404) """
        if format != 'matrices':
            raise NotImplementedError("format='gvar' needs gvar")
        hp = self._gethp(hp, rng)
        gp = self._gp(hp, x_test, weights, self.fit.gpfactorykw)
        with torch.no_grad():
            data = self.fit.data(hp, **self.fit.gpfactorykw)
            label = 'train' if x_test is None else 'test'
            if not error:
                label += 'mean'
            outmean, outcov = gp.predfromdata(data, label, raw=True)
        m = hp.get('mean', self._mu_mu) if hasattr(hp, 'get') else self._mu_mu
        return outmean + float(m), outcov

    @classmethod
    def _to_structured(cls, x):
        if hasattr(x, 'columns'):
            x = _array.StructuredArray.from_dataframe(x)
        elif isinstance(x, _array.StructuredArray):
            pass
        elif numpy.asarray(x).dtype.names is None:
            x = _array.unstructured_to_structured(numpy.asarray(x))
        else:
            x = _array.StructuredArray(x)
        assert x.ndim == 1
        return x

    @staticmethod
    def _toindices(x, splits):
        ix = BART.indices_from_coord(x, splits)
        return _array.unstructured_to_structured(ix.astype(numpy.int32), names=list(x.dtype.names))

    def __repr__(self):
        fmt = lambda v: f'{v[0]:.3g} +/- {v[1]:.2g}' if isinstance(v, tuple) else f'{v:.3g}'
        out = f"""BART fit:
alpha = {fmt(self.alpha)} (0 -> intercept only, 1 -> any)
beta = {fmt(self.beta)} (0 -> any, inf -> no interactions)
mean = {fmt(self.mean)}
latent sdev = {fmt(self.meansdev)} (large -> conservative extrapolation)
data total sdev = {self._ystd:.3g}"""
        if self._no_weights:
            out += f"\nerror sdev = {fmt(self.sigma)}"
        else:
            w = numpy.asarray(self.fit.gpfactorykw['weights'])
            avg = numpy.sqrt(numpy.mean(self.sigma[0] ** 2 / w))
            out += f"\nerror sdev (avg weighted) = {avg:.3g}\nerror sdev (unweighted) = {fmt(self.sigma)}"
        return out
