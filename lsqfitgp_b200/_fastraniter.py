"""Random samples from a multivariate Gaussian given mean and covariance separately.

Mirror of src/lsqfitgp/_fastraniter.py:36-121 (`raniter`, `sample`): the covariance is decomposed with `Chol`
(device: equilibrated, jittered Cholesky on the DMMA pipe) and every sample is `mean + L z` through `Chol.correlate`
(lgp_chol_mult).  The standard normal variates come from numpy's generator on the host exactly as in the reference, so the
same seed gives the same samples; `sample_batch` draws many samples with ONE product L Z (the case the reference's docs
name as the bottleneck when there are many more prediction points than data, docs/userguide/optim.rst:47-53), optionally
with variates generated on the device.
"""

import itertools

import numpy

from . import _linalg

__all__ = ['raniter', 'sample', 'sample_batch']


def _flatten(mean, cov):
    """ -> (flat mean, square covariance, unflatten function) """
    if hasattr(mean, 'keys'):
        keys = list(mean.keys())
        shapes = [numpy.shape(mean[k]) for k in keys]
        sizes = [int(numpy.prod(s, dtype=int)) for s in shapes]
        offs = numpy.cumsum([0] + sizes)
        flatmean = numpy.concatenate([numpy.asarray(mean[k], dtype=float).reshape(-1) for k in keys])
        squarecov = numpy.empty((len(flatmean), len(flatmean)))
        for i, k1 in enumerate(keys):
            for j, k2 in enumerate(keys):
                squarecov[offs[i]:offs[i + 1], offs[j]:offs[j + 1]] = numpy.asarray(cov[k1, k2], dtype=float).reshape(
                    sizes[i], sizes[j])

        def unflat(buf):
            return {k: buf[..., offs[i]:offs[i + 1]].reshape(buf.shape[:-1] + shapes[i]) for i, k in enumerate(keys)}
        return flatmean, squarecov, unflat
    mean = numpy.asarray(mean, dtype=float)
    cov = numpy.asarray(cov, dtype=float)
    flatmean = mean.reshape(-1)
    squarecov = cov.reshape(len(flatmean), len(flatmean))

    def unflat(buf):
        out = buf.reshape(buf.shape[:-1] + mean.shape)
        return out if (mean.shape or buf.ndim > 1) else out.item()
    return flatmean, squarecov, unflat


def _decompose(squarecov, eps):
    try:
        return _linalg.Chol(squarecov, epsrel='auto' if eps is None else eps)
    except numpy.linalg.LinAlgError:
        raise numpy.linalg.LinAlgError('covariance matrix not positive definite with eps={}'.format(eps))


def raniter(mean, cov, n=None, eps=None, rng=None):
    """Generator of random samples from N(mean, cov); `mean` scalar, array or dictionary of arrays, `cov` array or
    dictionary keyed by pairs of keys; `n` maximum number of iterations; `eps` relative jitter of the decomposition
    (default: number of variables times machine epsilon); `rng` seed or numpy generator (reference :36-116)."""
    flatmean, squarecov, unflat = _flatten(mean, cov)
    covdec = _decompose(squarecov, eps)
    rng = numpy.random.default_rng(rng)
    iterable = itertools.count() if n is None else range(n)
    for _ in iterable:
        iidsamp = rng.standard_normal(covdec.m)
        yield unflat(flatmean + covdec.correlate(iidsamp))


def sample(*args, **kw):
    """ Shortcut for ``next(raniter(..., n=1))`` (reference :117-121) """
    return next(raniter(*args, n=1, **kw))


def sample_batch(mean, cov, nsamples, eps=None, rng=None, device_rng=False):
    """ `nsamples` samples at once, leading axis = sample index: one factorisation and ONE triangular product L Z on the
    device (Z: (m, nsamples)).  device_rng=False draws Z with numpy's generator in the order `raniter` would (sample s =
    the s-th item of raniter with the same seed); device_rng=True draws it on the GPU with torch's generator seeded
    from `rng` (no host->device copy of Z). """
    import torch
    flatmean, squarecov, unflat = _flatten(mean, cov)
    covdec = _decompose(squarecov, eps)
    m = covdec.m
    if device_rng:
        dev = torch.device('cuda', torch.cuda.current_device())
        g = torch.Generator(device=dev)
        g.manual_seed(int(numpy.random.default_rng(rng).integers(0, 2 ** 62)))
        Z = torch.randn((m, nsamples), dtype=torch.float64, device=dev, generator=g)
        out = covdec.correlate(Z).T.cpu().numpy()
    else:
        rng = numpy.random.default_rng(rng)
        Z = numpy.stack([rng.standard_normal(m) for _ in range(nsamples)], axis=1)
        out = covdec.correlate(Z).T
    return unflat(flatmean + out)
