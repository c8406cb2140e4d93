"""Oracle: Chol decomposition (NumPy/SciPy restatement of src/lsqfitgp/_linalg/_decomp.py).

  eigval_bound          :349-354
  diag_scale_pow2       :356-361
  Decomposition._parseeps :245-255
  Chol.__init__         :380-393   (scipy.linalg.cholesky(lower=True) == LAPACK dpotrf, the routine
                                    jax.scipy.linalg.cholesky dispatches to on CPU)
  Chol solves           :395-439
  minus_log_normal_density :441-586 (derivative inputs are explicit arrays/callables instead of jax vjp/jvp)
"""

import numpy as np
from scipy import linalg


def eigval_bound(K):
    return np.max(np.sum(np.abs(K), axis=1))


def diag_scale_pow2(K):
    d = np.diag(K)
    with np.errstate(divide='ignore', invalid='ignore'):
        return np.where(d, np.exp2(np.rint(0.5 * np.log2(d))), 1)


def chol_inplace(K, *, epsrel='auto', epsabs=0):
    """ Chol.__init__ (:380-393) on ONE buffer: K (C-contiguous, exactly symmetric) is equilibrated, jittered and
    factored in place; returns (L, Chol.eps) with L = diag(s) Lt a Fortran-ordered view of the same memory.  Same
    arithmetic as the `Chol` class below (division by powers of two is exact whatever the order; dpotrf on the
    transposed view of a symmetric matrix is the same call), without its four n x n temporaries. """
    n = len(K)
    assert K.flags.c_contiguous and K.shape == (n, n)
    s = diag_scale_pow2(K)
    K /= s
    K /= s[:, None]
    machine_eps = np.finfo(float).eps
    if isinstance(epsrel, str) and epsrel == 'auto':
        epsrel = n * machine_eps
    if isinstance(epsabs, str) and epsabs == 'auto':
        epsabs = machine_eps
    maxeigv = max(np.max(np.sum(np.abs(K[i:i + 1024]), axis=1)) for i in range(0, n, 1024))
    eps = epsrel * maxeigv + epsabs
    K[np.diag_indices_from(K)] += eps
    L = linalg.cholesky(K.T, lower=True, overwrite_a=True, check_finite=False)  # K.T: Fortran-ordered view, no copy
    if not np.all(np.isfinite(np.diagonal(L))):
        raise np.linalg.LinAlgError('cholesky decomposition not finite, probably matrix not pos def numerically')
    L *= s[:, None]
    return L, eps * np.min(s * s)


class Chol:

    def __init__(self, K, *, epsrel='auto', epsabs=0):
        K = np.asarray(K, dtype=float)
        self._K = K
        s = diag_scale_pow2(K)
        K = K / s / s[:, None]
        eps = self._parseeps(K, epsrel, epsabs)
        K = K.copy()
        K[np.diag_indices_from(K)] += eps
        try:
            L = linalg.cholesky(K, lower=True, check_finite=False)
        except linalg.LinAlgError:
            L = np.full_like(K, np.nan)  # jax returns NaNs instead of raising
        if not np.all(np.isfinite(L)):
            raise np.linalg.LinAlgError('cholesky decomposition not finite, probably matrix not pos def numerically')
        self._L = L * s[:, None]
        self._eps = eps * np.min(s * s)

    def _parseeps(self, K, epsrel, epsabs, maxeigv=None):
        machine_eps = np.finfo(float).eps
        if isinstance(epsrel, str) and epsrel == 'auto':
            epsrel = len(K) * machine_eps
        if isinstance(epsabs, str) and epsabs == 'auto':
            epsabs = machine_eps
        if maxeigv is None:
            maxeigv = eigval_bound(K)
        self._eps = epsrel * maxeigv + epsabs
        return self._eps

    @property
    def eps(self):
        return self._eps

    @property
    def n(self):
        return len(self._L)

    m = n

    def matrix(self):
        return self._K

    def ginv(self):
        return self.ginv_quad(np.eye(self.n))

    def ginv_linear(self, X):
        invLX = linalg.solve_triangular(self._L, X, lower=True)
        return linalg.solve_triangular(self._L.T, invLX, lower=False)

    def pinv_bilinear(self, A, r):
        invLr = linalg.solve_triangular(self._L, r, lower=True)
        invLA = linalg.solve_triangular(self._L, A, lower=True)
        return invLA.T @ invLr

    def ginv_quad(self, A):
        invLA = linalg.solve_triangular(self._L, A, lower=True)
        return invLA.T @ invLA

    def ginv_diagquad(self, A):
        invLA = linalg.solve_triangular(self._L, A, lower=True)
        return np.einsum('ji,ji->i', invLA, invLA)

    def correlate(self, x):
        return self._L @ x

    def back_correlate(self, X):
        return self._L.T @ X

    def pinv_correlate(self, x):
        return linalg.solve_triangular(self._L, x, lower=True)

    def minus_log_normal_density(self, r, *, dr_vjp=None, dK_vjp=None, dr_jvp_vec=None, dK_jvp_vec=None, dr=None,
                                 dK=None, value=False, gradrev=False, gradfwd=False, fisher=False, fishvec=False):
        L = self._L
        out = {}
        grad = ((gradrev and (dK_vjp is not None or dr_vjp is not None))
                or (gradfwd and (dK is not None or dr is not None)))
        if value or grad:
            invLr = linalg.solve_triangular(L, r, lower=True)
        if grad:
            invKr = linalg.solve_triangular(L.T, invLr, lower=False)
        if (gradrev and dK_vjp is not None) or (gradfwd and dK is not None):
            invL = linalg.solve_triangular(L, np.eye(len(L)), lower=True)
            invK = invL.T @ invL

        if value:
            out['value'] = 1 / 2 * (len(L) * np.log(2 * np.pi) + 2 * np.sum(np.log(np.diag(L))) + invLr @ invLr)
        else:
            out['value'] = None

        if gradrev:
            out['gradrev'] = 0
            if dK_vjp is not None:
                tr_invK_dK = dK_vjp(invK)
                r_invK_dK_invK_r = dK_vjp(np.outer(invKr, invKr))
                out['gradrev'] += 1 / 2 * (tr_invK_dK - r_invK_dK_invK_r)
            if dr_vjp is not None:
                out['gradrev'] += dr_vjp(invKr)
        else:
            out['gradrev'] = None

        if gradfwd:
            out['gradfwd'] = 0
            if dK is not None:
                tr_invK_dK = np.einsum('ij,ijk->k', invK, dK)
                r_invK_dK_invK_r = np.einsum('i,ijk,j->k', invKr, dK, invKr)
                out['gradfwd'] += 1 / 2 * (tr_invK_dK - r_invK_dK_invK_r)
            if dr is not None:
                out['gradfwd'] += invKr @ dr
        else:
            out['gradfwd'] = None

        if fisher:
            out['fisher'] = 0
            if dK is not None:
                dKk = np.moveaxis(dK, 2, 0)
                invL_dK = np.stack([linalg.solve_triangular(L, m, lower=True) for m in dKk])
                invL_dK_invL = np.stack([linalg.solve_triangular(L, m.T, lower=True) for m in invL_dK])
                out['fisher'] += 1 / 2 * np.einsum('kij,qij->kq', invL_dK_invL, invL_dK_invL)
            if dr is not None:
                invLdr = linalg.solve_triangular(L, dr, lower=True)
                out['fisher'] += invLdr.T @ invLdr
        else:
            out['fisher'] = None

        if fishvec:
            out['fishvec'] = 0
            if not (dK_jvp_vec is None and dK_vjp is None):
                invL_dKv = linalg.solve_triangular(L, dK_jvp_vec, lower=True)
                invK_dKv = linalg.solve_triangular(L.T, invL_dKv, lower=False)
                invL_dKv_invK = linalg.solve_triangular(L, invK_dKv.T, lower=True)
                invK_dKv_invK = linalg.solve_triangular(L.T, invL_dKv_invK, lower=False)
                out['fishvec'] += 1 / 2 * dK_vjp(invK_dKv_invK)
            if not (dr_jvp_vec is None and dr_vjp is None):
                invL_drv = linalg.solve_triangular(L, dr_jvp_vec, lower=True)
                invK_drv = linalg.solve_triangular(L.T, invL_drv, lower=False)
                out['fishvec'] += dr_vjp(invK_drv)
        else:
            out['fishvec'] = None

        return tuple(out.values())
