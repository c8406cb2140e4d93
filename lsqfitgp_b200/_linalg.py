"""Decompositions of positive semidefinite matrices: the `Decomposition` interface and `Chol`.

Mirror of src/lsqfitgp/_linalg/_decomp.py (Decomposition ABC :126-278, Chol :376-658).  The arithmetic
(equilibration, Gershgorin jitter, blocked Cholesky on the FP64 tensor pipe, triangular solves, inverse
from the factor) runs in liblgpb200.so through lsqfitgp_b200._ops.

Array convention: methods accept numpy arrays or torch tensors; torch CUDA inputs give torch CUDA outputs,
anything else gives numpy arrays.
"""

import abc
import math

import numpy
import torch

from . import _lib
from . import _ops

__all__ = ['Decomposition', 'Chol', 'solve_triangular_python']

f64 = torch.float64


def _device():
    _lib.require_cuda()
    return torch.device('cuda', torch.cuda.current_device())


def _todev(x):
    """ -> (device float64 tensor, was_torch_cuda) """
    if isinstance(x, torch.Tensor):
        if x.is_cuda:
            return x.to(f64), True
        return x.detach().to(_device(), f64), False
    return torch.as_tensor(numpy.asarray(x, dtype=numpy.float64)).to(_device()), False


def _out(t, like_torch):
    return t if like_torch else t.cpu().numpy()


class Decomposition(abc.ABC):
    """ Abstract base class for decompositions of positive semidefinite matrices (reference _decomp.py:126-278) """

    @abc.abstractmethod
    def __init__(self, *args, **kw):
        pass

    @abc.abstractmethod
    def matrix(self):
        """ The input matrix """

    @abc.abstractmethod
    def ginv_linear(self, X):
        """ Compute K⁻X """

    @abc.abstractmethod
    def pinv_bilinear(self, A, r):
        """ Compute A'K⁺r """

    @abc.abstractmethod
    def pinv_bilinear_robj(self, A, r):
        """ Compute A'K⁺r, where r can be an array of objects """

    @abc.abstractmethod
    def ginv_quad(self, A):
        """ Compute A'K⁻A """

    @abc.abstractmethod
    def ginv_diagquad(self, A):
        """ Compute diag(A'K⁻A) """

    @abc.abstractmethod
    def correlate(self, x):
        """ Compute Zx where K = ZZ' """

    @abc.abstractmethod
    def back_correlate(self, X):
        """ Compute Z'X """

    @abc.abstractmethod
    def pinv_correlate(self, x):
        """ Compute Z⁺x """

    @abc.abstractmethod
    def minus_log_normal_density(self, r, *, dr_vjp=None, dK_vjp=None, dr_jvp_vec=None, dK_jvp_vec=None, dr=None,
                                 dK=None, value=False, gradrev=False, gradfwd=False, fisher=False, fishvec=False):
        """ minus log Normal density and its derivatives; returns (value, gradrev, gradfwd, fisher, fishvec) """

    @property
    def eps(self):
        """ The threshold below which eigenvalues are too small to be determined """
        return self._eps

    @property
    @abc.abstractmethod
    def n(self):
        """ Number of rows/columns of the matrix """

    @property
    @abc.abstractmethod
    def m(self):
        """ Number of columns of Z """

    def ginv(self):
        """ Compute K⁻ """
        return self.ginv_quad(numpy.eye(self.n))


def solve_triangular_python(a, b, *, lower=False):
    """ pure-python triangular solve for object arrays (reference _decomp.py:280-309) """
    a = numpy.asarray(a)
    x = numpy.copy(b)
    vec = x.ndim < 2
    if vec:
        x = x[:, None]
    n = a.shape[-1]
    assert x.shape[-2] == n
    if not lower:
        a = a[..., ::-1, ::-1]
        x = x[..., ::-1, :]
    x[..., 0, :] /= a[..., 0, 0, None]
    for i in range(1, n):
        x[..., i:, :] -= x[..., None, i - 1, :] * a[..., i:, i - 1, None]
        x[..., i, :] /= a[..., i, i, None]
    if not lower:
        x = x[..., ::-1, :]
    if vec:
        x = numpy.squeeze(x, -1)
    return x


class Chol(Decomposition):
    """Cholesky decomposition, regularised by adding a small multiple of the identity
    (reference _decomp.py:376-393):

        s = 2^rint(log2(diag K)/2);  Kt = K/s/s';  eps = epsrel * max_i sum_j |Kt_ij| + epsabs;
        Kt += eps I;  Lt = chol(Kt);  L = s Lt;  Chol.eps = eps * min(s^2)

    `K` may be a numpy array or a torch tensor.  `_addmat` / `_adddiag` (device tensors) are added to K inside
    the fused equilibration pass (GPCompute._solver `Kxx + ycov`).  Raises numpy.linalg.LinAlgError if the
    factor is not finite, like the reference does outside jit.
    """

    def __init__(self, K, *, epsrel='auto', epsabs=0, _addmat=None, _adddiag=None, _check=True, _inverse_stream=None):
        Kd, self._torch_in = _todev(K)
        if Kd.ndim != 2 or Kd.shape[0] != Kd.shape[1] or Kd.shape[0] < 1:
            raise ValueError(f'matrix must be square and non-empty, found shape {tuple(Kd.shape)}')
        self._K = K
        self._Kd = Kd
        self._addmat = _addmat
        self._adddiag = _adddiag
        self._low = self._low_stream = None
        if _inverse_stream is not None:
            # the caller will want (K + eps)^-1 (gradient): factorisation and inverse in one overlapped library call, the
            # inverse on `_inverse_stream` (lgp_chol_factor_inverse)
            self._st, self._low = _ops.chol_factor_inverse(Kd, _inverse_stream, addmat=_addmat, adddiag=_adddiag,
                                                           epsrel=epsrel, epsabs=epsabs)
            self._low_stream = _inverse_stream
        else:
            self._st = _ops.chol_factor(Kd, addmat=_addmat, adddiag=_adddiag, epsrel=epsrel, epsabs=epsabs)
        self._scal = None
        if _check:
            info = int(self._st.info.item())  # device -> host sync, like the eager isfinite check of the reference
            if info != 0:
                raise numpy.linalg.LinAlgError(
                    'cholesky decomposition not finite, probably matrix not pos def numerically')

    @classmethod
    def _from_kernel(cls, descs, xd, *, epsrel='auto', epsabs=0, _check=True, _inverse_stream=None):
        """ Decomposition of the Gram matrix of `descs` on the points xd WITHOUT materialising it: Gram build fused with
        the equilibration pass (lgp_gram_iso_prepare).  Returns None when the kernel is outside the fused family. """
        out = _ops.gram_chol_factor(descs, xd, side=_inverse_stream, epsrel=epsrel, epsabs=epsabs)
        if out is None:
            return None
        self = object.__new__(cls)
        self._torch_in = True
        self._K = self._Kd = None
        self._addmat = self._adddiag = None
        self._low = self._low_stream = None
        if _inverse_stream is not None:
            self._st, self._low = out
            self._low_stream = _inverse_stream
        else:
            self._st = out
        self._scal = None
        if _check:
            info = int(self._st.info.item())
            if info != 0:
                raise numpy.linalg.LinAlgError(
                    'cholesky decomposition not finite, probably matrix not pos def numerically')
        return self

    # ---- small helpers
    def _scalars(self):
        if self._scal is None:
            self._scal = self._st.scalars().cpu().numpy()
        return self._scal

    @property
    def _eps(self):
        s = self._scalars()
        return float(s[1] * s[3])

    @property
    def n(self):
        return self._st.n

    m = n

    def _rhs(self, X):
        """ -> (2-d device tensor (n, m), was_vector, like_torch) """
        Xd, like = _todev(X)
        vec = Xd.ndim < 2
        if vec:
            Xd = Xd[:, None]
        if Xd.shape[0] != self.n:
            raise ValueError(f'shape mismatch: matrix is {self.n}x{self.n}, right-hand side has {Xd.shape[0]} rows')
        return Xd, vec, like

    def _solve(self, Xd, trans, inplace=False):
        return _ops.chol_solve(self._st, Xd, trans, inplace=inplace)

    @staticmethod
    def _matmul_tn(A, B):
        """ A' B on the device through the DMMA GEMM: A (n, a), B (n, b) -> (a, b) """
        A = _ops.as_aligned(A)
        B = _ops.as_aligned(B)
        n, a = A.shape
        b = B.shape[1]
        C = _ops.aligned_empty(a, b, A.device)
        _ops.dgemm(A, B, C, a_kmajor=False, b_kmajor=False, M=a, N=b, K=n, flags=_lib.GEMM_BETA0)
        return C

    # ---- Decomposition interface
    def matrix(self):
        if self._addmat is None and self._adddiag is None:
            return self._K
        K = self._Kd.clone()
        if self._addmat is not None:
            _ops.axpby(1.0, self._addmat, 1.0, K)
        if self._adddiag is not None:
            K.diagonal().add_(self._adddiag)
        return _out(K, self._torch_in)

    def factor(self):
        """ the lower-triangular factor L (device tensor, zeros above the diagonal) """
        return _ops.chol_get_factor(self._st)

    def ginv_linear(self, X):
        # K⁻¹X = L'⁻¹(L⁻¹X)                                                  (reference _decomp.py:398-403)
        Xd, vec, like = self._rhs(X)
        Y = self._solve(Xd, False)
        Y = self._solve(Y, True, inplace=True)
        return _out(Y[:, 0] if vec else Y, like)

    def pinv_bilinear(self, A, r):
        # A'K⁻¹r = (L⁻¹A)'(L⁻¹r)                                             (reference _decomp.py:405-409)
        Ad, avec, like = self._rhs(A)
        rd, rvec, _ = self._rhs(r)
        invLA = self._solve(Ad, False)
        invLr = self._solve(rd, False)
        out = self._matmul_tn(invLA, invLr)
        if avec and rvec:
            out = out[0, 0]
        elif avec:
            out = out[0]
        elif rvec:
            out = out[:, 0]
        return _out(out, like)

    def pinv_bilinear_robj(self, A, r):
        # r may hold arbitrary objects (gvars in the reference): pure-python forward substitution  (:411-415)
        L = self.factor().cpu().numpy()
        invLr = solve_triangular_python(L, r, lower=True)
        Ad, avec, _ = self._rhs(A)
        invLA = self._solve(Ad, False).cpu().numpy()
        if avec:
            invLA = invLA[:, 0]
        return numpy.asarray(invLA).T @ invLr

    def ginv_quad(self, A):
        # A'K⁻¹A = (L⁻¹A)'(L⁻¹A)                                             (reference _decomp.py:417-420)
        Ad, vec, like = self._rhs(A)
        invLA = self._solve(Ad, False)
        out = self._matmul_tn(invLA, invLA)
        return _out(out[0, 0] if vec else out, like)

    def ginv_diagquad(self, A):
        # diag(A'K⁻¹A) = sum_j (L⁻¹A)_ji^2                                   (reference _decomp.py:422-427)
        Ad, vec, like = self._rhs(A)
        invLA = self._solve(Ad, False)
        out = _ops.colsumsq(invLA)
        return _out(out[0] if vec else out, like)

    def correlate(self, x):
        # Lx                                                                  (reference _decomp.py:429-431)
        xd, vec, like = self._rhs(x)
        y = _ops.chol_mult(self._st, xd, False)
        return _out(y[:, 0] if vec else y, like)

    def back_correlate(self, X):
        # L'X                                                                 (reference _decomp.py:433-435)
        Xd, vec, like = self._rhs(X)
        y = _ops.chol_mult(self._st, Xd, True)
        return _out(y[:, 0] if vec else y, like)

    def pinv_correlate(self, x):
        # L⁻¹x                                                                (reference _decomp.py:437-439)
        xd, vec, like = self._rhs(x)
        y = self._solve(xd, False)
        return _out(y[:, 0] if vec else y, like)

    def inverse_lower(self):
        """ device (n, n) view whose lower triangle holds (K + eps)⁻¹ (TRTRI + LAUUM, 2n³/3 flop) """
        if self._low is not None:
            # computed alongside the factorisation on another stream: order the current stream behind it
            cur = torch.cuda.current_stream()
            if self._low_stream is not None and self._low_stream != cur:
                cur.wait_stream(self._low_stream)
                self._low.record_stream(cur)
            return self._low
        return _ops.chol_inverse(self._st)

    def logdet_quad(self, r=None):
        """ device tensor [sum_i log L_ii, |L⁻¹r|²] """
        a = None
        if r is not None:
            rd, _, _ = self._rhs(r)
            a = self._solve(rd, False)[:, 0].contiguous()
        return _ops.chol_logdet_quad(self._st, a), a

    def minus_log_normal_density(self, r, *, dr_vjp=None, dK_vjp=None, dr_jvp_vec=None, dK_jvp_vec=None, dr=None,
                                 dK=None, value=False, gradrev=False, gradfwd=False, fisher=False, fishvec=False):
        """ reference _decomp.py:441-586; derivative inputs are callables on / arrays of device tensors or numpy """
        n = self.n
        rd, _, like = self._rhs(r)
        out = {}
        grad = ((gradrev and (dK_vjp is not None or dr_vjp is not None))
                or (gradfwd and (dK is not None or dr is not None)))
        invLr = invKr = invK = None
        if value or grad:
            invLr = self._solve(rd, False)
        if grad:
            invKr = self._solve(invLr, True)
        low = None
        if (gradrev and dK_vjp is not None) or (gradfwd and dK is not None):
            low = self.inverse_lower()   # lower triangle of (K + eps)^-1: the contractions below read only that

        def conv(x):
            return x if like else x.cpu().numpy()

        def todev(x):
            return _todev(x)[0]

        if value:
            ldq = _ops.chol_logdet_quad(self._st, invLr[:, 0].contiguous())
            ld, q = (float(v) for v in ldq.cpu().numpy())
            val = 1 / 2 * (n * math.log(2 * math.pi) + 2 * ld + q)
            out['value'] = torch.tensor(val, dtype=f64, device=rd.device) if like else val
        else:
            out['value'] = None

        if gradrev:
            g = 0
            if dK_vjp is not None:
                # the VJP is linear: dK_vjp(invK) - dK_vjp(outer(b, b)) = dK_vjp(invK - outer(b, b)), one call on the
                # full symmetric matrix written by one kernel pass from the lower triangle
                b = invKr[:, 0].contiguous()
                g = g + 1 / 2 * todev(dK_vjp(conv(_ops.sym_expand_sub(low, b))))
            if dr_vjp is not None:
                g = g + todev(dr_vjp(conv(invKr[:, 0])))
            out['gradrev'] = conv(g) if isinstance(g, torch.Tensor) else g
        else:
            out['gradrev'] = None

        if gradfwd:
            g = 0
            if dK is not None:
                # einsum('ij,ijk->k', invK, dK) - einsum('i,ijk,j->k', b, dK, b), one fused pass per k over the lower
                # triangle of invK (lgp_symlower_dot); dK: (n, n, k) array or a sequence of k (n, n) matrices
                if isinstance(dK, (list, tuple)):
                    dlist = [todev(m) for m in dK]
                else:
                    dKd = todev(dK)
                    dlist = [dKd[:, :, q] for q in range(dKd.shape[2])]
                b = invKr[:, 0].contiguous()
                g = g + 1 / 2 * torch.cat([_ops.symlower_dot(low, b, m) for m in dlist])
            if dr is not None:
                g = g + invKr[:, 0] @ todev(dr)
            out['gradfwd'] = conv(g) if isinstance(g, torch.Tensor) else g
        else:
            out['gradfwd'] = None

        if fisher:
            fm = 0
            if dK is not None:
                # B_q = L⁻¹ dK_q L⁻ᵀ (two blocked TRSMs on the DMMA GEMM path), F_kq = 1/2 sum_ij B_k,ij B_q,ij
                # (_decomp.py:547-554).  dK: (n, n, k) array as in the reference, or a sequence of k (n, n) matrices.
                if isinstance(dK, (list, tuple)):
                    dlist = [todev(m) for m in dK]
                else:
                    dKd = todev(dK)
                    dlist = [dKd[:, :, q] for q in range(dKd.shape[2])]
                mats = []
                for m in dlist:
                    t1 = self._solve(_ops.as_aligned(m), False)      # L⁻¹ dK_q
                    t2 = self._solve(_ops.as_aligned(t1.T), False)   # L⁻¹ (L⁻¹ dK_q)'
                    del t1
                    mats.append(t2)
                k = len(mats)
                fm = torch.zeros(k, k, dtype=f64, device=rd.device)
                for a in range(k):
                    for b in range(a + 1):
                        fm[a, b] = fm[b, a] = 0.5 * _ops.frob_dot(mats[a], mats[b])[0]
                del mats
            if dr is not None:
                invLdr = self._solve(todev(dr), False)
                fm = fm + self._matmul_tn(invLdr, invLdr)
            out['fisher'] = conv(fm) if isinstance(fm, torch.Tensor) else fm
        else:
            out['fisher'] = None

        if fishvec:
            fv = 0
            if not (dK_jvp_vec is None and dK_vjp is None):
                dKv = todev(dK_jvp_vec)
                t = self._solve(dKv, False)
                t = self._solve(t, True, inplace=True)        # K⁻¹ dKv
                t = self._solve(t.T.contiguous(), False)
                t = self._solve(t, True, inplace=True)        # K⁻¹ dKv K⁻¹
                fv = fv + 1 / 2 * todev(dK_vjp(conv(t)))
            if not (dr_jvp_vec is None and dr_vjp is None):
                t = self._solve(todev(dr_jvp_vec)[:, None], False)
                t = self._solve(t, True, inplace=True)
                fv = fv + todev(dr_vjp(conv(t[:, 0])))
            out['fishvec'] = conv(fv) if isinstance(fv, torch.Tensor) else fv
        else:
            out['fishvec'] = None

        return tuple(out.values())
