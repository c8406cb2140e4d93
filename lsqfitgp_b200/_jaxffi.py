"""JAX side of the XLA-FFI binding (boundary B3 of SURVEY.md section 8b).

STATUS: UNTESTED -- jax / jaxlib cannot be installed in the build image, so nothing in this module has ever run; importing
it raises ImportError without jax.  It is the counterpart of csrc/xla_ffi_shim.cc (built by `make -C lsqfitgp_b200/csrc
ffi` on a box with jax): registration of the handlers, `ffi_call` wrappers with the output shapes the C ABI prescribes
(include/lgp_b200.h), and the `jax.custom_vjp` rules that keep the hyperparameters differentiable under tracing:

  gram_iso(structure, devpar, x, y)       K = kernel(x, y); VJP w.r.t. devpar through lgp_gram_iso_vjp_dev
  chol_factor / chol_solve / chol_inverse / chol_logdet_quad
  neg_log_density(K, r)                   1/2 (n log 2pi + log det K + r' K^-1 r) with the closed-form reverse rule
                                          dK = 1/2 (K^-1 - b b'), dr = b (src/lsqfitgp/_linalg/_decomp.py:505-512)

`structure` is the static part of a kernel descriptor (kind, term, dimmask, ipar, par0: numpy int32 / float64 arrays,
FFI attributes); `devpar` the traced (nfactors, 6) array of scale_x, scale_y, loc_x, loc_y, par1, amp.  What the tested
torch path does with host descriptors (lsqfitgp_b200/_GP.py: _GramFn, _FusedNegLogMLFn) this module does with tracers.
"""

import ctypes
import math
import pathlib

import numpy

import jax  # noqa: E402  (ImportError here is the documented behaviour without jax)
import jax.numpy as jnp

_SO = pathlib.Path(__file__).resolve().parent / 'csrc' / 'liblgpb200_xla.so'
_TARGETS = ('lgp_xla_gram_iso', 'lgp_xla_gram_iso_vjp', 'lgp_xla_gram_iso_jvp', 'lgp_xla_chol_factor',
            'lgp_xla_chol_solve', 'lgp_xla_chol_inverse', 'lgp_xla_chol_logdet_quad')
_registered = False


def register():
    """ load liblgpb200_xla.so and register its handlers for the CUDA platform (idempotent) """
    global _registered
    if _registered:
        return
    so = ctypes.CDLL(str(_SO))
    for name in _TARGETS:
        jax.ffi.register_ffi_target(name, jax.ffi.pycapsule(getattr(so, name)), platform='CUDA')
    _registered = True


def _npad(n):
    return -(-n // 128) * 128


def _attrs(structure):
    kind, term, dimmask, ipar, par0 = structure
    return dict(kind=numpy.asarray(kind, numpy.int32), term=numpy.asarray(term, numpy.int32),
                dimmask=numpy.asarray(dimmask, numpy.int32), ipar=numpy.asarray(ipar, numpy.int32),
                par0=numpy.asarray(par0, numpy.float64))


def _gram_call(structure, devpar, x, y):
    register()
    out = jax.ShapeDtypeStruct((x.shape[1], y.shape[1]), jnp.float64)
    return jax.ffi.ffi_call('lgp_xla_gram_iso', out)(devpar, x, y, **_attrs(structure))


def gram_iso(structure, devpar, x, y):
    """ Gram matrix (n, m); differentiable w.r.t. devpar (amp: column 5; scales: columns 0-1 through d/d log scale;
    par1: column 4).  x: (ndim, n), y: (ndim, m) float64. """
    structure = tuple(tuple(numpy.asarray(a).tolist()) for a in structure)   # hashable static argument

    @jax.custom_vjp
    def f(devpar, x, y):
        return _gram_call(structure, devpar, x, y)

    def fwd(devpar, x, y):
        return _gram_call(structure, devpar, x, y), (devpar, x, y)

    def bwd(res, G):
        devpar, x, y = res
        nf = devpar.shape[0]
        out = jax.ShapeDtypeStruct((3 * nf + 8,), jnp.float64)
        v = jax.ffi.ffi_call('lgp_xla_gram_iso_vjp', out)(devpar, x, y, G, **_attrs(structure))[:3 * nf].reshape(nf, 3)
        # v[:, 0] = d/d amp, v[:, 1] = d/d log(scale) (scale_x == scale_y), v[:, 2] = d/d par1
        g = jnp.zeros_like(devpar)
        g = g.at[:, 5].set(v[:, 0])
        g = g.at[:, 0].set(v[:, 1] / devpar[:, 0])     # the caller ties scale_y to scale_x; the whole derivative goes to x
        g = g.at[:, 4].set(v[:, 2])
        return g, jnp.zeros_like(x), jnp.zeros_like(y)
    f.defvjp(fwd, bwd)
    return f(devpar, x, y)


def chol_factor(K, epsrel='auto', epsabs=0.0):
    """ (W, aux, info) of lgp_chol_factor: Chol.__init__, src/lsqfitgp/_linalg/_decomp.py:380-393 """
    register()
    n = K.shape[0]
    npad = _npad(n)
    outs = (jax.ShapeDtypeStruct((npad, npad), jnp.float64),
            jax.ShapeDtypeStruct((3 * npad + 16 + npad * 128,), jnp.float64),
            jax.ShapeDtypeStruct((), jnp.int32))
    return jax.ffi.ffi_call('lgp_xla_chol_factor', outs)(K, epsrel=-1.0 if epsrel == 'auto' else float(epsrel),
                                                         epsabs=float(epsabs))


def chol_solve(W, aux, B, trans):
    """ L^-1 B (trans=0) or L^-T B (trans=1); B (n, m) with m even (pad a vector to two columns) """
    register()
    return jax.ffi.ffi_call('lgp_xla_chol_solve', jax.ShapeDtypeStruct(B.shape, jnp.float64))(W, aux, B,
                                                                                             trans=numpy.int32(trans))


def chol_inverse(W, aux, n):
    """ lower triangle of (L L^T)^-1 in an (npad, npad) buffer """
    register()
    npad = W.shape[0]
    outs = (jax.ShapeDtypeStruct((npad, npad), jnp.float64), jax.ShapeDtypeStruct((npad, npad), jnp.float64))
    return jax.ffi.ffi_call('lgp_xla_chol_inverse', outs)(W, aux, n=numpy.int64(n))[0]


def chol_logdet_quad(aux, a):
    register()
    return jax.ffi.ffi_call('lgp_xla_chol_logdet_quad', jax.ShapeDtypeStruct((2,), jnp.float64))(aux, a)


@jax.custom_vjp
def neg_log_density(K, r):
    """ value of Chol(K).minus_log_normal_density(r) (_decomp.py:484-488) """
    return _nld_fwd(K, r)[0]


def _nld_fwd(K, r):
    n = K.shape[0]
    W, aux, info = chol_factor(K)
    a = chol_solve(W, aux, jnp.stack([r, jnp.zeros_like(r)], axis=1), 0)
    ldq = chol_logdet_quad(aux, a[:, 0])
    value = 0.5 * (n * math.log(2 * math.pi) + 2 * ldq[0] + ldq[1])
    return value, (W, aux, a, n)


def _nld_bwd(res, g):
    W, aux, a, n = res
    b = chol_solve(W, aux, a, 1)[:, 0]                         # K^-1 r
    low = chol_inverse(W, aux, n)[:n, :n]
    invK = jnp.tril(low) + jnp.tril(low, -1).T
    return 0.5 * g * (invK - jnp.outer(b, b)), g * b           # _decomp.py:505-512


neg_log_density.defvjp(_nld_fwd, _nld_bwd)
