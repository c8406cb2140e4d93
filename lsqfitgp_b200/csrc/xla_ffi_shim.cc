// XLA-FFI handlers over the C ABI of liblgpb200.so (boundary B3 of SURVEY.md section 8b: the JAX primitive level at which
// the reference calls jax.scipy.linalg.cholesky / solve_triangular, src/lsqfitgp/_linalg/_decomp.py:388,402-403, and
// evaluates kernels under jax.jit, src/lsqfitgp/_GP/_elements.py:554-579).
//
// STATUS: UNTESTED.  jax / jaxlib are not installable in the build image (no wheels, no network), so this file has never
// been compiled: it is written against the public XLA FFI C++ API (xla/ffi/api/ffi.h, jax >= 0.4.31) and is built only by
// `make ffi`, which needs `python -c "import jax.ffi; print(jax.ffi.include_dir())"` to succeed.  Everything it calls IS
// tested: each handler is a direct forwarding of one C-ABI entry point (tests/test_gpu_kernels.py,
// tests/test_gpu_api.py), and the device-resident-hyperparameter entry points exist for exactly this file
// (lgp_gram_iso_dev / _vjp_dev / _jvp_dev: traced scalars are device buffers and are never read on the host).
//
// Conventions honoured: buffers are XLA-owned device pointers valid for the call only; outputs are
// pre-allocated by XLA; the stream comes from PlatformStream; errors are returned as ffi::Error, numerical failure of the
// factorisation is the device-side `info` word (the analogue of the NaN propagation of the traced reference,
// _decomp.py:389-391); no global mutable state, no host synchronisation.
//
// The Python side (registration, ffi_call wrappers, jax.custom_vjp rules) is lsqfitgp_b200/_jaxffi.py.
#include <cuda_runtime.h>

#include <cstdint>
#include <vector>

#include "../../include/lgp_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

namespace {

ffi::Error Check(int rc, const char *what) {
    if (rc == LGP_OK) return ffi::Error::Success();
    return ffi::Error(rc == LGP_ERR_CUDA ? ffi::ErrorCode::kInternal : ffi::ErrorCode::kInvalidArgument, what);
}

// structural part of the kernel descriptor: static under tracing, travels as FFI attributes
bool BuildFactors(ffi::Span<const int32_t> kind, ffi::Span<const int32_t> term, ffi::Span<const int32_t> dimmask,
                  ffi::Span<const int32_t> ipar, ffi::Span<const double> par0, std::vector<lgp_factor_t> &out) {
    const size_t nf = kind.size();
    if (nf < 1 || nf > LGP_MAX_FACTORS || term.size() != nf || dimmask.size() != nf || ipar.size() != nf ||
        par0.size() != nf)
        return false;
    out.assign(nf, lgp_factor_t{});
    for (size_t i = 0; i < nf; i++) {
        out[i].kind = kind[i];
        out[i].term = term[i];
        out[i].dimmask = (uint32_t)dimmask[i];
        out[i].ipar = ipar[i];
        out[i].par0 = par0[i];
    }
    return true;
}

// ---- Gram matrix: K (n, m) from points x (ndim, n), y (ndim, m) and the device-resident hyperparameters (nf, 6)
ffi::Error GramIsoImpl(cudaStream_t stream, ffi::Buffer<ffi::F64> devpar, ffi::Buffer<ffi::F64> x,
                       ffi::Buffer<ffi::F64> y, ffi::ResultBuffer<ffi::F64> K, ffi::Span<const int32_t> kind,
                       ffi::Span<const int32_t> term, ffi::Span<const int32_t> dimmask, ffi::Span<const int32_t> ipar,
                       ffi::Span<const double> par0) {
    std::vector<lgp_factor_t> f;
    if (!BuildFactors(kind, term, dimmask, ipar, par0, f)) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "descriptor");
    const auto xd = x.dimensions(), yd = y.dimensions();
    if (xd.size() != 2 || yd.size() != 2 || xd[0] != yd[0]) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "x, y");
    const int64_t ndim = xd[0], n = xd[1], m = yd[1];
    return Check(lgp_gram_iso_dev(stream, f.data(), (int)f.size(), (int)ndim, devpar.typed_data(), x.typed_data(), n, n,
                                  y.typed_data(), m, m, K->typed_data(), m, 0),
                 "lgp_gram_iso_dev");
}

// ---- reverse mode of the Gram build: out (nf, 3) = sum_ij G_ij dK_ij / d(amp, log scale, par1)
ffi::Error GramIsoVjpImpl(cudaStream_t stream, ffi::Buffer<ffi::F64> devpar, ffi::Buffer<ffi::F64> x,
                          ffi::Buffer<ffi::F64> y, ffi::Buffer<ffi::F64> G, ffi::ResultBuffer<ffi::F64> out,
                          ffi::Span<const int32_t> kind, ffi::Span<const int32_t> term,
                          ffi::Span<const int32_t> dimmask, ffi::Span<const int32_t> ipar,
                          ffi::Span<const double> par0) {
    std::vector<lgp_factor_t> f;
    if (!BuildFactors(kind, term, dimmask, ipar, par0, f)) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "descriptor");
    const auto xd = x.dimensions(), yd = y.dimensions();
    const int64_t ndim = xd[0], n = xd[1], m = yd[1];
    // `out` is declared with 3*nf + 8 doubles on the Python side (scratch tail of the C ABI)
    return Check(lgp_gram_iso_vjp_dev(stream, f.data(), (int)f.size(), (int)ndim, devpar.typed_data(), x.typed_data(), n,
                                      n, y.typed_data(), m, m, G.typed_data(), m, nullptr, 0, out->typed_data()),
                 "lgp_gram_iso_vjp_dev");
}

// ---- forward mode of the Gram build along one tangent (nf, 3)
ffi::Error GramIsoJvpImpl(cudaStream_t stream, ffi::Buffer<ffi::F64> devpar, ffi::Buffer<ffi::F64> x,
                          ffi::Buffer<ffi::F64> y, ffi::Buffer<ffi::F64> tangent, ffi::ResultBuffer<ffi::F64> D,
                          ffi::Span<const int32_t> kind, ffi::Span<const int32_t> term,
                          ffi::Span<const int32_t> dimmask, ffi::Span<const int32_t> ipar,
                          ffi::Span<const double> par0) {
    std::vector<lgp_factor_t> f;
    if (!BuildFactors(kind, term, dimmask, ipar, par0, f)) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "descriptor");
    const auto xd = x.dimensions(), yd = y.dimensions();
    const int64_t ndim = xd[0], n = xd[1], m = yd[1];
    return Check(lgp_gram_iso_jvp_dev(stream, f.data(), (int)f.size(), (int)ndim, devpar.typed_data(), x.typed_data(), n,
                                      n, y.typed_data(), m, m, tangent.typed_data(), D->typed_data(), m),
                 "lgp_gram_iso_jvp_dev");
}

// ---- Chol.__init__ (_decomp.py:380-393): W (npad, npad), aux, info from K (n, n)
ffi::Error CholFactorImpl(cudaStream_t stream, ffi::Buffer<ffi::F64> K, ffi::ResultBuffer<ffi::F64> W,
                          ffi::ResultBuffer<ffi::F64> aux, ffi::ResultBuffer<ffi::S32> info, double epsrel,
                          double epsabs) {
    const auto kd = K.dimensions();
    if (kd.size() != 2 || kd[0] != kd[1]) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "K must be square");
    const int64_t n = kd[0], npad = lgp_chol_npad(n);
    return Check(lgp_chol_factor(stream, K.typed_data(), n, nullptr, 0, nullptr, n, epsrel, epsabs, W->typed_data(), npad,
                                 aux->typed_data(), info->typed_data()),
                 "lgp_chol_factor");
}

// ---- triangular solves (_decomp.py:402-403,407-409,419-420,439): X = L^-1 B or L^-T B, B (n, m) with m even
ffi::Error CholSolveImpl(cudaStream_t stream, ffi::Buffer<ffi::F64> W, ffi::Buffer<ffi::F64> aux,
                         ffi::Buffer<ffi::F64> B, ffi::ResultBuffer<ffi::F64> X, int32_t trans) {
    const auto bd = B.dimensions();
    if (bd.size() != 2) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "B must be 2-D");
    const int64_t n = bd[0], m = bd[1], npad = W.dimensions()[0];
    if (cudaMemcpyAsync(X->typed_data(), B.typed_data(), sizeof(double) * n * m, cudaMemcpyDeviceToDevice, stream) !=
        cudaSuccess)
        return ffi::Error(ffi::ErrorCode::kInternal, "copy");
    return Check(lgp_chol_solve(stream, W.typed_data(), npad, aux.typed_data(), n, X->typed_data(), m, m, trans),
                 "lgp_chol_solve");
}

// ---- inverse from the factor (_decomp.py:471-472): lower triangle of (L L^T)^-1, (npad, npad); scratch (npad, npad)
ffi::Error CholInverseImpl(cudaStream_t stream, ffi::Buffer<ffi::F64> W, ffi::Buffer<ffi::F64> aux,
                           ffi::ResultBuffer<ffi::F64> Kinv, ffi::ResultBuffer<ffi::F64> scratch, int64_t n) {
    const int64_t npad = W.dimensions()[0];
    return Check(lgp_chol_inverse(stream, W.typed_data(), npad, aux.typed_data(), n, scratch->typed_data(),
                                  Kinv->typed_data(), npad),
                 "lgp_chol_inverse");
}

// ---- reductions of the normal density (_decomp.py:484-488): out = [sum log L_ii, |a|^2]
ffi::Error CholLogdetQuadImpl(cudaStream_t stream, ffi::Buffer<ffi::F64> aux, ffi::Buffer<ffi::F64> a,
                              ffi::ResultBuffer<ffi::F64> out) {
    const int64_t n = a.dimensions()[0];
    return Check(lgp_chol_logdet_quad(stream, aux.typed_data(), n, a.typed_data(), out->typed_data()),
                 "lgp_chol_logdet_quad");
}

}  // namespace

#define LGP_DESC_ATTRS()                              \
    .Attr<ffi::Span<const int32_t>>("kind")           \
        .Attr<ffi::Span<const int32_t>>("term")       \
        .Attr<ffi::Span<const int32_t>>("dimmask")    \
        .Attr<ffi::Span<const int32_t>>("ipar")       \
        .Attr<ffi::Span<const double>>("par0")

XLA_FFI_DEFINE_HANDLER_SYMBOL(lgp_xla_gram_iso, GramIsoImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>() LGP_DESC_ATTRS());
XLA_FFI_DEFINE_HANDLER_SYMBOL(lgp_xla_gram_iso_vjp, GramIsoVjpImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>() LGP_DESC_ATTRS());
XLA_FFI_DEFINE_HANDLER_SYMBOL(lgp_xla_gram_iso_jvp, GramIsoJvpImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>() LGP_DESC_ATTRS());
XLA_FFI_DEFINE_HANDLER_SYMBOL(lgp_xla_chol_factor, CholFactorImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::S32>>()
                                  .Attr<double>("epsrel")
                                  .Attr<double>("epsabs"));
XLA_FFI_DEFINE_HANDLER_SYMBOL(lgp_xla_chol_solve, CholSolveImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Attr<int32_t>("trans"));
XLA_FFI_DEFINE_HANDLER_SYMBOL(lgp_xla_chol_inverse, CholInverseImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Attr<int64_t>("n"));
XLA_FFI_DEFINE_HANDLER_SYMBOL(lgp_xla_chol_logdet_quad, CholLogdetQuadImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>());
