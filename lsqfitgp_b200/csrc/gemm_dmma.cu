// Launcher for the FP64 DMMA GEMM (see gemm_dmma.cuh).
#include "gemm_dmma.cuh"
#include "internal.h"
#include <string.h>

namespace lgp {

template <class Cfg, bool AK, bool BK_>
static int launch_cfg(cudaStream_t stream, GemmParams &p) {
    p.tiles_m = (p.M + Cfg::BM - 1) / Cfg::BM;
    p.tiles_n = (p.N + Cfg::BN - 1) / Cfg::BN;
    int64_t grid = (p.flags & GEMM_LOWER) ? (int64_t)(Cfg::BM >= Cfg::BN ? Cfg::BM / Cfg::BN : 1) * p.tiles_m * (p.tiles_m + 1) / 2
                                          : (int64_t)p.tiles_m * p.tiles_n;
    if (grid > 2147483647LL) return LGP_ERR_UNSUPPORTED;
    static DeviceOnce attr_set;  // one flag per template instantiation and device
    const int dev = current_device();
    if (!attr_set.done(dev)) {
        if (cudaFuncSetAttribute(gemm_dmma_kernel<Cfg, AK, BK_>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 Cfg::SMEM_BYTES) != cudaSuccess)
            return LGP_ERR_CUDA;
        attr_set.set(dev);
    }
    gemm_dmma_kernel<Cfg, AK, BK_><<<(unsigned)grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(p);
    LGP_CUDA_CHECK_LAUNCH();
    return LGP_OK;
}

template <class Cfg>
static int launch_layout(cudaStream_t stream, bool a_kmaj, bool b_kmaj, GemmParams &p) {
    if (a_kmaj && b_kmaj) return launch_cfg<Cfg, true, true>(stream, p);
    if (a_kmaj && !b_kmaj) return launch_cfg<Cfg, true, false>(stream, p);
    if (!a_kmaj && b_kmaj) return launch_cfg<Cfg, false, true>(stream, p);
    return launch_cfg<Cfg, false, false>(stream, p);
}

// Host-side launcher (internal; the C ABI wrapper is lgp_dgemm in capi.cu).
int gemm_launch(cudaStream_t stream, bool a_kmaj, bool b_kmaj, int M, int N, int K, double alpha, const double *A,
                int64_t lda, const double *B, int64_t ldb, double *C, int64_t ldc, int flags, const GemmMirror *mir) {
    if (M <= 0 || N <= 0) return LGP_OK;
    if (mir && (mir->n < 0 || mir->n > GEMM_MAX_MIRRORS || (mir->multimem && mir->n != 1))) return LGP_ERR_BADARG;
    if (mir)
        for (int i = 0; i < mir->n; i++)
            if (reinterpret_cast<uintptr_t>(mir->dst[i]) & 15) return LGP_ERR_ALIGN;
    if ((lda & 1) || (ldb & 1) || (reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15))
        return LGP_ERR_ALIGN;
    if ((flags & GEMM_LOWER) && M != N) return LGP_ERR_BADARG;
    if ((flags & GEMM_INPLACE_A) && N > 128) return LGP_ERR_BADARG;
    if ((flags & GEMM_INPLACE_B) && M > 128) return LGP_ERR_BADARG;
    GemmParams p;
    p.A = A; p.B = B; p.C = C;
    p.M = M; p.N = N; p.K = K;
    p.lda = lda; p.ldb = ldb; p.ldc = ldc;
    p.alpha = alpha;
    p.flags = flags;
    if (mir)
        p.mir = *mir;
    else
        memset(&p.mir, 0, sizeof(p.mir));
    // tile configuration: the in-place products need the aliased dimension inside one tile
    if (flags & GEMM_INPLACE_B) {
        if (a_kmaj && !b_kmaj) return launch_cfg<GemmWide, true, false>(stream, p);
        if (!a_kmaj && !b_kmaj) return launch_cfg<GemmWide, false, false>(stream, p);
        return LGP_ERR_UNSUPPORTED;
    }
    if (flags & GEMM_INPLACE_A) {
        if (a_kmaj && b_kmaj) return launch_cfg<GemmTall, true, true>(stream, p);
        return LGP_ERR_UNSUPPORTED;
    }
    return launch_layout<GemmBig>(stream, a_kmaj, b_kmaj, p);
}

}  // namespace lgp
