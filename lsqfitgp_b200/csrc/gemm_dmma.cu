// Launcher for the FP64 DMMA GEMM (see gemm_dmma.cuh).
#include "gemm_dmma.cuh"
#include "internal.h"

namespace lgp {

// Host-side launcher (internal; the C ABI wrapper is lgp_dgemm in capi.cu).
int gemm_launch(cudaStream_t stream, bool a_kmaj, bool b_kmaj, int M, int N, int K, double alpha,
                       const double *A, int64_t lda, const double *B, int64_t ldb, double *C, int64_t ldc,
                       int flags) {
    if (M <= 0 || N <= 0) return LGP_OK;
    if ((lda & 1) || (ldb & 1) || (reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15))
        return LGP_ERR_ALIGN;
    if ((flags & GEMM_LOWER) && M != N) return LGP_ERR_BADARG;
    GemmParams p;
    p.A = A; p.B = B; p.C = C;
    p.M = M; p.N = N; p.K = K;
    p.lda = lda; p.ldb = ldb; p.ldc = ldc;
    p.alpha = alpha;
    p.flags = flags;
    p.tiles_m = (M + GEMM_BM - 1) / GEMM_BM;
    p.tiles_n = (N + GEMM_BN - 1) / GEMM_BN;
    int grid = (flags & GEMM_LOWER) ? p.tiles_m * (p.tiles_m + 1) / 2 : p.tiles_m * p.tiles_n;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(gemm_dmma_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES);
        cudaFuncSetAttribute(gemm_dmma_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES);
        cudaFuncSetAttribute(gemm_dmma_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES);
        cudaFuncSetAttribute(gemm_dmma_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES);
        attr_set = true;
    }
    if (a_kmaj && b_kmaj)
        gemm_dmma_kernel<true, true><<<grid, GEMM_THREADS, GEMM_SMEM_BYTES, stream>>>(p);
    else if (a_kmaj && !b_kmaj)
        gemm_dmma_kernel<true, false><<<grid, GEMM_THREADS, GEMM_SMEM_BYTES, stream>>>(p);
    else if (!a_kmaj && b_kmaj)
        gemm_dmma_kernel<false, true><<<grid, GEMM_THREADS, GEMM_SMEM_BYTES, stream>>>(p);
    else
        gemm_dmma_kernel<false, false><<<grid, GEMM_THREADS, GEMM_SMEM_BYTES, stream>>>(p);
    LGP_CUDA_CHECK_LAUNCH();
    return LGP_OK;
}

}  // namespace lgp
