// Misc C-ABI entry points (version, GEMM wrapper).
#include "../../include/lgp_b200.h"
#include "common.cuh"
#include "internal.h"

extern "C" {

int lgp_abi_version(void) { return LGP_ABI_VERSION; }

const char *lgp_build_info(void) {
    return "liblgpb200 abi=1 arch=sm_100a nvcc=" LGP_STR(__CUDACC_VER_MAJOR__) "." LGP_STR(__CUDACC_VER_MINOR__)
           " fp64-tensor=mma.sync.m8n8k4 (DMMA.8x8x4)";
}

int lgp_dgemm(lgp_stream_t stream, int a_kmajor, int b_kmajor, int64_t M, int64_t N, int64_t K, double alpha,
              const double *A, int64_t lda, const double *B, int64_t ldb, double *C, int64_t ldc, int flags) {
    if (M < 0 || N < 0 || K < 0 || M > (1 << 30) || N > (1 << 30) || K > (1 << 30)) return LGP_ERR_BADARG;
    if (!A || !B || !C) return LGP_ERR_BADARG;
    return lgp::gemm_launch((cudaStream_t)stream, a_kmajor != 0, b_kmajor != 0, (int)M, (int)N, (int)K, alpha, A,
                            lda, B, ldb, C, ldc, flags);
}

}  // extern "C"
