// Misc C-ABI entry points (version, GEMM wrapper).
#include "../../include/lgp_b200.h"
#include "common.cuh"
#include "internal.h"

#include <atomic>
namespace lgp {
static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
// Y = alpha * X + beta * Y (+ gamma on the diagonal), elementwise over an n x m block
__global__ void axpby_kernel(int64_t n, int64_t m, double alpha, const double *__restrict__ X, int64_t ldx, double beta,
                             double *__restrict__ Y, int64_t ldy, double gamma) {
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t i = blockIdx.y;
    if (j >= m || i >= n) return;
    double v = (X ? alpha * X[i * ldx + j] : 0.0);
    if (beta != 0.0) v += beta * Y[i * ldy + j];
    if (i == j) v += gamma;
    Y[i * ldy + j] = v;
}
}  // namespace lgp

extern "C" {

int lgp_abi_version(void) { return LGP_ABI_VERSION; }

long long lgp_launch_count(void) { return lgp::g_launches.load(); }

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

int lgp_axpby(lgp_stream_t stream, int64_t n, int64_t m, double alpha, const double *X, int64_t ldx, double beta,
              double *Y, int64_t ldy, double gamma) {
    if (n < 0 || m < 0 || !Y) return LGP_ERR_BADARG;
    if (n == 0 || m == 0) return LGP_OK;
    if (n > 2147483647LL) return LGP_ERR_UNSUPPORTED;
    dim3 grid((unsigned)((m + 255) / 256), (unsigned)(n > 65535 ? 65535 : n));
    if (n > 65535) {
        // rows beyond the grid.y limit: launch in slabs
        for (int64_t r0 = 0; r0 < n; r0 += 65535) {
            int64_t rows = n - r0 < 65535 ? n - r0 : 65535;
            dim3 g((unsigned)((m + 255) / 256), (unsigned)rows);
            // the diagonal offset changes with the slab: shift pointers so that i == j still marks it
            if (gamma != 0.0) return LGP_ERR_UNSUPPORTED;
            lgp::axpby_kernel<<<g, 256, 0, (cudaStream_t)stream>>>(rows, m, alpha, X ? X + r0 * ldx : nullptr, ldx,
                                                                    beta, Y + r0 * ldy, ldy, 0.0);
        }
    } else {
        lgp::axpby_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(n, m, alpha, X, ldx, beta, Y, ldy, gamma);
    }
    LGP_CUDA_CHECK_LAUNCH();
    return LGP_OK;
}

}  // extern "C"
