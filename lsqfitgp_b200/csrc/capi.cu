// Misc C-ABI entry points (version, GEMM wrapper).
#include "../../include/lgp_b200.h"
#include "common.cuh"
#include "internal.h"

#include <atomic>
#include <mutex>
namespace lgp {
int current_device() {
    int d = -1;
    return cudaGetDevice(&d) == cudaSuccess && d >= 0 && d < MAX_DEVICES ? d : -1;
}
static std::mutex &once_mutex() {
    static std::mutex m;
    return m;
}
bool DeviceOnce::done(int dev) const {
    std::lock_guard<std::mutex> lock(once_mutex());
    return dev >= 0 && ((mask >> dev) & 1ull);
}
void DeviceOnce::set(int dev) {
    std::lock_guard<std::mutex> lock(once_mutex());
    if (dev >= 0) mask |= 1ull << dev;
}
cudaEvent_t ring_event() {
    constexpr int RING = 64;
    struct Ring {
        cudaEvent_t ev[RING];
        int next, made;
    };
    static thread_local Ring *rings[MAX_DEVICES];
    const int dev = current_device();
    if (dev < 0) return nullptr;
    Ring *r = rings[dev];
    if (!r) {
        r = rings[dev] = new Ring();
        r->next = r->made = 0;
    }
    if (r->made < RING) {
        if (cudaEventCreateWithFlags(&r->ev[r->made], cudaEventDisableTiming) != cudaSuccess) return nullptr;
        return r->ev[r->made++];
    }
    cudaEvent_t e = r->ev[r->next];
    r->next = (r->next + 1) % RING;
    return e;
}
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

// register-resident loops used to MEASURE the FP64 roofline denominators in the run that reports them
template <int NACC>
__global__ void __launch_bounds__(256) peak_dmma_kernel(double *out, int iters) {
    double c[NACC][2];
    double a = threadIdx.x * 1e-9, b = 1.0 + threadIdx.x * 1e-9;
#pragma unroll
    for (int i = 0; i < NACC; i++) {
        c[i][0] = i;
        c[i][1] = -i;
    }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) dmma884(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
__global__ void __launch_bounds__(256) peak_dfma_kernel(double *out, int iters) {
    double c[NACC];
    double a = 1.0 + threadIdx.x * 1e-12, b = threadIdx.x * 1e-9;
#pragma unroll
    for (int i = 0; i < NACC; i++) c[i] = i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) c[i] = fma(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
}  // namespace lgp

extern "C" {

int lgp_peak_probe(lgp_stream_t stream, int kind, int iters, double *scratch, int64_t scratch_doubles, double *flops_out) {
    if (iters < 1 || !scratch || !flops_out || (kind != LGP_PEAK_DMMA && kind != LGP_PEAK_DFMA)) return LGP_ERR_BADARG;
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
        return LGP_ERR_CUDA;
    const int ctas = 2 * sms;
    if (scratch_doubles < (int64_t)ctas * 256) return LGP_ERR_BADARG;
    if (kind == LGP_PEAK_DMMA) {
        lgp::peak_dmma_kernel<16><<<ctas, 256, 0, (cudaStream_t)stream>>>(scratch, iters);
        // per warp-level m8n8k4: 8*8*4 FMA = 512 flop; 8 warps per CTA, 16 accumulators
        *flops_out = 512.0 * 8.0 * 16.0 * (double)iters * (double)ctas;
    } else {
        lgp::peak_dfma_kernel<16><<<ctas, 256, 0, (cudaStream_t)stream>>>(scratch, iters);
        *flops_out = 2.0 * 16.0 * 256.0 * (double)iters * (double)ctas;
    }
    LGP_CUDA_CHECK_LAUNCH();
    return LGP_OK;
}

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
