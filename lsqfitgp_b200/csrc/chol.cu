// Blocked recursive FP64 Cholesky with lsqfitgp's equilibration + Gershgorin jitter, triangular
// solves, triangular products, and inverse-from-factor, all built on the DMMA GEMM.
//
// Reference semantics: src/lsqfitgp/_linalg/_decomp.py:245-255 (_parseeps), :349-361
// (eigval_bound, diag_scale_pow2), :380-393 (Chol.__init__), :398-439 (solves), :466-472.
#include <limits.h>
#include <math.h>

#include "../../include/lgp_b200.h"
#include "common.cuh"
#include "gemm_dmma.cuh"
#include "internal.h"

namespace lgp {

// ------------------------------------------------------------------------------------------------
// 1. equilibration + jitter (one pass over K)
// ------------------------------------------------------------------------------------------------

// s_i = 2^rint(log2(K_ii)/2), 1 if K_ii == 0 (reference: diag_scale_pow2, _decomp.py:356-361)
__global__ void chol_diag_scale_kernel(const double *__restrict__ K, int64_t ldk, const double *__restrict__ addmat,
                                       int64_t ldadd, const double *__restrict__ adddiag, int n, int npad,
                                       double *__restrict__ aux) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npad) return;
    double s = 1.0;
    if (i < n) {
        double d = K[(int64_t)i * ldk + i];
        if (addmat) d = d + addmat[(int64_t)i * ldadd + i];
        if (adddiag) d = d + adddiag[i];
        if (d != 0.0) s = exp2(rint(0.5 * log2(d)));
    }
    aux[LGP_AUX_S(npad) + i] = s;
    aux[LGP_AUX_SINV(npad) + i] = 1.0 / s;
    if (i < 16) aux[LGP_AUX_SCALARS(npad) + i] = 0.0;
}

__device__ __forceinline__ void atomic_max_nonneg(double *addr, double v) {
    // non-negative doubles (and NaN, which compares above +inf as bits) order like their bit patterns
    atomicMax(reinterpret_cast<unsigned long long *>(addr), (unsigned long long)__double_as_longlong(v));
}

// One warp per row: Kt = (K + add)/s_i/s_j -> lower triangle of W, row abs-sums -> running max.
// Rows >= n are identity padding.
__global__ void __launch_bounds__(256) chol_prepare_kernel(const double *__restrict__ K, int64_t ldk,
                                                           const double *__restrict__ addmat, int64_t ldadd,
                                                           const double *__restrict__ adddiag, int n, int npad,
                                                           double *__restrict__ W, int64_t ldw,
                                                           double *__restrict__ aux) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + warp;
    if (i >= npad) return;
    double *wrow = W + (int64_t)i * ldw;
    if (i >= n) {
        for (int j = lane; j <= i; j += 32) wrow[j] = (j == i) ? 1.0 : 0.0;
        return;
    }
    const double *sinv = aux + LGP_AUX_SINV(npad);
    const double si = sinv[i];
    const double *krow = K + (int64_t)i * ldk;
    const double *arow = addmat ? addmat + (int64_t)i * ldadd : nullptr;
    double sum = 0.0;
    for (int j = lane; j < n; j += 32) {
        double v = krow[j];
        if (arow) v = v + arow[j];
        if (adddiag && j == i) v = v + adddiag[i];
        v = (v * sinv[j]) * si;  // exact: powers of two
        sum += fabs(v);
        if (j <= i) wrow[j] = v;
    }
    sum = warp_sum(sum);
    if (lane == 0) atomic_max_nonneg(aux + LGP_AUX_SCALARS(npad) + 0, sum);
}

// eps = epsrel * maxrowsum + epsabs; W_ii += eps (i < n); min s^2
__global__ void __launch_bounds__(1024) chol_jitter_kernel(int n, int npad, double epsrel, double epsabs,
                                                           double *__restrict__ W, int64_t ldw,
                                                           double *__restrict__ aux, int32_t *__restrict__ info) {
    double *sc = aux + LGP_AUX_SCALARS(npad);
    const double eps = epsrel * sc[0] + epsabs;
    const double *s = aux + LGP_AUX_S(npad);
    double mn = INFINITY;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        W[(int64_t)i * ldw + i] += eps;
        double v = s[i] * s[i];
        mn = (v < mn || v != v) ? v : mn;
    }
    __shared__ double red[32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double u = __shfl_xor_sync(0xffffffffu, mn, o);
        mn = (u < mn || u != u) ? u : mn;
    }
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mn;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); w++) {
            double u = red[w];
            mn = (u < mn || u != u) ? u : mn;
        }
        sc[1] = eps;
        sc[3] = mn;
        *info = INT_MAX;
    }
}

// ------------------------------------------------------------------------------------------------
// 2. leaf: 128x128 Cholesky + inverse of the factor, one CTA, matrix held in registers
// ------------------------------------------------------------------------------------------------
// Combined storage M[r][c]: r >= c holds A/L[r][c]; r < c holds row c of X = L^-1 (X[c][r]).
// Thread (lane, w) owns rows lane+32a (a<4) and columns w+16b (b<8).  At step k the 32 lanes of warp
// k%16 own column k, so the pivot is broadcast with one shuffle and there is a single barrier per step.
constexpr int LEAF_THREADS = 512;
constexpr int LEAF_SMEM_BYTES = (NB * (NB + 1) + 2 * NB) * 8;

__global__ void __launch_bounds__(LEAF_THREADS, 1) potrf_leaf_kernel(double *__restrict__ Wblk, int64_t ld,
                                                                     double *__restrict__ invd,
                                                                     double *__restrict__ dvec,
                                                                     int32_t *__restrict__ info, int j0) {
    extern __shared__ __align__(16) double leaf_sm[];
    double *stage = leaf_sm;
    double *colbuf = leaf_sm + NB * (NB + 1);
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;

    for (int idx = tid; idx < NB * NB; idx += LEAF_THREADS) {
        int r = idx >> 7, c = idx & 127;
        stage[r * (NB + 1) + c] = Wblk[(int64_t)r * ld + c];
    }
    __syncthreads();
    double M[4][8];
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 8; b++) {
            int r = lane + 32 * a, c = w + 16 * b;
            M[a][b] = (r >= c) ? stage[r * (NB + 1) + c] : 0.0;
        }
    __syncthreads();

    for (int k = 0; k < NB; k++) {
        double *col = colbuf + (k & 1) * NB;
        const int kb = k >> 4, kw = k & 15, ka = k >> 5, kl = k & 31;
        if (w == kw) {
            double cv[4];
#pragma unroll
            for (int b = 0; b < 8; b++)
                if (b == kb) {
#pragma unroll
                    for (int a = 0; a < 4; a++) cv[a] = M[a][b];
                }
            double pv = 0.0;
#pragma unroll
            for (int a = 0; a < 4; a++)
                if (a == ka) pv = cv[a];
            pv = __shfl_sync(0xffffffffu, pv, kl);
            const double l = sqrt(pv);
            const double rl = 1.0 / l;
            if (lane == 0) {
                if (!(pv > 0.0) || !(l < INFINITY)) atomicMin(info, j0 + k + 1);
                dvec[j0 + k] = l;
            }
#pragma unroll
            for (int a = 0; a < 4; a++) {
                int r = lane + 32 * a;
                double v = cv[a] * rl;
                col[r] = (r == k) ? rl : v;
                cv[a] = (r == k) ? l : v;
            }
#pragma unroll
            for (int b = 0; b < 8; b++)
                if (b == kb) {
#pragma unroll
                    for (int a = 0; a < 4; a++) M[a][b] = cv[a];
                }
        }
        __syncthreads();
        double cr[4];
#pragma unroll
        for (int a = 0; a < 4; a++) cr[a] = col[lane + 32 * a];
#pragma unroll
        for (int b = 0; b < 8; b++) {
            const int c = w + 16 * b;
            if (c > k) {
                const double cc = col[c];
#pragma unroll
                for (int a = 0; a < 4; a++) {
                    const int r = lane + 32 * a;
                    if (r >= c || r <= k) M[a][b] -= cr[a] * cc;
                }
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 8; b++) stage[(lane + 32 * a) * (NB + 1) + (w + 16 * b)] = M[a][b];
    __syncthreads();
    for (int idx = tid; idx < NB * NB; idx += LEAF_THREADS) {
        int r = idx >> 7, c = idx & 127;
        double lv = (r >= c) ? stage[r * (NB + 1) + c] : 0.0;
        Wblk[(int64_t)r * ld + c] = lv;
        double xv = (r > c) ? stage[c * (NB + 1) + r] : ((r == c) ? 1.0 / stage[r * (NB + 1) + r] : 0.0);
        invd[idx] = xv;
    }
}

// ------------------------------------------------------------------------------------------------
// 3. small elementwise / reduction kernels
// ------------------------------------------------------------------------------------------------
__global__ void row_scale_kernel(double *__restrict__ B, int64_t ldb, int n, int m, const double *__restrict__ f) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)n * m) return;
    int i = (int)(idx / m), j = (int)(idx % m);
    B[(int64_t)i * ldb + j] *= f[i];
}

__global__ void row_scale_copy_kernel(const double *__restrict__ X, int64_t ldx, double *__restrict__ Y, int64_t ldy,
                                      int n, int m, const double *__restrict__ f) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)n * m) return;
    int i = (int)(idx / m), j = (int)(idx % m);
    Y[(int64_t)i * ldy + j] = X[(int64_t)i * ldx + j] * (f ? f[i] : 1.0);
}

__global__ void get_factor_kernel(const double *__restrict__ W, int64_t ldw, const double *__restrict__ s, int n,
                                  double *__restrict__ L, int64_t ldl) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    int i = blockIdx.y;
    if (j >= n) return;
    L[(int64_t)i * ldl + j] = (j <= i) ? W[(int64_t)i * ldw + j] * s[i] : 0.0;
}

// Kinv_ij *= sinv_i * sinv_j on the lower triangle (undo the equilibration of the inverse)
__global__ void sym_scale_lower_kernel(double *__restrict__ A, int64_t lda, int n, const double *__restrict__ f) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    int i = blockIdx.y;
    if (j > i || i >= n) return;
    A[(int64_t)i * lda + j] = (A[(int64_t)i * lda + j] * f[j]) * f[i];
}

__global__ void copy_block_kernel(const double *__restrict__ src, int64_t lds, double *__restrict__ dst, int64_t ldd,
                                  int rows, int cols) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    int i = blockIdx.y;
    if (j >= cols || i >= rows) return;
    dst[(int64_t)i * ldd + j] = src[(int64_t)i * lds + j];
}

// out[0] = sum_i log(d_i * s_i), out[1] = sum_i a_i^2 ; single CTA, fixed summation order (deterministic)
__global__ void __launch_bounds__(1024) logdet_quad_kernel(const double *__restrict__ d, const double *__restrict__ s,
                                                           const double *__restrict__ a, int n,
                                                           double *__restrict__ out) {
    double ld = 0.0, q = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        ld += log(d[i] * s[i]);
        if (a) q += a[i] * a[i];
    }
    __shared__ double r0[32], r1[32];
    ld = warp_sum(ld);
    q = warp_sum(q);
    if ((threadIdx.x & 31) == 0) {
        r0[threadIdx.x >> 5] = ld;
        r1[threadIdx.x >> 5] = q;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        int nw = blockDim.x >> 5;
        ld = threadIdx.x < nw ? r0[threadIdx.x] : 0.0;
        q = threadIdx.x < nw ? r1[threadIdx.x] : 0.0;
        ld = warp_sum(ld);
        q = warp_sum(q);
        if (threadIdx.x == 0) {
            out[0] = ld;
            out[1] = q;
        }
    }
}

__global__ void finalize_info_kernel(int32_t *info, int n) {
    int v = *info;
    if (v == INT_MAX || v > n) *info = 0;  // pivots in the identity padding cannot fail first
}

// ------------------------------------------------------------------------------------------------
// 4. host-side recursion (all sizes in 128-blocks of the padded matrix)
// ------------------------------------------------------------------------------------------------
struct CholCtx {
    cudaStream_t st;
    double *W;
    int64_t ldw;
    double *invd;  // per-block 128x128 inverses
    double *dvec;
    int32_t *info;
    int rc;
};

#define RC(x)                   \
    do {                        \
        int r__ = (x);          \
        if (r__ != LGP_OK) {    \
            c.rc = r__;         \
            return;             \
        }                       \
    } while (0)

static inline double *Wp(const CholCtx &c, int rb, int cb) { return c.W + (int64_t)rb * NB * c.ldw + (int64_t)cb * NB; }

// X * Lt[jb..jb+nb, jb..jb+nb]^T = B in place; B = W[rb.., jb..] with `rows` rows
static void trsm_right_rec(CholCtx &c, int rb, int rows, int jb, int nb) {
    if (c.rc) return;
    if (nb == 1) {
        double *B = Wp(c, rb, jb);
        RC(gemm_launch(c.st, true, true, rows, NB, NB, 1.0, B, c.ldw, c.invd + (int64_t)jb * NB * NB, NB, B, c.ldw,
                       GEMM_BETA0 | GEMM_B_LOWER_K));
        return;
    }
    int n1 = nb / 2, n2 = nb - n1;
    trsm_right_rec(c, rb, rows, jb, n1);
    if (c.rc) return;
    // B2 -= B1 * L21^T
    RC(gemm_launch(c.st, true, true, rows, n2 * NB, n1 * NB, -1.0, Wp(c, rb, jb), c.ldw, Wp(c, jb + n1, jb), c.ldw,
                   Wp(c, rb, jb + n1), c.ldw, 0));
    trsm_right_rec(c, rb, rows, jb + n1, n2);
}

static void potrf_rec(CholCtx &c, int jb, int nb) {
    if (c.rc) return;
    if (nb == 1) {
        potrf_leaf_kernel<<<1, LEAF_THREADS, LEAF_SMEM_BYTES, c.st>>>(Wp(c, jb, jb), c.ldw,
                                                                      c.invd + (int64_t)jb * NB * NB, c.dvec, c.info,
                                                                      jb * NB);
        if (cudaGetLastError() != cudaSuccess) c.rc = LGP_ERR_CUDA;
        return;
    }
    int n1 = nb / 2, n2 = nb - n1;
    potrf_rec(c, jb, n1);
    trsm_right_rec(c, jb + n1, n2 * NB, jb, n1);
    if (c.rc) return;
    RC(gemm_launch(c.st, true, true, n2 * NB, n2 * NB, n1 * NB, -1.0, Wp(c, jb + n1, jb), c.ldw, Wp(c, jb + n1, jb),
                   c.ldw, Wp(c, jb + n1, jb + n1), c.ldw, GEMM_LOWER));
    potrf_rec(c, jb + n1, n2);
}

struct SolveCtx {
    cudaStream_t st;
    const double *W;
    int64_t ldw;
    const double *invd;
    int n;  // true size (rows of B)
    double *B;
    int64_t ldb;
    int m;
    int rc;
};
static inline const double *Wc(const SolveCtx &c, int rb, int cb) {
    return c.W + (int64_t)rb * NB * c.ldw + (int64_t)cb * NB;
}
static inline int rows_in(const SolveCtx &c, int jb, int nb) {
    int lo = jb * NB, hi = (jb + nb) * NB;
    if (hi > c.n) hi = c.n;
    return hi > lo ? hi - lo : 0;
}

// Lt X = B (forward), blocks [jb, jb+nb)
static void solve_lower_rec(SolveCtx &c, int jb, int nb) {
    if (c.rc) return;
    int rows = rows_in(c, jb, nb);
    if (rows == 0) return;
    double *Bj = c.B + (int64_t)jb * NB * c.ldb;
    if (nb == 1) {
        RC(gemm_launch(c.st, true, false, rows, c.m, rows, 1.0, c.invd + (int64_t)jb * NB * NB, NB, Bj, c.ldb, Bj,
                       c.ldb, GEMM_BETA0 | GEMM_A_LOWER_K));
        return;
    }
    int n1 = nb / 2, n2 = nb - n1;
    solve_lower_rec(c, jb, n1);
    if (c.rc) return;
    int rows2 = rows_in(c, jb + n1, n2);
    if (rows2 > 0) {
        RC(gemm_launch(c.st, true, false, rows2, c.m, n1 * NB, -1.0, Wc(c, jb + n1, jb), c.ldw, Bj, c.ldb,
                       Bj + (int64_t)n1 * NB * c.ldb, c.ldb, 0));
        solve_lower_rec(c, jb + n1, n2);
    }
}

// Lt^T X = B (backward)
static void solve_upper_rec(SolveCtx &c, int jb, int nb) {
    if (c.rc) return;
    int rows = rows_in(c, jb, nb);
    if (rows == 0) return;
    double *Bj = c.B + (int64_t)jb * NB * c.ldb;
    if (nb == 1) {
        RC(gemm_launch(c.st, false, false, rows, c.m, rows, 1.0, c.invd + (int64_t)jb * NB * NB, NB, Bj, c.ldb, Bj,
                       c.ldb, GEMM_BETA0 | GEMM_A_UPPER_K));
        return;
    }
    int n1 = nb / 2, n2 = nb - n1;
    int rows2 = rows_in(c, jb + n1, n2);
    if (rows2 > 0) {
        solve_upper_rec(c, jb + n1, n2);
        if (c.rc) return;
        // B1 -= L21^T X2 : Aop[i][k] = L21[k][i]
        RC(gemm_launch(c.st, false, false, n1 * NB, c.m, rows2, -1.0, Wc(c, jb + n1, jb), c.ldw,
                       Bj + (int64_t)n1 * NB * c.ldb, c.ldb, Bj, c.ldb, 0));
    }
    solve_upper_rec(c, jb, n1);
}

// X = Lt^-1 (lower) out of place into X (ld = ldx); the strict upper triangle of X is scratch.
struct InvCtx {
    cudaStream_t st;
    const double *W;
    int64_t ldw;
    const double *invd;
    double *X;
    int64_t ldx;
    int rc;
};
static void trtri_rec(InvCtx &c, int jb, int nb) {
    if (c.rc) return;
    if (nb == 1) {
        dim3 g((NB + 127) / 128, NB);
        copy_block_kernel<<<g, 128, 0, c.st>>>(c.invd + (int64_t)jb * NB * NB, NB,
                                               c.X + (int64_t)jb * NB * c.ldx + (int64_t)jb * NB, c.ldx, NB, NB);
        if (cudaGetLastError() != cudaSuccess) c.rc = LGP_ERR_CUDA;
        return;
    }
    int n1 = nb / 2, n2 = nb - n1;
    trtri_rec(c, jb, n1);
    trtri_rec(c, jb + n1, n2);
    if (c.rc) return;
    const double *L21 = c.W + (int64_t)(jb + n1) * NB * c.ldw + (int64_t)jb * NB;
    double *X11 = c.X + (int64_t)jb * NB * c.ldx + (int64_t)jb * NB;
    double *X22 = c.X + (int64_t)(jb + n1) * NB * c.ldx + (int64_t)(jb + n1) * NB;
    double *X21 = c.X + (int64_t)(jb + n1) * NB * c.ldx + (int64_t)jb * NB;
    double *Tt = c.X + (int64_t)jb * NB * c.ldx + (int64_t)(jb + n1) * NB;  // n1 x n2 scratch (upper block)
    // Tt[j][i] = sum_k X11[k][j] * L21[i][k]   (k >= j)
    RC(gemm_launch(c.st, false, true, n1 * NB, n2 * NB, n1 * NB, 1.0, X11, c.ldx, L21, c.ldw, Tt, c.ldx,
                   GEMM_BETA0 | GEMM_A_UPPER_K));
    // X21[i][j] = - sum_k X22[i][k] * Tt[j][k]  (k <= i)
    RC(gemm_launch(c.st, true, true, n2 * NB, n1 * NB, n2 * NB, -1.0, X22, c.ldx, Tt, c.ldx, X21, c.ldx,
                   GEMM_BETA0 | GEMM_A_LOWER_K));
}

}  // namespace lgp

using namespace lgp;

extern "C" {

int64_t lgp_chol_npad(int64_t n) { return (n + NB - 1) / NB * NB; }
int64_t lgp_chol_aux_doubles(int64_t n) {
    int64_t npad = lgp_chol_npad(n);
    return 3 * npad + 16 + (npad / NB) * (int64_t)NB * NB;
}

int lgp_chol_factor(lgp_stream_t stream, const double *K, int64_t ldk, const double *addmat, int64_t ldadd,
                    const double *adddiag, int64_t n64, double epsrel, double epsabs, double *W, int64_t ldw,
                    double *aux, int32_t *info) {
    if (n64 < 1 || n64 > (1 << 30) || !K || !W || !aux || !info) return LGP_ERR_BADARG;
    const int n = (int)n64, npad = (int)lgp_chol_npad(n);
    if (ldw < npad || ldk < n || (addmat && ldadd < n)) return LGP_ERR_BADARG;
    if ((ldw & 1) || (reinterpret_cast<uintptr_t>(W) & 15) || (reinterpret_cast<uintptr_t>(aux) & 15))
        return LGP_ERR_ALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    static bool attr = false;
    if (!attr) {
        if (cudaFuncSetAttribute(potrf_leaf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LEAF_SMEM_BYTES) !=
            cudaSuccess)
            return LGP_ERR_CUDA;
        attr = true;
    }
    if (epsrel < 0) epsrel = (double)n * 2.220446049250313e-16;
    chol_diag_scale_kernel<<<(npad + 255) / 256, 256, 0, st>>>(K, ldk, addmat, ldadd, adddiag, n, npad, aux);
    LGP_CUDA_CHECK_LAUNCH();
    chol_prepare_kernel<<<(npad + 7) / 8, 256, 0, st>>>(K, ldk, addmat, ldadd, adddiag, n, npad, W, ldw, aux);
    LGP_CUDA_CHECK_LAUNCH();
    chol_jitter_kernel<<<1, 1024, 0, st>>>(n, npad, epsrel, epsabs, W, ldw, aux, info);
    LGP_CUDA_CHECK_LAUNCH();
    CholCtx c{st, W, ldw, aux + LGP_AUX_INVDIAG(npad), aux + LGP_AUX_DIAG(npad), info, LGP_OK};
    potrf_rec(c, 0, npad / NB);
    if (c.rc) return c.rc;
    finalize_info_kernel<<<1, 1, 0, st>>>(info, n);
    LGP_CUDA_CHECK_LAUNCH();
    logdet_quad_kernel<<<1, 1024, 0, st>>>(aux + LGP_AUX_DIAG(npad), aux + LGP_AUX_S(npad), nullptr, n,
                                           aux + LGP_AUX_SCALARS(npad) + 4);
    LGP_CUDA_CHECK_LAUNCH();
    return LGP_OK;
}

int lgp_chol_solve(lgp_stream_t stream, const double *W, int64_t ldw, const double *aux, int64_t n64, double *B,
                   int64_t ldb, int64_t m64, int trans) {
    if (n64 < 1 || m64 < 1 || !W || !aux || !B) return LGP_ERR_BADARG;
    const int n = (int)n64, m = (int)m64, npad = (int)lgp_chol_npad(n);
    if ((ldb & 1) || (reinterpret_cast<uintptr_t>(B) & 15)) return LGP_ERR_ALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    const double *sinv = aux + LGP_AUX_SINV(npad);
    int64_t total = (int64_t)n * m;
    SolveCtx c{st, W, ldw, aux + LGP_AUX_INVDIAG(npad), n, B, ldb, m, LGP_OK};
    if (!trans) {
        row_scale_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(B, ldb, n, m, sinv);
        LGP_CUDA_CHECK_LAUNCH();
        solve_lower_rec(c, 0, npad / NB);
    } else {
        solve_upper_rec(c, 0, npad / NB);
        if (c.rc) return c.rc;
        row_scale_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(B, ldb, n, m, sinv);
        LGP_CUDA_CHECK_LAUNCH();
    }
    return c.rc;
}

int lgp_chol_mult(lgp_stream_t stream, const double *W, int64_t ldw, const double *aux, int64_t n64, const double *X,
                  int64_t ldx, int64_t m64, double *Y, int64_t ldy, double *tmp, int64_t ldt, int trans) {
    if (n64 < 1 || m64 < 1 || !W || !aux || !X || !Y) return LGP_ERR_BADARG;
    const int n = (int)n64, m = (int)m64, npad = (int)lgp_chol_npad(n);
    cudaStream_t st = (cudaStream_t)stream;
    const double *s = aux + LGP_AUX_S(npad);
    int64_t total = (int64_t)n * m;
    int rc;
    if (!trans) {
        // Y = S (Lt X): Aop[i][k] = Lt[i][k] (k <= i), Bop[j][k] = X[k][j]
        rc = gemm_launch(st, true, false, n, m, n, 1.0, W, ldw, X, ldx, Y, ldy, GEMM_BETA0 | GEMM_A_LOWER_K);
        if (rc) return rc;
        row_scale_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(Y, ldy, n, m, s);
        LGP_CUDA_CHECK_LAUNCH();
    } else {
        // Y = Lt^T (S X): tmp = S X, then Aop[i][k] = Lt[k][i] (k >= i)
        if (!tmp) return LGP_ERR_BADARG;
        row_scale_copy_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(X, ldx, tmp, ldt, n, m, s);
        LGP_CUDA_CHECK_LAUNCH();
        rc = gemm_launch(st, false, false, n, m, n, 1.0, W, ldw, tmp, ldt, Y, ldy, GEMM_BETA0 | GEMM_A_UPPER_K);
        if (rc) return rc;
    }
    return LGP_OK;
}

int lgp_chol_get_factor(lgp_stream_t stream, const double *W, int64_t ldw, const double *aux, int64_t n64,
                        double *Lout, int64_t ldl) {
    if (n64 < 1 || !W || !aux || !Lout) return LGP_ERR_BADARG;
    const int n = (int)n64, npad = (int)lgp_chol_npad(n);
    dim3 g((n + 255) / 256, n);
    get_factor_kernel<<<g, 256, 0, (cudaStream_t)stream>>>(W, ldw, aux + LGP_AUX_S(npad), n, Lout, ldl);
    LGP_CUDA_CHECK_LAUNCH();
    return LGP_OK;
}

int lgp_chol_inverse(lgp_stream_t stream, const double *W, int64_t ldw, const double *aux, int64_t n64,
                     double *scratch, double *Kinv, int64_t ldk) {
    if (n64 < 1 || !W || !aux || !scratch || !Kinv) return LGP_ERR_BADARG;
    const int n = (int)n64, npad = (int)lgp_chol_npad(n);
    if (ldk < npad) return LGP_ERR_BADARG;
    if ((ldk & 1) || (reinterpret_cast<uintptr_t>(Kinv) & 15) || (reinterpret_cast<uintptr_t>(scratch) & 15))
        return LGP_ERR_ALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    InvCtx c{st, W, ldw, aux + LGP_AUX_INVDIAG(npad), scratch, (int64_t)npad, LGP_OK};
    trtri_rec(c, 0, npad / NB);
    if (c.rc) return c.rc;
    // Kinv[i][j] = sum_{k >= i} X[k][i] X[k][j], j <= i   (LAUUM as one triangular-K SYRK launch)
    int rc = gemm_launch(st, false, false, npad, npad, npad, 1.0, scratch, npad, scratch, npad, Kinv, ldk,
                         GEMM_BETA0 | GEMM_LOWER | GEMM_A_UPPER_K);
    if (rc) return rc;
    dim3 g((n + 255) / 256, n);
    sym_scale_lower_kernel<<<g, 256, 0, st>>>(Kinv, ldk, n, aux + LGP_AUX_SINV(npad));
    LGP_CUDA_CHECK_LAUNCH();
    return LGP_OK;
}

int lgp_chol_logdet_quad(lgp_stream_t stream, const double *aux, int64_t n64, const double *a, double *out) {
    if (n64 < 1 || !aux || !out) return LGP_ERR_BADARG;
    const int n = (int)n64, npad = (int)lgp_chol_npad(n);
    logdet_quad_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(aux + LGP_AUX_DIAG(npad), aux + LGP_AUX_S(npad), a, n,
                                                             out);
    LGP_CUDA_CHECK_LAUNCH();
    return LGP_OK;
}

}  // extern "C"
