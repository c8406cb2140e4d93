// Local (per-GPU) kernels of the 2-D block-cyclic multi-GPU Cholesky (lsqfitgp_b200/_dist.py: DistChol).
//
// The reference is single-device; what is distributed here is Chol.__init__
// (src/lsqfitgp/_linalg/_decomp.py:380-393) for matrices that do not fit one GPU.  Tile (I, J) of the T x T tiling
// lives on process (I mod Pr, J mod Pc) at local tile position (I div Pr, J div Pc) of a dense row-major local
// matrix, so that every local operation below is one large strided GEMM / streaming pass.
#include <math.h>

#include "../../include/lgp_b200.h"
#include "common.cuh"
#include "gemm_dmma.cuh"
#include "internal.h"

namespace lgp {

struct Grid {
    int64_t n;
    int T, NT, Pr, Pc, pr, pc, LR, LC;
};

static bool grid_ok(const lgp_grid_t *g, Grid &o) {
    if (!g || g->n < 1 || g->tile < NB || g->tile % NB || g->nprow < 1 || g->npcol < 1 || g->prow < 0 ||
        g->prow >= g->nprow || g->pcol < 0 || g->pcol >= g->npcol)
        return false;
    o.n = g->n;
    o.T = g->tile;
    o.NT = (int)((g->n + g->tile - 1) / g->tile);
    o.Pr = g->nprow;
    o.Pc = g->npcol;
    o.pr = g->prow;
    o.pc = g->pcol;
    o.LR = o.NT > o.pr ? (o.NT - o.pr + o.Pr - 1) / o.Pr : 0;
    o.LC = o.NT > o.pc ? (o.NT - o.pc + o.Pc - 1) / o.Pc : 0;
    return true;
}

// number of local tile rows of process row r whose global tile index is < J
static inline int tiles_before(int J, int r, int P) { return J > r ? (J - r + P - 1) / P : 0; }

__device__ __forceinline__ int64_t glob_index(int64_t loc, int T, int p, int P) {
    return ((int64_t)p + (int64_t)P * (loc / T)) * T + loc % T;
}

// d[gi] = A_ii for the diagonal entries stored on this process
__global__ void dist_diag_kernel(const double *__restrict__ A, int64_t lda, Grid g, double *__restrict__ d) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.n) return;
    int I = (int)(i / g.T);
    if (I % g.Pr != g.pr || I % g.Pc != g.pc) return;
    int64_t lr = (int64_t)(I / g.Pr) * g.T + i % g.T, lc = (int64_t)(I / g.Pc) * g.T + i % g.T;
    d[i] = A[lr * lda + lc];
}

// s_i = 2^rint(log2(d_i)/2), 1 if d_i == 0 (diag_scale_pow2, _decomp.py:356-361); entries i >= n (padding): 1
__global__ void dist_scale_kernel(const double *__restrict__ d, int64_t n, int64_t npad, double *__restrict__ s,
                                  double *__restrict__ sinv) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npad) return;
    double v = 1.0;
    if (i < n && d[i] != 0.0) v = exp2(rint(0.5 * log2(d[i])));
    s[i] = v;
    sinv[i] = 1.0 / v;
}

// One warp per local row: A <- A/s_i/s_j (identity in the padding), rowsum[gi] = sum_j |A_ij| over the local columns
__global__ void __launch_bounds__(256) dist_prepare_kernel(double *__restrict__ A, int64_t lda, Grid g,
                                                           const double *__restrict__ sinv,
                                                           double *__restrict__ rowsum) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * 8 + warp;
    if (r >= (int64_t)g.LR * g.T) return;
    const int64_t gi = glob_index(r, g.T, g.pr, g.Pr);
    double *row = A + r * lda;
    const int64_t ncol = (int64_t)g.LC * g.T;
    const double si = gi < g.n ? sinv[gi] : 1.0;
    double sum = 0.0;
    for (int64_t c = lane; c < ncol; c += 32) {
        const int64_t gj = glob_index(c, g.T, g.pc, g.Pc);
        double v;
        if (gi >= g.n || gj >= g.n) {
            v = (gi == gj) ? 1.0 : 0.0;
        } else {
            v = (row[c] * sinv[gj]) * si;  // exact: powers of two
            sum += fabs(v);
        }
        row[c] = v;
    }
    sum = warp_sum(sum);
    if (lane == 0 && gi < g.n) rowsum[gi] = sum;
}

// ---- lower-packed local storage: local tile column lj (global tile column J = pc + Pc lj) is ONE contiguous panel
// (rows x T, leading dimension T) holding only the local tile rows li >= first = tiles_before(J, pr, Pr), i.e. the tiles
// I >= J.  Half the memory of the dense local matrix: n = 150000 fits one B200 (84 GiB), SURVEY.md section 8(e).
struct PanelGeom {
    int64_t n;
    int T, Pr, pr, J, first, rows_t;  // rows_t: number of tile rows stored in the panel
};
__device__ __forceinline__ int64_t panel_grow(const PanelGeom &p, int64_t r) {
    return ((int64_t)p.pr + (int64_t)p.Pr * (p.first + r / p.T)) * p.T + r % p.T;
}

// d[gi] = A_ii for the diagonal entries of the panel's first tile, if that tile is the diagonal tile (J, J)
__global__ void panel_diag_kernel(const double *__restrict__ P, PanelGeom p, double *__restrict__ d) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= p.T) return;
    const int64_t gi = (int64_t)p.J * p.T + t;
    if (gi < p.n) d[gi] = P[(int64_t)t * p.T + t];
}
__global__ void panel_add_diag_kernel(double *__restrict__ P, PanelGeom p, const double *__restrict__ eps) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= p.T) return;
    const int64_t gi = (int64_t)p.J * p.T + t;
    if (gi < p.n) P[(int64_t)t * p.T + t] += eps[0];
}

// P <- P / s_i / s_j (identity in the padding); rowsum[gi] += sum_j |P_ij| and, for tiles strictly below the diagonal,
// rowsum[gj] += sum_i |P_ij| (the mirrored entries that are not stored).  CTA = 64 rows x T columns, thread = 4+ columns.
constexpr int PP_ROWS = 64;
__global__ void __launch_bounds__(256) panel_prepare_kernel(double *__restrict__ P, PanelGeom p,
                                                            const double *__restrict__ sinv,
                                                            double *__restrict__ rowsum) {
    __shared__ double rs[PP_ROWS];
    const int tid = threadIdx.x;
    const int64_t r0 = (int64_t)blockIdx.x * PP_ROWS;
    const int64_t nrows = (int64_t)p.rows_t * p.T;
    if (tid < PP_ROWS) rs[tid] = 0.0;
    __syncthreads();
    const bool diag_tile = (p.pr + p.Pr * (p.first + (int)(r0 / p.T))) == p.J;  // a CTA never straddles tiles (T % 64 == 0)
    for (int c = tid; c < p.T; c += 256) {
        const int64_t gj = (int64_t)p.J * p.T + c;
        const double sj = gj < p.n ? sinv[gj] : 1.0;
        double colsum = 0.0;
        for (int q = 0; q < PP_ROWS; q++) {
            const int64_t r = r0 + q;
            if (r >= nrows) break;
            const int64_t gi = panel_grow(p, r);
            double v;
            if (gi >= p.n || gj >= p.n) {
                v = (gi == gj) ? 1.0 : 0.0;
            } else {
                v = (P[r * p.T + c] * sj) * sinv[gi];  // exact: powers of two
                const double a = fabs(v);
                colsum += a;
                atomicAdd(&rs[q], a);
            }
            P[r * p.T + c] = v;
        }
        if (!diag_tile && gj < p.n && colsum != 0.0) atomicAdd(rowsum + gj, colsum);
    }
    __syncthreads();
    if (tid < PP_ROWS) {
        const int64_t r = r0 + tid;
        if (r < nrows) {
            const int64_t gi = panel_grow(p, r);
            if (gi < p.n) atomicAdd(rowsum + gi, rs[tid]);
        }
    }
}

// out[0] = max_i rowsum_i, out[1] = eps = epsrel*max + epsabs  (eigval_bound + _parseeps, _decomp.py:245-255,349-354)
__global__ void __launch_bounds__(1024) dist_eps_kernel(const double *__restrict__ rowsum, int64_t n, double epsrel,
                                                        double epsabs, double *__restrict__ out) {
    double m = 0.0;
    bool nan = false;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        double v = rowsum[i];
        nan |= (v != v);
        m = fmax(m, v);
    }
    __shared__ double red[32];
    __shared__ int rnan[32];
    m = warp_max(m);
    nan = __any_sync(0xffffffffu, nan);
    if ((threadIdx.x & 31) == 0) {
        red[threadIdx.x >> 5] = m;
        rnan[threadIdx.x >> 5] = nan;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); w++) {
            m = fmax(m, red[w]);
            nan |= rnan[w] != 0;
        }
        if (nan) m = NAN;
        out[0] = m;
        out[1] = epsrel * m + epsabs;
    }
}

__global__ void dist_add_diag_kernel(double *__restrict__ A, int64_t lda, Grid g, const double *__restrict__ eps) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.n) return;
    int I = (int)(i / g.T);
    if (I % g.Pr != g.pr || I % g.Pc != g.pc) return;
    int64_t lr = (int64_t)(I / g.Pr) * g.T + i % g.T, lc = (int64_t)(I / g.Pc) * g.T + i % g.T;
    A[lr * lda + lc] += eps[0];
}

// y[r] += alpha * sum_c P[r][c] x[c]   (one warp per row, 8 rows per CTA)
__global__ void __launch_bounds__(256) gemv_n_kernel(const double *__restrict__ P, int64_t ldp, int64_t rows, int cols,
                                                     const double *__restrict__ x, double *__restrict__ y,
                                                     double alpha) {
    extern __shared__ double xs[];
    for (int c = threadIdx.x; c < cols; c += blockDim.x) xs[c] = x[c];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * 8 + warp;
    if (r >= rows) return;
    const double *row = P + r * ldp;
    double acc = 0.0;
    for (int c = lane; c < cols; c += 32) acc += row[c] * xs[c];
    acc = warp_sum(acc);
    if (lane == 0) y[r] += alpha * acc;
}

// y[c] += alpha * sum_r P[r][c] x[r]   (CTA = 128 rows x 256 columns, atomicAdd per column)
__global__ void __launch_bounds__(256) gemv_t_kernel(const double *__restrict__ P, int64_t ldp, int64_t rows, int cols,
                                                     const double *__restrict__ x, double *__restrict__ y,
                                                     double alpha) {
    const int c = blockIdx.x * 256 + threadIdx.x;
    const int64_t r0 = (int64_t)blockIdx.y * 128;
    __shared__ double xs[128];
    if (threadIdx.x < 128) xs[threadIdx.x] = (r0 + threadIdx.x < rows) ? x[r0 + threadIdx.x] : 0.0;
    __syncthreads();
    if (c >= cols) return;
    const int nr = (int)((rows - r0 < 128) ? rows - r0 : 128);
    const double *p = P + r0 * ldp + c;
    double acc = 0.0;
#pragma unroll 8
    for (int r = 0; r < nr; r++) acc += p[(int64_t)r * ldp] * xs[r];
    atomicAdd(y + c, alpha * acc);
}

__global__ void copy2d_kernel(const double *__restrict__ src, int64_t lds, double *__restrict__ dst, int64_t ldd,
                              int64_t rows, int cols) {
    int c2 = blockIdx.x * blockDim.x + threadIdx.x;  // pairs of columns
    int64_t r = blockIdx.y + (int64_t)blockIdx.z * 65535;
    if (r >= rows || 2 * c2 >= cols) return;
    if (2 * c2 + 1 < cols)
        *reinterpret_cast<double2 *>(dst + r * ldd + 2 * c2) = *reinterpret_cast<const double2 *>(src + r * lds + 2 * c2);
    else
        dst[r * ldd + 2 * c2] = src[r * lds + 2 * c2];
}

// copy to several destinations at once (peer-mapped buffers or one multicast address): see GemmMirror
__global__ void copy2d_bcast_kernel(const double *__restrict__ src, int64_t lds, GemmMirror mir, int64_t rows, int cols) {
    int c2 = blockIdx.x * blockDim.x + threadIdx.x;  // pairs of columns
    int64_t r = blockIdx.y + (int64_t)blockIdx.z * 65535;
    if (r >= rows || 2 * c2 >= cols) return;
    const int64_t off = r * mir.ld + 2 * c2;
    if (2 * c2 + 1 < cols) {
        const double2 v = *reinterpret_cast<const double2 *>(src + r * lds + 2 * c2);
        gemm_mirror_store2(mir, off, v.x, v.y);
    } else {
        gemm_mirror_store(mir, off, src[r * lds + 2 * c2]);
    }
}

}  // namespace lgp

using namespace lgp;

extern "C" {

int lgp_dist_local_shape(const lgp_grid_t *grid, int64_t *rows, int64_t *cols) {
    Grid g;
    if (!grid_ok(grid, g) || !rows || !cols) return LGP_ERR_BADARG;
    *rows = (int64_t)g.LR * g.T;
    *cols = (int64_t)g.LC * g.T;
    return LGP_OK;
}

int lgp_dist_diag(lgp_stream_t stream, const lgp_grid_t *grid, const double *A, int64_t lda, double *d) {
    Grid g;
    if (!grid_ok(grid, g) || !A || !d) return LGP_ERR_BADARG;
    dist_diag_kernel<<<(unsigned)((g.n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(A, lda, g, d);
    LGP_CUDA_CHECK_LAUNCH();
    return LGP_OK;
}

int lgp_dist_scale_from_diag(lgp_stream_t stream, const double *d, int64_t n, int64_t npad, double *s, double *sinv) {
    if (n < 1 || npad < n || !d || !s || !sinv) return LGP_ERR_BADARG;
    dist_scale_kernel<<<(unsigned)((npad + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d, n, npad, s, sinv);
    LGP_CUDA_CHECK_LAUNCH();
    return LGP_OK;
}

int lgp_dist_prepare(lgp_stream_t stream, const lgp_grid_t *grid, double *A, int64_t lda, const double *sinv,
                     double *rowsum) {
    Grid g;
    if (!grid_ok(grid, g) || !A || !sinv || !rowsum) return LGP_ERR_BADARG;
    const int64_t rows = (int64_t)g.LR * g.T;
    if (rows == 0 || g.LC == 0) return LGP_OK;
    dist_prepare_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(A, lda, g, sinv, rowsum);
    LGP_CUDA_CHECK_LAUNCH();
    return LGP_OK;
}

int lgp_dist_eps(lgp_stream_t stream, const double *rowsum, int64_t n, double epsrel, double epsabs, double *out) {
    if (n < 1 || !rowsum || !out) return LGP_ERR_BADARG;
    if (epsrel < 0) epsrel = (double)n * 2.220446049250313e-16;
    dist_eps_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(rowsum, n, epsrel, epsabs, out);
    LGP_CUDA_CHECK_LAUNCH();
    return LGP_OK;
}

int lgp_dist_add_diag(lgp_stream_t stream, const lgp_grid_t *grid, double *A, int64_t lda, const double *eps) {
    Grid g;
    if (!grid_ok(grid, g) || !A || !eps) return LGP_ERR_BADARG;
    dist_add_diag_kernel<<<(unsigned)((g.n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(A, lda, g, eps);
    LGP_CUDA_CHECK_LAUNCH();
    return LGP_OK;
}

int lgp_dist_trailing_update(lgp_stream_t stream, const lgp_grid_t *grid, double *A, int64_t lda, int64_t k,
                             const double *const *panel, int64_t lj_begin, int64_t lj_end) {
    Grid g;
    if (!grid_ok(grid, g) || !A || !panel || k < 0 || k >= g.NT) return LGP_ERR_BADARG;
    if (lj_begin < 0) lj_begin = 0;
    if (lj_end > g.LC) lj_end = g.LC;
    const int64_t TT = (int64_t)g.T * g.T;
    const int li0_mine = tiles_before((int)k + 1, g.pr, g.Pr);
    for (int64_t lj = lj_begin; lj < lj_end; lj++) {
        const int J = g.pc + g.Pc * (int)lj;
        if (J <= k) continue;
        const int li_s = tiles_before(J, g.pr, g.Pr);  // first local tile row with I >= J
        if (li_s >= g.LR) continue;
        const int rJ = J % g.Pr;
        const double *Aop = panel[g.pr] + (int64_t)(li_s - li0_mine) * TT;
        const double *Bop = panel[rJ] + (int64_t)(J / g.Pr - tiles_before((int)k + 1, rJ, g.Pr)) * TT;
        double *C = A + (int64_t)li_s * g.T * lda + lj * g.T;
        int rc = gemm_launch((cudaStream_t)stream, true, true, (g.LR - li_s) * g.T, g.T, g.T, -1.0, Aop, g.T, Bop, g.T,
                             C, lda, 0);
        if (rc) return rc;
    }
    return LGP_OK;
}

int lgp_dgemv(lgp_stream_t stream, int trans, const double *P, int64_t ldp, int64_t rows, int64_t cols,
              const double *x, double *y, double alpha) {
    if (rows < 0 || cols < 0 || cols > (1 << 30) || !P || !x || !y) return LGP_ERR_BADARG;
    if (rows == 0 || cols == 0) return LGP_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (!trans) {
        // x is staged in shared memory: wide matrices go in column chunks (y accumulates)
        constexpr int64_t CH = 4096;
        for (int64_t c0 = 0; c0 < cols; c0 += CH) {
            const int64_t cc = cols - c0 < CH ? cols - c0 : CH;
            gemv_n_kernel<<<(unsigned)((rows + 7) / 8), 256, (size_t)cc * 8, st>>>(P + c0, ldp, rows, (int)cc, x + c0, y,
                                                                                  alpha);
            LGP_CUDA_CHECK_LAUNCH();
        }
        return LGP_OK;
    } else {
        if (cols > (1 << 20)) return LGP_ERR_UNSUPPORTED;
        dim3 grid((unsigned)((cols + 255) / 256), (unsigned)((rows + 127) / 128));
        if (grid.y > 65535) return LGP_ERR_UNSUPPORTED;
        gemv_t_kernel<<<grid, 256, 0, st>>>(P, ldp, rows, (int)cols, x, y, alpha);
    }
    LGP_CUDA_CHECK_LAUNCH();
    return LGP_OK;
}

int lgp_copy2d(lgp_stream_t stream, const double *src, int64_t lds, double *dst, int64_t ldd, int64_t rows,
               int64_t cols) {
    if (rows < 0 || cols < 0 || !src || !dst) return LGP_ERR_BADARG;
    if (rows == 0 || cols == 0) return LGP_OK;
    if ((lds & 1) || (ldd & 1) || (reinterpret_cast<uintptr_t>(src) & 15) || (reinterpret_cast<uintptr_t>(dst) & 15))
        return LGP_ERR_ALIGN;
    dim3 grid((unsigned)((cols / 2 + 1 + 127) / 128), (unsigned)(rows > 65535 ? 65535 : rows),
              (unsigned)((rows + 65534) / 65535));
    copy2d_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(src, lds, dst, ldd, rows, (int)cols);
    LGP_CUDA_CHECK_LAUNCH();
    return LGP_OK;
}

int lgp_copy2d_bcast(lgp_stream_t stream, const double *src, int64_t lds, int64_t rows, int64_t cols, int n_dst,
                     void *const *dst, int64_t ld_dst, int multimem) {
    if (rows < 0 || cols < 0 || !src || n_dst < 1 || n_dst > GEMM_MAX_MIRRORS || !dst || ld_dst < cols ||
        (multimem && n_dst != 1))
        return LGP_ERR_BADARG;
    if (rows == 0 || cols == 0) return LGP_OK;
    if ((lds & 1) || (ld_dst & 1) || (reinterpret_cast<uintptr_t>(src) & 15)) return LGP_ERR_ALIGN;
    GemmMirror mir;
    mir.n = n_dst;
    mir.multimem = multimem ? 1 : 0;
    mir.ld = ld_dst;
    for (int i = 0; i < n_dst; i++) {
        if (!dst[i] || (reinterpret_cast<uintptr_t>(dst[i]) & 15)) return LGP_ERR_ALIGN;
        mir.dst[i] = (double *)dst[i];
    }
    dim3 grid((unsigned)((cols / 2 + 1 + 127) / 128), (unsigned)(rows > 65535 ? 65535 : rows),
              (unsigned)((rows + 65534) / 65535));
    copy2d_bcast_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(src, lds, mir, rows, (int)cols);
    LGP_CUDA_CHECK_LAUNCH();
    return LGP_OK;
}

// ---- lower-packed panels (see PanelGeom)
static bool panel_geom(const lgp_grid_t *grid, int64_t lj, Grid &g, PanelGeom &p) {
    if (!grid_ok(grid, g) || lj < 0 || lj >= g.LC) return false;
    p.n = g.n;
    p.T = g.T;
    p.Pr = g.Pr;
    p.pr = g.pr;
    p.J = g.pc + g.Pc * (int)lj;
    p.first = tiles_before(p.J, g.pr, g.Pr);
    p.rows_t = g.LR - p.first;
    return true;
}

int lgp_dist_panel_rows(const lgp_grid_t *grid, int64_t lj, int64_t *first_tile_row, int64_t *tile_rows) {
    Grid g;
    PanelGeom p;
    if (!panel_geom(grid, lj, g, p) || !first_tile_row || !tile_rows) return LGP_ERR_BADARG;
    *first_tile_row = p.first;
    *tile_rows = p.rows_t > 0 ? p.rows_t : 0;
    return LGP_OK;
}

int lgp_dist_panel_diag(lgp_stream_t stream, const lgp_grid_t *grid, int64_t lj, const double *panel, double *d) {
    Grid g;
    PanelGeom p;
    if (!panel_geom(grid, lj, g, p) || !d) return LGP_ERR_BADARG;
    if (p.rows_t <= 0 || p.J % g.Pr != g.pr) return LGP_OK;  // the diagonal tile (J, J) is not stored here
    if (!panel) return LGP_ERR_BADARG;
    panel_diag_kernel<<<(g.T + 255) / 256, 256, 0, (cudaStream_t)stream>>>(panel, p, d);
    LGP_CUDA_CHECK_LAUNCH();
    return LGP_OK;
}

int lgp_dist_panel_prepare(lgp_stream_t stream, const lgp_grid_t *grid, int64_t lj, double *panel, const double *sinv,
                           double *rowsum) {
    Grid g;
    PanelGeom p;
    if (!panel_geom(grid, lj, g, p) || !sinv || !rowsum) return LGP_ERR_BADARG;
    if (p.rows_t <= 0) return LGP_OK;
    if (!panel) return LGP_ERR_BADARG;
    const int64_t nrows = (int64_t)p.rows_t * p.T;
    panel_prepare_kernel<<<(unsigned)((nrows + PP_ROWS - 1) / PP_ROWS), 256, 0, (cudaStream_t)stream>>>(panel, p, sinv,
                                                                                                      rowsum);
    LGP_CUDA_CHECK_LAUNCH();
    return LGP_OK;
}

int lgp_dist_panel_add_diag(lgp_stream_t stream, const lgp_grid_t *grid, int64_t lj, double *panel, const double *eps) {
    Grid g;
    PanelGeom p;
    if (!panel_geom(grid, lj, g, p) || !eps) return LGP_ERR_BADARG;
    if (p.rows_t <= 0 || p.J % g.Pr != g.pr) return LGP_OK;
    if (!panel) return LGP_ERR_BADARG;
    panel_add_diag_kernel<<<(g.T + 255) / 256, 256, 0, (cudaStream_t)stream>>>(panel, p, eps);
    LGP_CUDA_CHECK_LAUNCH();
    return LGP_OK;
}

int lgp_dist_trailing_update_packed(lgp_stream_t stream, const lgp_grid_t *grid, double *const *colpanels, int64_t k,
                                    const double *const *panel, int64_t lj_begin, int64_t lj_end) {
    Grid g;
    if (!grid_ok(grid, g) || !colpanels || !panel || k < 0 || k >= g.NT) return LGP_ERR_BADARG;
    if (lj_begin < 0) lj_begin = 0;
    if (lj_end > g.LC) lj_end = g.LC;
    const int64_t TT = (int64_t)g.T * g.T;
    const int li0_mine = tiles_before((int)k + 1, g.pr, g.Pr);
    for (int64_t lj = lj_begin; lj < lj_end; lj++) {
        const int J = g.pc + g.Pc * (int)lj;
        if (J <= k) continue;
        const int li_s = tiles_before(J, g.pr, g.Pr);  // first local tile row with I >= J = first stored row of the panel
        if (li_s >= g.LR) continue;
        if (!colpanels[lj]) return LGP_ERR_BADARG;
        const int rJ = J % g.Pr;
        const double *Aop = panel[g.pr] + (int64_t)(li_s - li0_mine) * TT;
        const double *Bop = panel[rJ] + (int64_t)(J / g.Pr - tiles_before((int)k + 1, rJ, g.Pr)) * TT;
        int rc = gemm_launch((cudaStream_t)stream, true, true, (g.LR - li_s) * g.T, g.T, g.T, -1.0, Aop, g.T, Bop, g.T,
                             colpanels[lj], g.T, 0);
        if (rc) return rc;
    }
    return LGP_OK;
}

}  // extern "C"
