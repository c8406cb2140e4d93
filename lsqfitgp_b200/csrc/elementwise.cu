// Small HBM-streaming kernels of the decomposition / assembly layer that used to be eager PyTorch passes:
//   lgp_sym_expand_sub  scale * (sym(lower triangle) - b b^T) as a full matrix      (_linalg/_decomp.py:505-509)
//   lgp_symlower_dot    sum_ij (invK_ij - b_i b_j) D_ij from the lower triangle      (_linalg/_decomp.py:524-531, gradfwd)
//   lgp_colsumsq        column sums of squares                                       (_linalg/_decomp.py:422-427, ginv_diagquad)
//   lgp_add_scalar      Y += c                                                       (_GP/_elements.py:581-601, scalar blocks)
//   lgp_searchsorted    bin indices of coordinates w.r.t. splitting points           (_kernels/_bart.py:294-299,503-514)
#include "../../include/lgp_b200.h"
#include "common.cuh"
#include "internal.h"

namespace lgp {

constexpr int ET = 32;  // tile edge; 32 x 8 threads, 4 rows each

// tile index (I >= J) of the t-th lower tile
__device__ __forceinline__ void lower_tile(int64_t t, int &I, int &J) {
    I = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
    while ((int64_t)(I + 1) * (I + 2) / 2 <= t) I++;
    while ((int64_t)I * (I + 1) / 2 > t) I--;
    J = (int)(t - (int64_t)I * (I + 1) / 2);
}

__global__ void __launch_bounds__(256) sym_expand_sub_kernel(const double *__restrict__ low, int64_t ldl,
                                                             const double *__restrict__ b, int n, double scale,
                                                             double *__restrict__ out, int64_t ldo) {
    __shared__ double tile[ET][ET + 1];
    int I, J;
    lower_tile(blockIdx.x, I, J);
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int i0 = I * ET, j0 = J * ET;
#pragma unroll
    for (int r = ty; r < ET; r += 8) {
        const int i = i0 + r, j = j0 + tx;
        double v = 0.0;
        if (i < n && j < n && j <= i) {
            v = low[(int64_t)i * ldl + j];
            if (b) v -= b[i] * b[j];
            v *= scale;
        }
        tile[r][tx] = v;
    }
    __syncthreads();
    if (I == J) {
        // diagonal tile: mirror the lower part inside the tile
#pragma unroll
        for (int r = ty; r < ET; r += 8) {
            const int i = i0 + r, j = j0 + tx;
            if (i < n && j < n) out[(int64_t)i * ldo + j] = (tx <= r) ? tile[r][tx] : tile[tx][r];
        }
        return;
    }
#pragma unroll
    for (int r = ty; r < ET; r += 8) {
        const int i = i0 + r, j = j0 + tx;
        if (i < n && j < n) out[(int64_t)i * ldo + j] = tile[r][tx];
        // transposed tile: row (j0 + r), columns i0 + tx
        const int jj = j0 + r, ii = i0 + tx;
        if (jj < n && ii < n) out[(int64_t)jj * ldo + ii] = tile[tx][r];
    }
}

__global__ void __launch_bounds__(256) symlower_dot_kernel(const double *__restrict__ low, int64_t ldl,
                                                           const double *__restrict__ b, const double *__restrict__ D,
                                                           int64_t ldd, int n, int64_t ntiles, double *__restrict__ out) {
    __shared__ double tileT[ET][ET + 1];
    __shared__ double red[8];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    double acc = 0.0;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        int I, J;
        lower_tile(t, I, J);
        const int i0 = I * ET, j0 = J * ET;
        __syncthreads();
        // D[J-tile rows][I-tile cols] -> tileT[r][c] = D[j0 + r][i0 + c]
#pragma unroll
        for (int r = ty; r < ET; r += 8) {
            const int jj = j0 + r, ii = i0 + tx;
            tileT[r][tx] = (jj < n && ii < n) ? D[(int64_t)jj * ldd + ii] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int r = ty; r < ET; r += 8) {
            const int i = i0 + r, j = j0 + tx;
            if (i < n && j < n && j <= i) {
                double g = low[(int64_t)i * ldl + j];
                if (b) g -= b[i] * b[j];
                const double dij = D[(int64_t)i * ldd + j];
                acc += (j < i) ? g * (dij + tileT[tx][r]) : g * dij;
            }
        }
    }
    acc = warp_sum(acc);
    __syncthreads();
    if (tx == 0) red[ty] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < 8; w++) s += red[w];
        atomicAdd(out, s);
    }
}

// one CTA per 32 columns: rows strided over the 8 warps... lanes along columns (coalesced), 8 row groups
__global__ void __launch_bounds__(256) colsumsq_kernel(const double *__restrict__ A, int64_t lda, int64_t rows, int cols,
                                                       double *__restrict__ out) {
    __shared__ double red[8][32];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + tx;
    double acc = 0.0;
    if (c < cols) {
        const int64_t r0 = (int64_t)blockIdx.y * 8 + ty, step = (int64_t)gridDim.y * 8;
        for (int64_t r = r0; r < rows; r += step) {
            const double v = A[r * lda + c];
            acc += v * v;
        }
    }
    red[ty][tx] = acc;
    __syncthreads();
    if (ty == 0 && c < cols) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < 8; w++) s += red[w][tx];
        atomicAdd(out + c, s);
    }
}

__global__ void add_scalar_kernel(int64_t n, int64_t m, double *__restrict__ Y, int64_t ldy, double c) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t i = blockIdx.y + (int64_t)65535 * blockIdx.z;
    if (i < n && j < m) Y[i * ldy + j] += c;
}

__global__ void searchsorted_kernel(const double *__restrict__ splits, int maxlen, int p, const double *__restrict__ x,
                                    int64_t ldx, int64_t n, int32_t *__restrict__ out, int64_t ldo) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int d = blockIdx.y;
    if (i >= n) return;
    const double v = x[(int64_t)d * ldx + i];
    // first index k in [0, maxlen] with splits[k][d] >= v  (numpy / jax searchsorted, side='left'); NaN sorts last
    int lo = 0, hi = maxlen;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        const double s = splits[(int64_t)mid * p + d];
        if (s < v || (v != v && s == s))
            lo = mid + 1;
        else
            hi = mid;
    }
    out[(int64_t)d * ldo + i] = lo;
}

}  // namespace lgp

using namespace lgp;

extern "C" {

int lgp_sym_expand_sub(lgp_stream_t stream, const double *low, int64_t ldl, const double *b, int64_t n, double scale,
                       double *out, int64_t ldo) {
    if (n < 1 || n > (1 << 30) || !low || !out || ldl < n || ldo < n) return LGP_ERR_BADARG;
    const int64_t nt = (n + ET - 1) / ET;
    const int64_t tiles = nt * (nt + 1) / 2;
    if (tiles > 2147483647LL) return LGP_ERR_UNSUPPORTED;
    sym_expand_sub_kernel<<<(unsigned)tiles, 256, 0, (cudaStream_t)stream>>>(low, ldl, b, (int)n, scale, out, ldo);
    LGP_CUDA_CHECK_LAUNCH();
    return LGP_OK;
}

int lgp_symlower_dot(lgp_stream_t stream, const double *low, int64_t ldl, const double *b, const double *D, int64_t ldd,
                     int64_t n, double *out) {
    if (n < 1 || n > (1 << 30) || !low || !D || !out || ldl < n || ldd < n) return LGP_ERR_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (cudaMemsetAsync(out, 0, sizeof(double), st) != cudaSuccess) return LGP_ERR_CUDA;
    const int64_t nt = (n + ET - 1) / ET;
    const int64_t tiles = nt * (nt + 1) / 2;
    const unsigned grid = (unsigned)(tiles < 148 * 16 ? tiles : 148 * 16);
    symlower_dot_kernel<<<grid, 256, 0, st>>>(low, ldl, b, D, ldd, (int)n, tiles, out);
    LGP_CUDA_CHECK_LAUNCH();
    return LGP_OK;
}

int lgp_colsumsq(lgp_stream_t stream, const double *A, int64_t lda, int64_t rows, int64_t cols, double *out) {
    if (rows < 0 || cols < 0 || cols > (1 << 30) || !out || (rows && cols && (!A || lda < cols))) return LGP_ERR_BADARG;
    if (cols == 0) return LGP_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (cudaMemsetAsync(out, 0, sizeof(double) * cols, st) != cudaSuccess) return LGP_ERR_CUDA;
    if (rows == 0) return LGP_OK;
    const unsigned gx = (unsigned)((cols + 31) / 32);
    // enough row groups to fill the GPU when the matrix is tall and narrow
    int64_t gy = (148 * 8 + gx - 1) / gx;
    const int64_t maxgy = (rows + 63) / 64;
    if (gy > maxgy) gy = maxgy;
    if (gy < 1) gy = 1;
    if (gy > 65535) gy = 65535;
    colsumsq_kernel<<<dim3(gx, (unsigned)gy), 256, 0, st>>>(A, lda, rows, (int)cols, out);
    LGP_CUDA_CHECK_LAUNCH();
    return LGP_OK;
}

int lgp_add_scalar(lgp_stream_t stream, int64_t n, int64_t m, double *Y, int64_t ldy, double c) {
    if (n < 0 || m < 0 || !Y || ldy < m) return LGP_ERR_BADARG;
    if (n == 0 || m == 0) return LGP_OK;
    dim3 g((unsigned)((m + 255) / 256), (unsigned)(n < 65535 ? n : 65535), (unsigned)((n + 65534) / 65535));
    add_scalar_kernel<<<g, 256, 0, (cudaStream_t)stream>>>(n, m, Y, ldy, c);
    LGP_CUDA_CHECK_LAUNCH();
    return LGP_OK;
}

int lgp_searchsorted(lgp_stream_t stream, const double *splits, int64_t maxlen, int p, const double *x, int64_t ldx,
                     int64_t n, int32_t *out, int64_t ldo) {
    if (maxlen < 0 || maxlen > 2147483647LL || p < 0 || p > 65535 || n < 0 || !out || (maxlen && p && !splits) ||
        (n && p && !x))
        return LGP_ERR_BADARG;
    if (n == 0 || p == 0) return LGP_OK;
    searchsorted_kernel<<<dim3((unsigned)((n + 255) / 256), (unsigned)p), 256, 0, (cudaStream_t)stream>>>(
        splits, (int)maxlen, p, x, ldx, n, out, ldo);
    LGP_CUDA_CHECK_LAUNCH();
    return LGP_OK;
}

}  // extern "C"
