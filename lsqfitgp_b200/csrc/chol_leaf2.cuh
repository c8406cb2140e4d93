// 128x128 Cholesky leaf, version 2: Cholesky factor AND its inverse of one diagonal block in one CTA, blocked 4 x 32.
//
// Same contract as potrf_leaf_kernel (chol.cu): reads the lower triangle of Wblk, writes L into the lower triangle
// (zeros above), X = L^-1 into invd (row-major 128 x 128, zeros above the diagonal), dvec[j0 + k] = L_kk, and
// atomicMin(info, j0 + k + 1) at the first non-positive / non-finite pivot (NaNs then propagate like a failed
// LAPACK/JAX factorisation, reference _linalg/_decomp.py:388-391).
//
// Version 1 is unblocked: 128 block-wide steps, each a rank-1 update of the whole register-resident matrix (177
// instructions per warp and step, IPC 0.33 with two warps per scheduler: 68 us).  Here the matrix lives in shared memory
// and the work is split the LAPACK way, so that only 32-column factorisations sit on the critical path:
//   for J = 0..3:   potrf of the 32 x 32 diagonal block by ONE warp, rows in registers, pivots by shuffle
//                   inverse of that block (warp 0) || triangular solve of the 32-row blocks below it (warps 1..3), one
//                     matrix row per lane, the factor's entries broadcast from shared memory
//                   rank-32 update of the trailing blocks on the FP64 tensor pipe (DMMA 8x8x4 from shared memory)
//   off-diagonal blocks of X by block distance d = 1, 2, 3:  X_IJ = -X_II (sum_K L_IK X_KJ), DMMA again.
#pragma once
#include "common.cuh"

namespace lgp {

constexpr int L2_B = 32;                    // sub-block
constexpr int L2_S = 133;                   // row stride of the matrix in shared memory (odd: one row per lane is
                                            // conflict-free; 5 lr + q spreads the DMMA fragment loads over the banks)
constexpr int L2_XS = 33;                   // row stride of the 32 x 32 scratch blocks
constexpr int L2_THREADS = 256;
// S[128][133] | XD[4][32][33] (inverted diagonal blocks) | T[3][32][33] (product scratch) | xd[128] (1 / L_kk) | col[2][32]
constexpr int L2_SMEM_DOUBLES = 128 * L2_S + 4 * L2_B * L2_XS + 3 * L2_B * L2_XS + 128 + 2 * L2_B;
constexpr int L2_SMEM_BYTES = L2_SMEM_DOUBLES * 8;

// sqrt(pv) and 1/sqrt(pv) to ~1 ulp from the hardware estimate (MUFU.RSQ64H, ~2^-20): two Goldschmidt iterations and a
// Newton correction of the root with the exact residual (cf. fastmath.cuh); ~12 dependent DP instructions instead of the
// library rsqrt's ~30 with special-case branches.  Outside the safe range (and for pv <= 0, NaN) the library functions
// keep the IEEE special values that the failure reporting relies on.  pv is warp-uniform: no divergence.
__device__ __forceinline__ void l2_sqrt_rsqrt(double pv, double &l, double &rl) {
    if (pv >= 1e-280 && pv <= 1e280) {
        double y0;
        asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(pv));
        double g = pv * y0, h = 0.5 * y0;
        double r = fma(-h, g, 0.5);
        g = fma(g, r, g);
        h = fma(h, r, h);
        r = fma(-h, g, 0.5);
        g = fma(g, r, g);
        h = fma(h, r, h);
        l = fma(fma(-g, g, pv), h, g);
        rl = h + h;
        rl = fma(fma(-l, rl, 1.0), rl, rl);  // one Newton step on the reciprocal of the final root
    } else {
        rl = rsqrt(pv);  // NaN for pv < 0, inf for pv == 0
        l = pv * rl;
    }
}

// ---- 32 x 32 Cholesky by one warp: lane r holds row r; column j of L is published through shared memory (one
// conflict-free store, then broadcast loads: the shuffle unit would need two SHFL per entry) -------------------------
__device__ __forceinline__ void l2_potrf32(double *__restrict__ Sd /* diag block in S */, double *__restrict__ xd,
                                           double *__restrict__ col /* 2 x 32 scratch */, double *__restrict__ dvec,
                                           int32_t *__restrict__ info, int jglob, int lane) {
    double a[L2_B];
#pragma unroll
    for (int c = 0; c < L2_B; c++) a[c] = Sd[lane * L2_S + c];
#pragma unroll
    for (int j = 0; j < L2_B; j++) {
        const double pv = __shfl_sync(0xffffffffu, a[j], j);
        double l, rl;
        l2_sqrt_rsqrt(pv, l, rl);
        if (lane == j) {
            if (!(pv > 0.0) || !(l < INFINITY)) atomicMin(info, jglob + j + 1);
            dvec[jglob + j] = l;
            xd[j] = rl;
        }
        a[j] = (lane == j) ? l : a[j] * rl;  // lanes r > j: L_rj; lanes r < j hold garbage that is never read
        const double v = a[j];
        double *cb = col + (j & 1) * L2_B;
        cb[lane] = v;
        __syncwarp();
#pragma unroll
        for (int c = j + 1; c < L2_B; c++) a[c] = fma(-v, cb[c], a[c]);  // (meaningful for lanes r >= c only)
    }
#pragma unroll
    for (int c = 0; c < L2_B; c++)
        if (c <= lane) Sd[lane * L2_S + c] = a[c];
}

// ---- inverse of the 32 x 32 lower-triangular diagonal block: lane c computes column c of X (right-looking: the
// updates of one step are independent FMAs) ---------------------------------------------------------------------
__device__ __forceinline__ void l2_trtri32(const double *__restrict__ Sd, const double *__restrict__ xd,
                                           double *__restrict__ XDb /* [32][33] */, int lane) {
    double x[L2_B];  // running sums  s_r = sum_{k < r} L_rk x_k, then the solution
#pragma unroll
    for (int r = 0; r < L2_B; r++) x[r] = 0.0;
#pragma unroll
    for (int k = 0; k < L2_B; k++) {
        const double rl = xd[k];
        const double xk = (k == lane) ? rl : ((k > lane) ? -rl * x[k] : 0.0);
        x[k] = xk;
#pragma unroll
        for (int r = k + 1; r < L2_B; r++) x[r] = fma(Sd[r * L2_S + k], xk, x[r]);  // broadcast loads
    }
#pragma unroll
    for (int r = 0; r < L2_B; r++) XDb[r * L2_XS + lane] = x[r];
}

// ---- P <- P L^-T for a 32-row block below the diagonal block: lane r holds row r of P ---------------------------------
__device__ __forceinline__ void l2_trsm32(double *__restrict__ Sp /* panel block in S */, const double *__restrict__ Sd,
                                          const double *__restrict__ xd, int lane) {
    double p[L2_B];
#pragma unroll
    for (int c = 0; c < L2_B; c++) p[c] = Sp[lane * L2_S + c];
#pragma unroll
    for (int j = 0; j < L2_B; j++) {
        p[j] *= xd[j];
#pragma unroll
        for (int c = j + 1; c < L2_B; c++) p[c] = fma(-p[j], Sd[c * L2_S + j], p[c]);
    }
#pragma unroll
    for (int c = 0; c < L2_B; c++) Sp[lane * L2_S + c] = p[c];
}

// ---- 32 x (8 NJT) x 32 product on the FP64 tensor pipe, operands in shared memory --------------------------------------
// acc (4 x NJT DMMA tiles, lane holds [row lr][cols 2q, 2q+1] of each) += sum_k A(r, k) B(c, k),
// A(r, k) = pa[r * sar + k * sak], B(c, k) = pb[c * sbr + k * sbk]   (pb already offset to the unit's first column)
template <int NJT>
__device__ __forceinline__ void l2_mma(double (&acc)[4][NJT][2], const double *__restrict__ pa, int sar, int sak,
                                       const double *__restrict__ pb, int sbr, int sbk, int lane) {
    const int lr = lane >> 2, q = lane & 3;
#pragma unroll
    for (int k0 = 0; k0 < L2_B; k0 += 4) {
        double af[4], bf[NJT];
#pragma unroll
        for (int i = 0; i < 4; i++) af[i] = pa[(8 * i + lr) * sar + (k0 + q) * sak];
#pragma unroll
        for (int j = 0; j < NJT; j++) bf[j] = pb[(8 * j + lr) * sbr + (k0 + q) * sbk];
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < NJT; j++) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
}

template <int NJT>
__device__ __forceinline__ void l2_zero(double (&acc)[4][NJT][2]) {
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < NJT; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
}

// C(r, c) = pc[r * scr + c * scc] (pc offset to the unit's first column):  C = (accumulate ? C : 0) + alpha acc
template <int NJT>
__device__ __forceinline__ void l2_store(const double (&acc)[4][NJT][2], double *__restrict__ pc, int scr, int scc,
                                         double alpha, bool accumulate, int lane) {
    const int lr = lane >> 2, q = lane & 3;
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < NJT; j++)
#pragma unroll
            for (int e = 0; e < 2; e++) {
                double *p = pc + (8 * i + lr) * scr + (8 * j + 2 * q + e) * scc;
                const double v = alpha * acc[i][j][e];
                *p = accumulate ? *p + v : v;
            }
}

__device__ long long l2_dbg[32];  // phase timestamps of the last launch (clock64 of thread 0), read by lgp_debug_leaf2_clocks
#define L2_STAMP(i)                          \
    do {                                     \
        if (tid == 0) l2_dbg[i] = clock64(); \
    } while (0)

__global__ void __launch_bounds__(L2_THREADS, 1) potrf_leaf2_kernel(double *__restrict__ Wblk, int64_t ld,
                                                                   double *__restrict__ invd,
                                                                   double *__restrict__ dvec,
                                                                   int32_t *__restrict__ info, int j0) {
    extern __shared__ __align__(16) double l2sm[];
    double *S = l2sm;
    double *XD = S + 128 * L2_S;
    double *T = XD + 4 * L2_B * L2_XS;
    double *xd = T + 3 * L2_B * L2_XS;
    double *colb = xd + 128;
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
    constexpr int NW = L2_THREADS / 32;

    // lower triangle in; the strict upper triangle of S will hold the off-diagonal blocks of X, transposed
    for (int idx = tid; idx < NB * NB; idx += L2_THREADS) {
        const int r = idx >> 7, c = idx & 127;
        S[r * L2_S + c] = (r >= c) ? Wblk[(int64_t)r * ld + c] : 0.0;
    }
    L2_STAMP(0);
    __syncthreads();
    L2_STAMP(1);

    for (int J = 0; J < 4; J++) {
        double *Sd = S + (J * L2_B) * L2_S + J * L2_B;
        if (w == 0) l2_potrf32(Sd, xd + J * L2_B, colb, dvec, info, j0 + J * L2_B, lane);
        __syncthreads();
        L2_STAMP(2 + 3 * J);
        if (w == 0) {
            l2_trtri32(Sd, xd + J * L2_B, XD + J * L2_B * L2_XS, lane);
        } else if (w <= 3 - J) {
            l2_trsm32(S + ((J + w) * L2_B) * L2_S + J * L2_B, Sd, xd + J * L2_B, lane);
        }
        __syncthreads();
        L2_STAMP(3 + 3 * J);
        // trailing blocks (I, K), J < K <= I <= 3:  S_IK -= P_I P_K^T  (full 32 x 32 also on the diagonal blocks: their
        // upper halves are scratch); work unit = 8 columns of a block, units dealt round-robin to the 8 warps
        const int nrem = 3 - J, nunit = 4 * (nrem * (nrem + 1) / 2);
        for (int u = w; u < nunit; u += NW) {
            const int b = u >> 2, cs = (u & 3) * 8;
            int I = 0, K = b;
            while (K > I) {  // row-major enumeration of the lower triangle of blocks
                K -= I + 1;
                I++;
            }
            I += J + 1;
            K += J + 1;
            double acc[4][1][2];
            l2_zero<1>(acc);
            l2_mma<1>(acc, S + (I * L2_B) * L2_S + J * L2_B, L2_S, 1, S + (K * L2_B + cs) * L2_S + J * L2_B, L2_S, 1, lane);
            l2_store<1>(acc, S + (I * L2_B) * L2_S + K * L2_B + cs, L2_S, 1, -1.0, true, lane);
        }
        __syncthreads();
        L2_STAMP(4 + 3 * J);
    }

    // off-diagonal blocks of X = L^-1 by block distance:  X_IJ = -XD_I (sum_{K=J}^{I-1} L_IK X_KJ), with X_JJ = XD_J and
    // X_KJ (K > J) stored transposed at S[J-block rows][K-block columns]; unit = 8 columns of a block
    for (int d = 1; d <= 3; d++) {
        const int nunit = 4 * (4 - d);
        for (int u = w; u < nunit; u += NW) {
            const int J = u >> 2, I = J + d, cs = (u & 3) * 8;
            double acc[4][1][2];
            l2_zero<1>(acc);
            // K = J: B(c, k) = XD_J[k][c]
            l2_mma<1>(acc, S + (I * L2_B) * L2_S + J * L2_B, L2_S, 1, XD + J * L2_B * L2_XS + cs, 1, L2_XS, lane);
            for (int K = J + 1; K < I; K++)  // B(c, k) = X_KJ[k][c] = S[J*32 + c][K*32 + k]
                l2_mma<1>(acc, S + (I * L2_B) * L2_S + K * L2_B, L2_S, 1, S + (J * L2_B + cs) * L2_S + K * L2_B, L2_S, 1,
                          lane);
            l2_store<1>(acc, T + J * L2_B * L2_XS + cs, L2_XS, 1, 1.0, false, lane);
        }
        __syncthreads();
        for (int u = w; u < nunit; u += NW) {
            const int J = u >> 2, I = J + d, cs = (u & 3) * 8;
            double acc[4][1][2];
            l2_zero<1>(acc);
            // X_IJ = -XD_I T_J:  A(r, k) = XD_I[r][k], B(c, k) = T_J[k][c]
            l2_mma<1>(acc, XD + I * L2_B * L2_XS, L2_XS, 1, T + J * L2_B * L2_XS + cs, 1, L2_XS, lane);
            // store transposed: X_IJ[r][c] -> S[J*32 + c][I*32 + r]
            l2_store<1>(acc, S + (J * L2_B + cs) * L2_S + I * L2_B, 1, L2_S, -1.0, false, lane);
        }
        __syncthreads();
        L2_STAMP(13 + d);
    }

    for (int idx = tid; idx < NB * NB; idx += L2_THREADS) {
        const int r = idx >> 7, c = idx & 127;
        Wblk[(int64_t)r * ld + c] = (r >= c) ? S[r * L2_S + c] : 0.0;
        double xv = 0.0;
        if (r >= c) {
            const int I = r >> 5, Jb = c >> 5;
            xv = (I == Jb) ? XD[I * L2_B * L2_XS + (r & 31) * L2_XS + (c & 31)] : S[c * L2_S + r];
        }
        invd[idx] = xv;
    }
    L2_STAMP(17);
}

}  // namespace lgp
