// BART prior correlation Gram matrix (fast path with <= 3 levels per bracket).
//
// Reference: BART._correlation, src/lsqfitgp/_kernels/_bart.py:628-757 (closed forms for bracket
// widths 1, 2, 3 and the `repeat` scan), called through BART.correlation :415-455 which folds
// `reset` brackets into rows of non-termination probabilities (done by the caller of this ABI).
// Point equality is tested exactly (any(ix != iy)) instead of through fasthash64 (:675-678).
#include <math.h>
#include <string.h>

#include "../../include/lgp_b200.h"
#include "common.cuh"

namespace lgp {

constexpr int BT = 64;           // CTA tile
constexpr int B_THREADS = 256;   // 16 x 16 threads, 4 x 4 pairs each
constexpr int B_MAX_P = 64;
constexpr int B_MAX_ROWS = 16;

struct BartDesc {
    int p;          // number of active (w != 0) covariates after compaction
    int width;      // 1, 2, 3
    int nrows;
    int dim[B_MAX_P];        // original column index of each active covariate
    int n[B_MAX_P];          // split counts
    double w[B_MAX_P];
    double wn[B_MAX_P];         // n ? w/n : 0
    double w_inv_Wn[B_MAX_P];   // w * inv_Wn
    double inv_Wnmod[B_MAX_P];  // 1/(Wn - (n ? w : 0))
    double psin[B_MAX_P];       // digamma(n or 1)
    double rows[B_MAX_ROWS][3];
    double Wn, inv_Wn, gamma, amp;
};

__global__ void __launch_bounds__(B_THREADS) gram_bart_kernel(const __grid_constant__ BartDesc d,
                                                              const double *__restrict__ psi,
                                                              const int32_t *__restrict__ ix, int64_t ldx, int64_t n,
                                                              const int32_t *__restrict__ iy, int64_t ldy, int64_t m,
                                                              double *__restrict__ K, int64_t ldk, int vec_ok) {
    extern __shared__ __align__(16) int32_t bsm[];
    int32_t *sx = bsm;                 // [p][64]
    int32_t *sy = bsm + d.p * BT;      // [p][64]
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int64_t i0 = (int64_t)blockIdx.y * BT, j0 = (int64_t)blockIdx.x * BT;
    for (int idx = tid; idx < d.p * BT; idx += B_THREADS) {
        int k = idx / BT, r = idx % BT;
        int64_t i = i0 + r, j = j0 + r;
        sx[idx] = (i < n) ? ix[(int64_t)d.dim[k] * ldx + i] : 0;
        sy[idx] = (j < m) ? iy[(int64_t)d.dim[k] * ldy + j] : 0;
    }
    __syncthreads();

    double S[4][4], sumi[4][4];
    bool any0[4][4];
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int c = 0; c < 4; c++) {
            S[a][c] = 0.0;
            sumi[a][c] = 0.0;
            any0[a][c] = false;
        }

    const int width = d.width;
    for (int k = 0; k < d.p; k++) {
        const int nk = d.n[k];
        const double wnk = d.wn[k];
        int xi[4], yj[4];
#pragma unroll
        for (int a = 0; a < 4; a++) xi[a] = sx[k * BT + ty + 16 * a];
#pragma unroll
        for (int b = 0; b < 2; b++) {
            int2 t = *reinterpret_cast<const int2 *>(&sy[k * BT + 2 * tx + 32 * b]);
            yj[2 * b] = t.x;
            yj[2 * b + 1] = t.y;
        }
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const int lo = min(xi[a], yj[c]), hi = max(xi[a], yj[c]);
                const int n0 = hi - lo;
                any0[a][c] |= (n0 != 0);
                if (width == 2) {
                    // sum_term = where(n, w/n, 0) @ |ix - iy|        (_bart.py:698-699)
                    S[a][c] = __dadd_rn(S[a][c], __dmul_rn(wnk, (double)n0));
                } else if (width == 3) {
                    // S = wn @ nout, nout = n - n0                    (_bart.py:720,727)
                    S[a][c] = __dadd_rn(S[a][c], __dmul_rn(wnk, (double)(nk - n0)));
                }
            }
    }
    // width 3 needs the complete S before the per-dimension terms: second pass
    if (width == 3) {
        for (int k = 0; k < d.p; k++) {
            const int nk = d.n[k];
            const double wk = d.w[k], wnk = d.wn[k], wiW = d.w_inv_Wn[k], iWmod = d.inv_Wnmod[k], psin = d.psin[k];
            int xi[4], yj[4];
#pragma unroll
            for (int a = 0; a < 4; a++) xi[a] = sx[k * BT + ty + 16 * a];
#pragma unroll
            for (int b = 0; b < 2; b++) {
                int2 t = *reinterpret_cast<const int2 *>(&sy[k * BT + 2 * tx + 32 * b]);
                yj[2 * b] = t.x;
                yj[2 * b + 1] = t.y;
            }
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    const int lo = min(xi[a], yj[c]), hi = max(xi[a], yj[c]);
                    const int n0 = hi - lo;
                    const int nminus0 = hi, nplus0 = nk - lo, nout = nk - n0;
                    const double fn0 = (double)n0;
                    const double inv_Wnminus = nplus0 ? d.inv_Wn : iWmod;
                    const double inv_Wnplus = nminus0 ? d.inv_Wn : iWmod;
                    const double t = __dmul_rn(wnk, fn0);
                    // terms1 = (S + t) * (inv_Wnminus + inv_Wnplus + inv_Wn * (nout - 2))
                    const double terms1 =
                        __dmul_rn(__dadd_rn(S[a][c], t),
                                  __dadd_rn(__dadd_rn(inv_Wnminus, inv_Wnplus), __dmul_rn(d.inv_Wn, (double)(nout - 2))));
                    // terms2
                    const double wiWn0 = __dmul_rn(wiW, fn0);
                    const double wmod = __dmul_rn(wk, iWmod);
                    const double t2a = nplus0 ? __ddiv_rn(wiWn0, (double)nplus0) : wmod;
                    const double t2b = nminus0 ? __ddiv_rn(wiWn0, (double)nminus0) : wmod;
                    const double terms2 = __dadd_rn(t2a, t2b);
                    // terms3 = w * inv_Wn * n0 * (2 psin - psiminus - psiplus)
                    const double psiminus = psi[1 + hi];
                    const double psiplus = psi[1 + nk - lo];
                    const double terms3 =
                        __dmul_rn(wiWn0, __dsub_rn(__dsub_rn(__dmul_rn(2.0, psin), psiminus), psiplus));
                    const double terms = __dsub_rn(__dsub_rn(terms1, terms2), terms3);
                    sumi[a][c] = __dadd_rn(sumi[a][c], __dmul_rn(wnk, terms));
                }
        }
    }

#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int c = 0; c < 4; c++) {
            double g = d.gamma;
            const bool an = any0[a][c];
            for (int r = 0; r < d.nrows; r++) {
                double res;
                if (width == 1) {
                    // 1 - (1 - gamma) * pnt[0]          (_bart.py:688)
                    res = __dsub_rn(1.0, __dmul_rn(__dsub_rn(1.0, g), d.rows[r][0]));
                } else if (width == 2) {
                    // Q = 1 - pnt[1] + gamma * pnt[1]; result = 1 - P0 + Q * (P0 - P0 / Wn * sum_term)   (:702-704)
                    const double P0 = d.rows[r][0], P1 = d.rows[r][1];
                    const double Q = __dadd_rn(__dsub_rn(1.0, P1), __dmul_rn(g, P1));
                    res = __dadd_rn(__dsub_rn(1.0, P0),
                                    __dmul_rn(Q, __dsub_rn(P0, __dmul_rn(__ddiv_rn(P0, d.Wn), S[a][c]))));
                } else {
                    // Q = 1 + pnt[2] * (gamma - 1); sump = S + pnt[1] * (Q * sumi - S);
                    // result = 1 + pnt[0] * (inv_Wn * sump - 1)                                        (:751-753)
                    const double Q = __dadd_rn(1.0, __dmul_rn(d.rows[r][2], __dsub_rn(g, 1.0)));
                    const double sump = __dadd_rn(
                        S[a][c], __dmul_rn(d.rows[r][1], __dsub_rn(__dmul_rn(Q, sumi[a][c]), S[a][c])));
                    res = __dadd_rn(1.0, __dmul_rn(d.rows[r][0], __dsub_rn(__dmul_rn(d.inv_Wn, sump), 1.0)));
                }
                g = an ? res : 1.0;
            }
            S[a][c] = __dmul_rn(d.amp, g);
        }

#pragma unroll
    for (int a = 0; a < 4; a++) {
        int64_t i = i0 + ty + 16 * a;
        if (i >= n) continue;
        double *krow = K + i * ldk;
#pragma unroll
        for (int b = 0; b < 2; b++) {
            int64_t j = j0 + 2 * tx + 32 * b;
            if (j >= m) continue;
            if (vec_ok && j + 1 < m) {
                *reinterpret_cast<double2 *>(krow + j) = make_double2(S[a][2 * b], S[a][2 * b + 1]);
            } else {
                krow[j] = S[a][2 * b];
                if (j + 1 < m) krow[j + 1] = S[a][2 * b + 1];
            }
        }
    }
}

__global__ void fill_kernel(double *K, int64_t ldk, int64_t n, int64_t m, double v) {
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t i = blockIdx.y;
    if (j < m && i < n) K[i * ldk + j] = v;
}

}  // namespace lgp

using namespace lgp;

extern "C" {

// psi_out[k] = digamma(k) for k = 1..len-1 (psi_out[0] = -inf), HOST memory; extended precision recurrence.
int lgp_bart_digamma_table(double *psi_out, int64_t len) {
    if (!psi_out || len < 1) return LGP_ERR_BADARG;
    psi_out[0] = -INFINITY;
    long double v = -0.577215664901532860606512090082402431L;  // digamma(1) = -EulerGamma
    for (int64_t k = 1; k < len; k++) {
        psi_out[k] = (double)v;
        v += 1.0L / (long double)k;
    }
    return LGP_OK;
}

int lgp_gram_bart(lgp_stream_t stream, int p, const int32_t *nsplits, const double *w, const double *rows, int nrows,
                  int width, double gamma, double amp, const double *psi, const int32_t *ix, int64_t ldx, int64_t n,
                  const int32_t *iy, int64_t ldy, int64_t m, double *K_out, int64_t ldk, int flags) {
    (void)flags;
    if (!K_out || n < 0 || m < 0 || p < 0) return LGP_ERR_BADARG;
    if (n == 0 || m == 0) return LGP_OK;
    if (ldk < m) return LGP_ERR_BADARG;
    if (width < 1 || width > 3 || nrows < 1 || nrows > B_MAX_ROWS || !rows) return LGP_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    BartDesc d;
    memset(&d, 0, sizeof(d));
    // compact away zero-weight covariates (reference masks them: _bart.py:669-672)
    int pa = 0;
    double Wn = 0.0;
    for (int k = 0; k < p; k++) {
        double wk = w ? w[k] : 1.0;
        if (wk == 0.0) continue;
        if (pa >= B_MAX_P) return LGP_ERR_UNSUPPORTED;
        d.dim[pa] = k;
        d.n[pa] = nsplits[k];
        d.w[pa] = wk;
        if (nsplits[k]) Wn += wk;  // Wn = sum(where(n, w, 0))   (_bart.py:694)
        pa++;
    }
    if (pa == 0) {
        // no covariates: correlation is identically 1 (_bart.py:659-661)
        dim3 g((unsigned)((m + 255) / 256), (unsigned)n);
        fill_kernel<<<g, 256, 0, st>>>(K_out, ldk, n, m, amp);
        LGP_CUDA_CHECK_LAUNCH();
        return LGP_OK;
    }
    if (width == 3 && !psi) return LGP_ERR_BADARG;
    d.p = pa;
    d.width = width;
    d.nrows = nrows;
    d.Wn = Wn;
    d.inv_Wn = 1.0 / Wn;
    d.gamma = gamma;
    d.amp = amp;
    for (int k = 0; k < pa; k++) {
        int nk = d.n[k];
        d.wn[k] = nk ? d.w[k] / (double)nk : 0.0;
        d.w_inv_Wn[k] = d.w[k] * d.inv_Wn;
        d.inv_Wnmod[k] = 1.0 / (Wn - (nk ? d.w[k] : 0.0));
        // digamma(n or 1) from the same extended-precision recurrence as the table
        long double v = -0.577215664901532860606512090082402431L;
        int nn = nk ? nk : 1;
        for (int q = 1; q < nn; q++) v += 1.0L / (long double)q;
        d.psin[k] = (double)v;
    }
    for (int r = 0; r < nrows; r++)
        for (int c = 0; c < width; c++) d.rows[r][c] = rows[r * width + c];
    size_t smem = (size_t)2 * pa * BT * sizeof(int32_t);
    dim3 grid((unsigned)((m + BT - 1) / BT), (unsigned)((n + BT - 1) / BT));
    if (grid.y > 65535) return LGP_ERR_UNSUPPORTED;
    int vec_ok = ((ldk & 1) == 0) && ((reinterpret_cast<uintptr_t>(K_out) & 15) == 0);
    gram_bart_kernel<<<grid, B_THREADS, smem, st>>>(d, psi, ix, ldx, n, iy, ldy, m, K_out, ldk, vec_ok);
    LGP_CUDA_CHECK_LAUNCH();
    return LGP_OK;
}

}  // extern "C"
