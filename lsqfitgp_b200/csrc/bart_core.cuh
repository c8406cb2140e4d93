// Per-pair arithmetic of the BART prior correlation (fast path, brackets of width <= 3), shared by the CUDA kernels of
// gram_bart.cu and by a host build used by the CPU accuracy tests (the test-side harness compiles this very header
// with g++).  Plain C++.
//
// Reference: BART._correlation, src/lsqfitgp/_kernels/_bart.py:628-757, with the bracket folding of BART.correlation
// (:415-455) done by the caller.  Operation order follows the reference (every operation individually rounded).
//
// What is restructured for the GPU (results unchanged):
//   * bin indices travel as doubles (integers < 2^53 are exact), so min / max / differences need no int<->double
//     conversions per pair;
//   * everything that depends on ONE index only -- digamma(1 + i), digamma(1 + n - i), 1/i, 1/(n - i) -- is computed
//     once per staged point (BartPoint) and SELECTED per pair by comparing the two indices, instead of gathered from the
//     digamma table / divided per pair;
//   * the two divisions per (pair, dimension) become a multiplication by the staged reciprocal with two exact-residual
//     corrections (fm_div_recip of fastmath.cuh: correctly rounded for the operand ranges that occur here).
//   * alpha / beta derivatives ride along the `repeat` scan as dual numbers (the scan is affine in gamma and
//     multilinear in the probabilities of each row): forward mode of `_bart.py:749-757`.
#pragma once
#include "fastmath.cuh"

namespace lgp {

#if defined(__CUDA_ARCH__)
#define LGP_B_MUL(a, b) __dmul_rn((a), (b))
#define LGP_B_ADD(a, b) __dadd_rn((a), (b))
#define LGP_B_SUB(a, b) __dsub_rn((a), (b))
#else
// host build: compiled with -ffp-contract=off
#define LGP_B_MUL(a, b) ((a) * (b))
#define LGP_B_ADD(a, b) ((a) + (b))
#define LGP_B_SUB(a, b) ((a) - (b))
#endif

// per-dimension constants (one covariate with non-zero weight)
struct BartDim {
    double n;      // number of splits
    double wn;     // n ? w/n : 0
    double w;      // weight
    double wiW;    // w * inv_Wn
    double wmod;   // w * inv_Wnmod,  inv_Wnmod = 1/(Wn - (n ? w : 0))
    double iWmod;  // inv_Wnmod
    double psin2;  // 2 * digamma(n or 1)
};

// per-point, per-dimension staged values
struct BartPoint {
    double v;   // bin index in [0, n]
    double pa;  // digamma(1 + v)
    double pb;  // digamma(1 + n - v)
    double rm;  // 1/v        (unused when v == 0)
    double rp;  // 1/(n - v)  (unused when v == n)
};

LGP_FM_HD BartDim bart_dim(double n, double w, double Wn, double inv_Wn, const double *psi /* psi[k] = digamma(k) */) {
    BartDim d;
    d.n = n;
    d.w = w;
    d.wn = n != 0.0 ? w / n : 0.0;
    d.wiW = LGP_B_MUL(w, inv_Wn);
    d.iWmod = 1.0 / LGP_B_SUB(Wn, n != 0.0 ? w : 0.0);
    d.wmod = LGP_B_MUL(w, d.iWmod);
    d.psin2 = psi ? LGP_B_MUL(2.0, psi[n != 0.0 ? (long)n : 1]) : 0.0;  // only the width-3 terms use it
    return d;
}

LGP_FM_HD BartPoint bart_point(int idx, int n, const double *psi) {
    BartPoint p;
    p.v = (double)idx;
    p.pa = psi[1 + idx];
    p.pb = psi[1 + n - idx];
    p.rm = 1.0 / (double)idx;
    p.rp = 1.0 / (double)(n - idx);
    return p;
}

// first pass over the dimensions: S2 += wn*|ix-iy| (width-2 brackets, :698-699), S3 += wn*(n - |ix-iy|) (width 3, :727)
template <bool NEED2, bool NEED3>
LGP_FM_HD void bart_pass1(const BartDim &d, double vx, double vy, double &S2, double &S3, bool &any0) {
    const double n0 = fabs(vx - vy);
    any0 = any0 || (n0 != 0.0);
    if (NEED2) S2 = LGP_B_ADD(S2, LGP_B_MUL(d.wn, n0));
    if (NEED3) S3 = LGP_B_ADD(S3, LGP_B_MUL(d.wn, LGP_B_SUB(d.n, n0)));
}

// second pass (width 3 only): sumi += wn * (terms1 - terms2 - terms3)   (:717-747), S = the complete S3
LGP_FM_HD void bart_pass2(const BartDim &d, double inv_Wn, double S, const BartPoint &x, const BartPoint &y, double &sumi) {
    const bool xge = x.v >= y.v, xle = x.v <= y.v;
    const double hi = xge ? x.v : y.v, lo = xge ? y.v : x.v;
    const double n0 = hi - lo;
    const double nplus0 = d.n - lo, nout = d.n - n0;  // nminus0 = hi
    const double inv_Wnminus = nplus0 != 0.0 ? inv_Wn : d.iWmod;
    const double inv_Wnplus = hi != 0.0 ? inv_Wn : d.iWmod;
    const double t = LGP_B_MUL(d.wn, n0);
    const double terms1 = LGP_B_MUL(
        LGP_B_ADD(S, t), LGP_B_ADD(LGP_B_ADD(inv_Wnminus, inv_Wnplus), LGP_B_MUL(inv_Wn, nout - 2.0)));
    const double wiWn0 = LGP_B_MUL(d.wiW, n0);
    // reciprocal of (n - lo): lo belongs to x when x <= y; reciprocal of hi: hi belongs to x when x >= y
    const double rplus = xle ? x.rp : y.rp, rminus = xge ? x.rm : y.rm;
    const double t2a = nplus0 != 0.0 ? fm_div_recip(wiWn0, nplus0, rplus) : d.wmod;
    const double t2b = hi != 0.0 ? fm_div_recip(wiWn0, hi, rminus) : d.wmod;
    const double terms2 = LGP_B_ADD(t2a, t2b);
    const double psiminus = xge ? x.pa : y.pa;  // digamma(1 + hi)
    const double psiplus = xle ? x.pb : y.pb;   // digamma(1 + n - lo)
    const double terms3 = LGP_B_MUL(wiWn0, LGP_B_SUB(LGP_B_SUB(d.psin2, psiminus), psiplus));
    const double terms = LGP_B_SUB(LGP_B_SUB(terms1, terms2), terms3);
    sumi = LGP_B_ADD(sumi, LGP_B_MUL(d.wn, terms));
}

// One row of the `repeat` scan, value and (DUAL) derivatives w.r.t. alpha and beta.
//   row[c], c < width: non-termination probabilities; da[c], db[c]: their derivatives w.r.t. alpha, beta.
//   g, ga, gb: gamma carried from the previous row (deeper bracket) and its derivatives; updated in place.
template <bool DUAL>
LGP_FM_HD void bart_row(int width, bool any0, double Wn, double inv_Wn, double S2, double S3, double sumi,
                        const double *row, const double *da, const double *db, double &g, double &ga, double &gb) {
    double res, ra = 0.0, rb = 0.0;
    if (width == 1) {
        // 1 - (1 - gamma) * pnt[0]                                                          (:688)
        const double omg = LGP_B_SUB(1.0, g);
        res = LGP_B_SUB(1.0, LGP_B_MUL(omg, row[0]));
        if (DUAL) {
            ra = ga * row[0] - omg * da[0];
            rb = gb * row[0] - omg * db[0];
        }
    } else if (width == 2) {
        // Q = 1 - pnt[1] + gamma * pnt[1]; result = 1 - P0 + Q * (P0 - P0 / Wn * sum_term)   (:702-704)
        const double P0 = row[0], P1 = row[1];
        const double Q = LGP_B_ADD(LGP_B_SUB(1.0, P1), LGP_B_MUL(g, P1));
        const double inner = LGP_B_SUB(P0, LGP_B_MUL(P0 / Wn, S2));
        res = LGP_B_ADD(LGP_B_SUB(1.0, P0), LGP_B_MUL(Q, inner));
        if (DUAL) {
            const double f = 1.0 - S2 / Wn;  // inner = P0 * f
            const double Qa = (g - 1.0) * da[1] + P1 * ga, Qb = (g - 1.0) * db[1] + P1 * gb;
            ra = -da[0] + Qa * inner + Q * f * da[0];
            rb = -db[0] + Qb * inner + Q * f * db[0];
        }
    } else {
        // Q = 1 + pnt[2] * (gamma - 1); sump = S + pnt[1] * (Q * sumi - S); result = 1 + pnt[0] * (inv_Wn * sump - 1)   (:751-753)
        const double gm1 = LGP_B_SUB(g, 1.0);
        const double Q = LGP_B_ADD(1.0, LGP_B_MUL(row[2], gm1));
        const double inner = LGP_B_SUB(LGP_B_MUL(Q, sumi), S3);
        const double sump = LGP_B_ADD(S3, LGP_B_MUL(row[1], inner));
        const double outer = LGP_B_SUB(LGP_B_MUL(inv_Wn, sump), 1.0);
        res = LGP_B_ADD(1.0, LGP_B_MUL(row[0], outer));
        if (DUAL) {
            const double Qa = da[2] * gm1 + row[2] * ga, Qb = db[2] * gm1 + row[2] * gb;
            const double sa = da[1] * inner + row[1] * Qa * sumi, sb = db[1] * inner + row[1] * Qb * sumi;
            ra = da[0] * outer + row[0] * inv_Wn * sa;
            rb = db[0] * outer + row[0] * inv_Wn * sb;
        }
    }
    g = any0 ? res : 1.0;
    if (DUAL) {
        ga = any0 ? ra : 0.0;
        gb = any0 ? rb : 0.0;
    }
}

}  // namespace lgp
