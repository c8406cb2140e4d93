/* TEST INFRASTRUCTURE: host build of lsqfitgp_b200/csrc/bart_core.cuh (the per-pair arithmetic of the BART Gram kernels:
 * value, alpha / beta duals, staged-reciprocal divisions) so that tests/test_bart_core_cpu.py can compare it with the
 * oracle restatement of the reference (oracle/bart.py) without a GPU.  Compiled as C++ by oracle/Makefile (g++ -x c++,
 * -ffp-contract=off).  Mirrors the flow of gram_bart_kernel: zero-weight compaction, pass 1 (S2, S3, equality), pass 2
 * (width-3 terms), the chained `repeat` scans. */
#include <vector>

#include "../lsqfitgp_b200/csrc/bart_core.cuh"

extern "C" {

/* ix, iy: npairs x p (row-major) bin indices; rows: total x 3; drows: 2 x total x 3 (may be NULL); out: npairs x 3
 * (corr, d corr/d alpha, d corr/d beta) */
int lgp_host_bart_pairs(int p, const int *nsplits, const double *w, int nstages, const int *stage_width,
                        const int *stage_nrows, const double *rows, const double *drows, double gamma, const int *ix,
                        const int *iy, long npairs, double *out) {
    int total = 0, need2 = 0, need3 = 0, nmax = 0;
    for (int s = 0; s < nstages; s++) {
        total += stage_nrows[s];
        need2 |= stage_width[s] == 2;
        need3 |= stage_width[s] == 3;
    }
    std::vector<int> dim;
    double Wn = 0.0;
    for (int k = 0; k < p; k++) {
        if (w[k] == 0.0) continue;
        dim.push_back(k);
        if (nsplits[k]) Wn += w[k];
        if (nsplits[k] > nmax) nmax = nsplits[k];
    }
    const double inv_Wn = 1.0 / Wn;
    std::vector<double> psi(nmax + 3);
    psi[0] = -INFINITY;
    long double v = -0.577215664901532860606512090082402431L;
    for (int k = 1; k < nmax + 3; k++) {
        psi[k] = (double)v;
        v += 1.0L / (long double)k;
    }
    std::vector<double> zero(3 * total, 0.0);
    const double *da = drows ? drows : zero.data(), *db = drows ? drows + 3 * total : zero.data();
    for (long q = 0; q < npairs; q++) {
        double S2 = 0.0, S3 = 0.0, sumi = 0.0;
        bool any0 = false;
        for (int k : dim) {
            const lgp::BartDim d = lgp::bart_dim((double)nsplits[k], w[k], Wn, inv_Wn, psi.data());
            lgp::bart_pass1<true, true>(d, (double)ix[q * p + k], (double)iy[q * p + k], S2, S3, any0);
        }
        if (need3)
            for (int k : dim) {
                const lgp::BartDim d = lgp::bart_dim((double)nsplits[k], w[k], Wn, inv_Wn, psi.data());
                const lgp::BartPoint px = lgp::bart_point(ix[q * p + k], nsplits[k], psi.data());
                const lgp::BartPoint py = lgp::bart_point(iy[q * p + k], nsplits[k], psi.data());
                lgp::bart_pass2(d, inv_Wn, S3, px, py, sumi);
            }
        double g = gamma, ga = 0.0, gb = 0.0;
        if (dim.empty()) {
            g = 1.0;
        } else {
            int r = 0;
            for (int s = 0; s < nstages; s++)
                for (int t = 0; t < stage_nrows[s]; t++, r++)
                    lgp::bart_row<true>(stage_width[s], any0, Wn, inv_Wn, S2, S3, sumi, rows + 3 * r, da + 3 * r, db + 3 * r,
                                        g, ga, gb);
        }
        out[3 * q] = g;
        out[3 * q + 1] = ga;
        out[3 * q + 2] = gb;
    }
    (void)need2;
    return 0;
}
}
