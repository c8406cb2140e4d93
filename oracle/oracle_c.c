/* C restatement (checker side, never linked into the product) of two integer/byte-level pieces of the path:
 *   fasthash64            src/lsqfitgp/_jaxext/_fasthash.py:56-97  (== tests/fast-hash/fasthash.c:34-66 of the reference)
 *   bart_pair_w3          src/lsqfitgp/_kernels/_bart.py:669-757   (depth-3 closed form + `repeat` scan), one pair
 * Used by tests/ to cross-check the NumPy oracle at sizes where pure-Python loops are too slow.
 * Build: make -C oracle  (gcc -O2 -ffp-contract=off, so every operation is individually rounded). */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

static uint64_t mix(uint64_t h) {
    h ^= h >> 23;
    h *= 0x2127599bf4325c37ULL;
    h ^= h >> 47;
    return h;
}

uint64_t oracle_fasthash64(const void *buf, size_t len, uint64_t seed) {
    const uint64_t m = 0x880355f21e6d1965ULL;
    const unsigned char *p = (const unsigned char *)buf;
    uint64_t h = seed ^ (len * m);
    size_t nw = len / 8;
    for (size_t i = 0; i < nw; i++) {
        uint64_t v;
        memcpy(&v, p + 8 * i, 8); /* little-endian host */
        h ^= mix(v);
        h *= m;
    }
    size_t tail = len & 7;
    if (tail) {
        uint64_t v = 0;
        for (size_t i = 0; i < tail; i++) v |= (uint64_t)p[8 * nw + i] << (8 * i);
        h ^= mix(v);
        h *= m;
    }
    return mix(h);
}

/* digamma of a positive integer: psi(1) = -gamma, psi(k+1) = psi(k) + 1/k, in long double */
static double psi_int(long k) {
    long double v = -0.577215664901532860606512090082402431L;
    for (long q = 1; q < k; q++) v += 1.0L / (long double)q;
    return (double)v;
}

/* rows: nrows x 3, deepest bracket first; n, ix, iy: p entries; w: p weights. */
double oracle_bart_pair_w3(int p, const int64_t *n_in, const int64_t *ix_in, const int64_t *iy_in, const double *w,
                           const double *rows, int nrows, double gamma) {
    if (p == 0) return 1.0;
    int anyn0 = 0;
    double Wn = 0.0;
    for (int i = 0; i < p; i++) {
        int64_t n = w[i] != 0 ? n_in[i] : 0, x = w[i] != 0 ? ix_in[i] : 0, y = w[i] != 0 ? iy_in[i] : 0;
        if (x != y) anyn0 = 1;
        if (n) Wn += w[i];
    }
    double inv_Wn = 1.0 / Wn;
    double S = 0.0;
    for (int i = 0; i < p; i++) {
        int64_t n = w[i] != 0 ? n_in[i] : 0, x = w[i] != 0 ? ix_in[i] : 0, y = w[i] != 0 ? iy_in[i] : 0;
        int64_t lo = x < y ? x : y, hi = x < y ? y : x, n0 = hi - lo, nout = n - n0;
        double wn = n ? w[i] / (double)n : 0.0;
        S = S + wn * (double)nout;
    }
    double sumi = 0.0;
    for (int i = 0; i < p; i++) {
        int64_t n = w[i] != 0 ? n_in[i] : 0, x = w[i] != 0 ? ix_in[i] : 0, y = w[i] != 0 ? iy_in[i] : 0;
        int64_t lo = x < y ? x : y, hi = x < y ? y : x, n0 = hi - lo;
        int64_t nminus0 = hi, nplus0 = n - lo, nout = n - n0;
        double wn = n ? w[i] / (double)n : 0.0;
        double inv_Wnmod = 1.0 / (Wn - (n ? w[i] : 0.0));
        double inv_Wnminus = nplus0 ? inv_Wn : inv_Wnmod;
        double inv_Wnplus = nminus0 ? inv_Wn : inv_Wnmod;
        double t = wn * (double)n0;
        double terms1 = (S + t) * (inv_Wnminus + inv_Wnplus + inv_Wn * (double)(nout - 2));
        double terms2 = (nplus0 ? w[i] * inv_Wn * (double)n0 / (double)nplus0 : w[i] * inv_Wnmod) +
                        (nminus0 ? w[i] * inv_Wn * (double)n0 / (double)nminus0 : w[i] * inv_Wnmod);
        double psin = psi_int(n ? n : 1), psiminus = psi_int(1 + hi), psiplus = psi_int(1 + n - lo);
        double terms3 = w[i] * inv_Wn * (double)n0 * (2 * psin - psiminus - psiplus);
        sumi = sumi + wn * (terms1 - terms2 - terms3);
    }
    double g = gamma;
    for (int r = 0; r < nrows; r++) {
        double Q = 1 + rows[3 * r + 2] * (g - 1);
        double sump = S + rows[3 * r + 1] * (Q * sumi - S);
        double result = 1 + rows[3 * r + 0] * (inv_Wn * sump - 1);
        g = anyn0 ? result : 1.0;
    }
    return g;
}
