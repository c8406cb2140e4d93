/* TEST INFRASTRUCTURE: host build of lsqfitgp_b200/csrc/fastmath.cuh (the short exp/sqrt of the Gram kernels) so that
 * tests/test_fastmath_cpu.py can measure their error against glibc without a GPU.  Compiled as C++ by oracle/Makefile
 * (g++ -x c++).  The hardware reciprocal-square-root estimate (MUFU.RSQ64H: upper 32 input bits, ~2^-21 relative) is
 * emulated by truncating the argument and the result to 20 mantissa bits. */
#include "../lsqfitgp_b200/csrc/fastmath.cuh"

static double trunc20(double v) {
    uint64_t u = lgp::fm_to_bits(v) & ~((1ull << 32) - 1);
    return lgp::fm_from_bits(u);
}

extern "C" {
void lgp_host_exp_neg(const double *a, double *out, long n) {
    for (long i = 0; i < n; i++) out[i] = lgp::fm_exp_neg(a[i], lgp::EXP_TAB_HOST);
}
void lgp_host_exp_neg_fast(const double *a, double *out, long n) {
    for (long i = 0; i < n; i++) out[i] = lgp::fm_exp_neg_fast(a[i], lgp::EXP_TAB_HOST);
}
void lgp_host_div_recip(const double *a, const double *b, double *out, long n) {
    for (long i = 0; i < n; i++) {
        const double y = 1.0 / b[i];
        out[i] = lgp::fm_div_recip_ok(a[i]) ? lgp::fm_div_recip(a[i], b[i], y) : a[i] / b[i];
    }
}
void lgp_host_log_ge1(const double *x, double *out, long n) {
    for (long i = 0; i < n; i++) out[i] = lgp::fm_log_ge1_fast(x[i], lgp::LOG_TAB_HOST);
}
/* rational quadratic core as the Gram fast path evaluates it: (1 + r2/beta)^(-beta/2) */
void lgp_host_ratquad(double beta, const double *r2, double *out, long n) {
    const double rb = 1.0 / beta, cexp = -0.5 * beta;
    for (long i = 0; i < n; i++) {
        const double t = lgp::fm_div_recip_ok(r2[i]) ? lgp::fm_div_recip(r2[i], beta, rb) : r2[i] / beta;
        const double y = cexp * lgp::fm_log_ge1_fast(1.0 + t, lgp::LOG_TAB_HOST);
        out[i] = lgp::fm_exp_neg_fast(y, lgp::EXP_TAB_HOST);
    }
}
void lgp_host_sqrt(const double *z, double *out, long n) {
    for (long i = 0; i < n; i++) {
        double y0 = trunc20(1.0 / sqrt(trunc20(z[i])));
        out[i] = lgp::fm_sqrt_from_rsqrt(z[i], y0);
    }
}
}

/* host build of the general-order Matern core (lsqfitgp_b200/csrc/bessel_k.cuh) for tests/test_bessel_cpu.py */
#include "../lsqfitgp_b200/csrc/bessel_k.cuh"

extern "C" {
int lgp_host_matern_nu(double nu, const double *r2, double *val, double *dr2, long n) {
    double par[lgp::MATERN_NPAR];
    if (!lgp::matern_nu_setup(nu, par)) return 1;
    for (long i = 0; i < n; i++) lgp::matern_nu_core(par, r2[i], true, val[i], dr2[i]);
    return 0;
}
int lgp_host_matern_nu_par(double nu, double *par) { return lgp::matern_nu_setup(nu, par) ? 0 : 1; }
}
