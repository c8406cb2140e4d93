// Matern kernel of general real order nu on the device: 2/Gamma(nu) (x/2)^nu K_nu(x), x = sqrt(2 nu r2), and its derivative
// with respect to the squared argument.
//
// Reference: src/lsqfitgp/_kernels/_matern.py:55-76 (Matern), src/lsqfitgp/_special/_bessel.py:70-99 (kvmodx2 and its JVP
// -kvmodx2(nu - 1, x2, 1)/4).  The reference ships every Gram entry to the host and calls scipy.special.kv (AMOS zbesk)
// through jax.pure_callback; here K_nu is evaluated in the kernel:
//   * x <= 2: Temme's series (N. M. Temme, J. Comput. Phys. 19 (1975) 324) for K_mu, K_{mu+1}, |mu| <= 1/2;
//   * x >  2: Steed's algorithm for the continued fraction CF2 of the Whittaker function (Thompson & Barnett 1986),
//             evaluated scaled by e^x so that the tails do not underflow prematurely;
//   * stable upward recurrence K_{a+1} = K_{a-1} + (2a/x) K_a to the requested order.
// The order is a kernel hyperparameter, identical for all entries: everything that depends on nu only (fractional
// order mu, Temme's Gamma-function combinations, 2/Gamma(nu)) is computed once on the host (matern_nu_setup) from the
// Taylor series of 1/Gamma(1+t) (rgamma_table.inc, generated with mpmath) and passed in the descriptor.
//
// Plain C++ so that tests/test_bessel_cpu.py can check the same source on the host against mpmath.
#pragma once
#include <math.h>

#include "fastmath.cuh"
#include "rgamma_table.inc"

namespace lgp {

constexpr int MATERN_NPAR = 10;
// layout of the parameter block
enum { MN_NU = 0, MN_MU, MN_NSTEPS, MN_FACT, MN_GAM1, MN_GAM2, MN_GAMPL, MN_GAMMI, MN_C0, MN_TWONU };
constexpr double MATERN_NU_MAX = 100.0;

// Host: constants of order nu.  Returns false if the order is not supported (nu < 0, nu > MATERN_NU_MAX, NaN).
static inline bool matern_nu_setup(double nu, double *par) {
    if (!(nu >= 0.0 && nu <= MATERN_NU_MAX)) return false;
    static const double a[LGP_RGAMMA_NCOEF] = {LGP_RGAMMA_COEFS};
    double mu, nsteps;
    if (nu < 0.5) {
        // start at mu = -nu: (K_mu, K_{mu+1}) = (K_nu, K_{1-nu}) = (K_nu, K_{nu-1})
        mu = -nu;
        nsteps = -1.0;
    } else {
        const double nl = floor(nu + 0.5);
        mu = nu - nl;         // in [-1/2, 1/2)
        nsteps = nl - 1.0;    // recurrence steps from (K_mu, K_{mu+1}) to (K_{nu-1}, K_nu)
    }
    // 1/Gamma(1 +- mu) and Temme's combinations from the even / odd parts of the series: no cancellation as mu -> 0
    double even = 0.0, odd = 0.0;  // sum_{k even} a_k mu^k, sum_{k odd} a_k mu^(k-1)
    const double m2 = mu * mu;
    for (int k = LGP_RGAMMA_NCOEF - 2; k >= 0; k -= 2) {
        even = even * m2 + a[k];
        odd = odd * m2 + a[k + 1];
    }
    const double pimu = 3.141592653589793238462643383279502884 * mu;
    par[MN_NU] = nu;
    par[MN_MU] = mu;
    par[MN_NSTEPS] = nsteps;
    par[MN_FACT] = fabs(pimu) < 1e-8 ? 1.0 : pimu / sin(pimu);
    par[MN_GAM1] = -odd;                 // (1/Gamma(1-mu) - 1/Gamma(1+mu)) / (2 mu)
    par[MN_GAM2] = even;                 // (1/Gamma(1-mu) + 1/Gamma(1+mu)) / 2
    par[MN_GAMPL] = even + mu * odd;     // 1/Gamma(1+mu)
    par[MN_GAMMI] = even - mu * odd;     // 1/Gamma(1-mu)
    par[MN_C0] = nu == 0.0 ? 0.0 : 2.0 / tgamma(nu);
    par[MN_TWONU] = 2.0 * (nu != 0.0 ? nu : 1.0);  // _matern.py:74: for nu = 0 the limit is white noise, avoid r2 * 0
    return true;
}

// K_mu(x) and K_{mu+1}(x), x > 0, |mu| <= 1/2.  For x > 2 the two values are scaled by e^x and *ex = e^-x; for x <= 2
// they are unscaled and *ex = 1.
LGP_FM_HD void bessel_k_start(const double *par, double x, double &k0, double &k1, double &ex) {
    const double mu = par[MN_MU];
    const double EPS = 1.0e-17;
    if (x <= 2.0) {
        const double b = 0.5 * x;
        const double d = -log(b);
        const double e = mu * d;
        const double e2 = e * e;
        // sinh(e)/e without cancellation
        const double shc = fabs(e) < 0.05 ? 1.0 + e2 * (1.0 / 6.0 + e2 * (1.0 / 120.0 + e2 * (1.0 / 5040.0 + e2 / 362880.0)))
                                          : sinh(e) / e;
        double ff = par[MN_FACT] * (par[MN_GAM1] * cosh(e) + par[MN_GAM2] * shc * d);
        const double em = exp(e);  // b^-mu
        double p = 0.5 * em / par[MN_GAMPL];
        double q = 0.5 / (em * par[MN_GAMMI]);
        double c = 1.0;
        const double b2 = b * b, mu2 = mu * mu;
        double sum = ff, sum1 = p;
        for (int i = 1; i <= 200; i++) {
            const double di = (double)i;
            // (reciprocals by fm_rcp_fast: ~1 ulp each, far inside the 1e-13 budget; the IEEE divisions were 3/4 of the loop)
            ff = (di * ff + p + q) * fm_rcp_fast(di * di - mu2);
            c *= b2 * fm_rcp_fast(di);
            p *= fm_rcp_fast(di - mu);
            q *= fm_rcp_fast(di + mu);
            const double del = c * ff;
            sum += del;
            sum1 += c * (p - di * ff);
            if (fabs(del) < fabs(sum) * EPS) break;
        }
        k0 = sum;
        k1 = sum1 * (2.0 / x);
        ex = 1.0;
        return;
    }
    // Steed's algorithm for CF2
    double b = 2.0 * (1.0 + x);
    double d = 1.0 / b;
    double h = d, delh = d;
    double q1 = 0.0, q2 = 1.0;
    const double a1 = 0.25 - mu * mu;
    double q = a1, c = a1, a = -a1;
    double s = 1.0 + q * delh;
    for (int i = 2; i <= 2000; i++) {
        a -= (double)(2 * (i - 1));
        c = -a * c * fm_rcp_fast((double)i);
        const double qnew = (q1 - b * q2) * fm_rcp_fast(a);
        q1 = q2;
        q2 = qnew;
        q += c * qnew;
        b += 2.0;
        d = fm_rcp_fast(b + a * d);
        delh = (b * d - 1.0) * delh;
        h += delh;
        const double dels = q * delh;
        s += dels;
        if (fabs(dels) < fabs(s) * EPS) break;
    }
    h = a1 * h;
    k0 = sqrt(3.141592653589793238462643383279502884 / (2.0 * x)) / s;
    k1 = k0 * (mu + x + 0.5 - h) / x;
    ex = exp(-x);
}

// Matern(nu) core at squared distance r2 (before the 2 nu factor), and d core / d r2.
// `want_deriv` is a compile-time-foldable flag (the derivative costs one more pow-free product).
LGP_FM_HD void matern_nu_core(const double *par, double r2, bool want_deriv, double &val, double &dr2) {
    const double z = par[MN_TWONU] * r2;
    dr2 = 0.0;
    if (!(z > 0.0)) {  // z == 0: the limit (kvmodx2: where(x2, normal, atzero), atzero = 1); NaN propagates
        val = z == 0.0 ? 1.0 : z;
        return;
    }
    const double x = sqrt(z);
    const double nu = par[MN_NU];
    double k0, k1, ex;
    bessel_k_start(par, x, k0, k1, ex);
    const double inv = 2.0 / x;
    const int nsteps = (int)par[MN_NSTEPS];
    double km1, k;  // K_{nu-1}, K_nu
    if (nsteps < 0) {
        k = k0;
        km1 = k1;
    } else {
        double aj = par[MN_MU];
        for (int j = 0; j < nsteps; j++) {
            aj += 1.0;
            const double kn = k0 + aj * inv * k1;
            k0 = k1;
            k1 = kn;
        }
        km1 = k0;
        k = k1;
    }
    const double hx = 0.5 * x;
    const double pw = pow(hx, nu);
    if (!(pw > 1e-280) && hx < 1.0) {
        // (x/2)^nu underflows while K_nu overflows: x is so small that the kernel equals 1 in double precision
        val = 1.0;
        if (nu > 1.0) dr2 = -0.25 * par[MN_TWONU] / (nu - 1.0);  // limit of -kvmodx2(nu - 1, z, 1)/4 (_bessel.py:73-82)
        return;
    }
    const double cp = par[MN_C0] * pw;
    val = cp * (ex * k);
    if (want_deriv) dr2 = -0.25 * par[MN_TWONU] * ((cp / hx) * (ex * km1));
}

}  // namespace lgp
