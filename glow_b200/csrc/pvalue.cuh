// Two-sided Student-t p-value, p = 2 * t.sf(|t|, dof) = I_x(dof/2, 1/2), x = dof / (dof + t^2).
// Replaces scipy's stats.distributions.t.sf at python/glow/gwas/lin_reg.py:245.
//
// Regularised incomplete beta with a = dof/2 (up to ~2.5e5 at biobank scale) and b = 1/2:
//   * bulk: power series of I_y(1/2, a), y = 1 - x, and p = 1 - I_y.  Used while the series
//     converges quickly AND p is not tiny: t^2 <= 20 (p >= 7.7e-6 for every dof, so the subtraction
//     costs at most 5 of the 16 digits -- the north-star bar on p is 1e-6) and y small enough that
//     the term ratio drops below 1/2 within 64 terms.  Nearly every null test takes this branch, so
//     warps do not diverge into the (much more expensive) continued fraction;
//   * tail: modified-Lentz continued fraction of I_x(a, 1/2) evaluated directly, with the
//     prefactor x^a (1-x)^(1/2) / (a B(a,1/2)) assembled in log space so that p stays
//     relatively accurate down to the double underflow threshold (~1e-308).
// ln B(a,1/2) is a per-run constant computed on the host in extended precision (PvalConst).
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace glr {

struct PvalConst {
  double nu;        // degrees of freedom
  double a;         // nu / 2
  double logc_cf;   // -ln B(a,1/2) - ln a
  double logc_ser;  // -ln B(a,1/2) + ln 2
  double y_switch;  // (b+1)/(a+b+2): below this the series is always the right choice
  double y_series;  // series also used for y <= y_series when t^2 <= 20
  double ser_coef[64];  // (a + 1/2 + n) / (3/2 + n): term ratio of the series without the division
};

#ifndef __CUDACC_RTC__
inline PvalConst make_pval_const(double nu) {
  PvalConst c;
  c.nu = nu;
  c.a = 0.5 * nu;
  long double a = 0.5L * (long double)nu;
  long double d;  // ln Gamma(a + 1/2) - ln Gamma(a)
  if (nu <= 0) {
    d = 0;
  } else if (a >= 100.0L) {
    long double ia = 1.0L / a, ia2 = ia * ia;
    d = 0.5L * logl(a) + ia * (-0.125L + ia2 * (1.0L / 192 + ia2 * (-1.0L / 640 + ia2 * (17.0L / 14336))));
  } else {
    d = lgammal(a + 0.5L) - lgammal(a);
  }
  long double lnB = 0.5L * logl(3.14159265358979323846264338327950288L) - d;  // ln Gamma(1/2) - d
  c.logc_cf = (double)(-lnB - logl(a > 0 ? a : 1.0L));
  c.logc_ser = (double)(-lnB + logl(2.0L));
  c.y_switch = 1.5 / (c.a + 2.5);
  c.y_series = 0.5 * 65.5 / (c.a + 64.5);
  if (c.y_series < c.y_switch) c.y_series = c.y_switch;
  for (int n = 0; n < 64; ++n) c.ser_coef[n] = (c.a + 0.5 + n) / (1.5 + n);
  return c;
}
#endif

__device__ __forceinline__ double t_two_sided_pvalue(double t, const PvalConst& c) {
  if (!(c.nu > 0.0) || isnan(t)) return nan("");
  const double t2 = t * t;
  if (isinf(t2)) return 0.0;
  if (t2 == 0.0) return 1.0;
  const double a = c.a;
  const double denom = c.nu + t2;
  const double y = t2 / denom;
  const double lx = -log1p(t2 / c.nu);
  if (y <= c.y_switch || (t2 <= 20.0 && y <= c.y_series)) {
    // I_y(1/2, a) = y^(1/2) (1-y)^a / (1/2 B) * sum_n (a+1/2)_n / (3/2)_n y^n
    double s = 1.0, term = 1.0;
    int n = 0;
    for (; n < 64; ++n) {
      term *= c.ser_coef[n] * y;
      s += term;
      if (term < 1e-16 * s) break;
    }
    if (n == 64) {
      for (; n < 2000; ++n) {
        term *= (a + 0.5 + n) / (1.5 + n) * y;
        s += term;
        if (term < 1e-16 * s) break;
      }
    }
    const double q = exp(0.5 * log(y) + a * lx + c.logc_ser) * s;
    return 1.0 - q;
  }
  // continued fraction for I_x(a, b), b = 1/2 (modified Lentz)
  const double x = c.nu / denom;
  const double b = 0.5, tiny = 1e-300;
  const double qab = a + b, qap = a + 1.0, qam = a - 1.0;
  double cc = 1.0, d = 1.0 - qab * x / qap;
  if (fabs(d) < tiny) d = tiny;
  d = 1.0 / d;
  double h = d;
  for (int m = 1; m < 5000; ++m) {
    const double m2 = 2.0 * m;
    double aa = m * (b - m) * x / ((qam + m2) * (a + m2));
    d = 1.0 + aa * d;
    if (fabs(d) < tiny) d = tiny;
    cc = 1.0 + aa / cc;
    if (fabs(cc) < tiny) cc = tiny;
    d = 1.0 / d;
    h *= d * cc;
    aa = -(a + m) * (qab + m) * x / ((a + m2) * (qap + m2));
    d = 1.0 + aa * d;
    if (fabs(d) < tiny) d = tiny;
    cc = 1.0 + aa / cc;
    if (fabs(cc) < tiny) cc = tiny;
    d = 1.0 / d;
    const double del = d * cc;
    h *= del;
    if (fabs(del - 1.0) < 4e-15) break;  // a few ulps: |del - 1| can hover at 1-2 ulp forever
  }
  return exp(a * lx + 0.5 * log(y) + c.logc_cf) * h;
}

}  // namespace glr
