/* incgamma.h — regularised incomplete gamma function and its inverse (C / CUDA dual header).
 *
 * Stands in for R::pgamma / R::qgamma of libRmath, which the reference's horseshoe update calls
 * once per iteration (up_tau_hs, src/utils_ct.cpp:903-906); libRmath is not available here.
 * Published algorithms: series for x < a+1 and modified-Lentz continued fraction otherwise
 * (Press et al., Numerical Recipes, gser/gcf); the inverse by a safeguarded Newton iteration
 * from the Wilson-Hilferty approximation.
 * Shared by the device update kernel and by the CPU oracle (the function is pinned on its own
 * against scipy.special.gammainc / gammaincinv in tests/test_oracle_chain.py).
 */
#ifndef TPPMX_INCGAMMA_H
#define TPPMX_INCGAMMA_H
#include <math.h>

#ifdef __CUDACC__
#define TPPMX_HD __host__ __device__
#else
#define TPPMX_HD
#endif

/* P(a, x) = gamma(a, x) / Gamma(a),  a > 0, x >= 0; lga = lgamma(a).
 * For a = n/2 (the only shapes the path produces: (Q*D + 1)/2) and x >= 1/2 the upper tail has
 * a finite closed form with positive terms,
 *     a integer   : Q = e^-x sum_{k<a} x^k / k!
 *     a = m + 1/2 : Q = erfc(sqrt x) + e^-x sum_{k<m} x^(k+1/2) / Gamma(k+3/2),
 * used when P = 1 - Q >= 1e-3 (relative error then <= 1e-13); otherwise the series (x < a+1) or
 * the modified-Lentz continued fraction, which are accurate relative to a tiny P. */
TPPMX_HD static inline double tppmx_gammainc_lga(double a, double x, double lga) {
  if (!(x > 0.0)) return 0.0;
  const double a2 = 2.0 * a;
  if (x >= 0.5 && a2 == floor(a2) && a2 <= 512.0) {
    const int n2 = (int)a2;
    const double ex = exp(-x);
    double q, t;
    int m;
    if ((n2 & 1) == 0) { /* integer a */
      m = n2 / 2;
      t = 1.0;
      q = 0.0;
      for (int k = 0; k < m; k++) {
        q += t;
        t *= x / (double)(k + 1);
      }
      q *= ex;
    } else {
      m = (n2 - 1) / 2;
      const double sx = sqrt(x);
      t = 1.1283791670955126 * sx; /* 2 sqrt(x) / sqrt(pi) */
      q = 0.0;
      for (int k = 0; k < m; k++) {
        q += t;
        t *= x / ((double)k + 1.5);
      }
      q = erfc(sx) + ex * q;
    }
    if (q < 0.999) return 1.0 - q; /* else P is tiny: fall through to the relatively accurate forms */
  }
  const double lead = exp(-x + a * log(x) - lga);
  if (x < a + 1.0) {
    double ap = a, del = 1.0 / a, sum = del;
    for (int n = 0; n < 2000; n++) {
      ap += 1.0;
      del *= x / ap;
      sum += del;
      if (fabs(del) < fabs(sum) * 1e-17) break;
    }
    const double p = sum * lead;
    return p > 1.0 ? 1.0 : p;
  }
  const double tiny = 1e-300;
  double b = x + 1.0 - a, c = 1.0 / tiny, d = 1.0 / b, h = d;
  for (int i = 1; i < 2000; i++) {
    const double an = -(double)i * ((double)i - a);
    b += 2.0;
    d = an * d + b;
    if (fabs(d) < tiny) d = tiny;
    c = b + an / c;
    if (fabs(c) < tiny) c = tiny;
    d = 1.0 / d;
    const double del = d * c;
    h *= del;
    if (fabs(del - 1.0) < 1e-16) break;
  }
  const double q = lead * h;
  return q > 1.0 ? 0.0 : 1.0 - q;
}
TPPMX_HD static inline double tppmx_gammainc(double a, double x) { return tppmx_gammainc_lga(a, x, lgamma(a)); }

/* standard normal quantile, Acklam's rational approximation (relative error < 1.2e-9): only the
 * starting point of the Newton iteration below */
TPPMX_HD static inline double tppmx_qnorm_approx(double p) {
  const double a1 = -3.969683028665376e+01, a2 = 2.209460984245205e+02, a3 = -2.759285104469687e+02,
               a4 = 1.383577518672690e+02, a5 = -3.066479806614716e+01, a6 = 2.506628277459239e+00;
  const double b1 = -5.447609879822406e+01, b2 = 1.615858368580409e+02, b3 = -1.556989798598866e+02,
               b4 = 6.680131188771972e+01, b5 = -1.328068155288572e+01;
  const double c1 = -7.784894002430293e-03, c2 = -3.223964580411365e-01, c3 = -2.400758277161838e+00,
               c4 = -2.549732539343734e+00, c5 = 4.374664141464968e+00, c6 = 2.938163982698783e+00;
  const double d1 = 7.784695709041462e-03, d2 = 3.224671290700398e-01, d3 = 2.445134137142996e+00,
               d4 = 3.754408661907416e+00;
  if (p < 0.02425) {
    const double q = sqrt(-2.0 * log(p));
    return (((((c1 * q + c2) * q + c3) * q + c4) * q + c5) * q + c6) / ((((d1 * q + d2) * q + d3) * q + d4) * q + 1.0);
  }
  if (p > 1.0 - 0.02425) {
    const double q = sqrt(-2.0 * log(1.0 - p));
    return -(((((c1 * q + c2) * q + c3) * q + c4) * q + c5) * q + c6) / ((((d1 * q + d2) * q + d3) * q + d4) * q + 1.0);
  }
  const double q = p - 0.5, r = q * q;
  return (((((a1 * r + a2) * r + a3) * r + a4) * r + a5) * r + a6) * q /
         (((((b1 * r + b2) * r + b3) * r + b4) * r + b5) * r + 1.0);
}

/* x with P(a, x) = p, 0 <= p < 1.  Wilson-Hilferty (or the small-p power law) start, then
 * Newton steps on P(a, x) - p with a bisection safeguard; typically 3-5 evaluations of P. */
TPPMX_HD static inline double tppmx_gammaincinv(double a, double p) {
  if (!(p > 0.0)) return 0.0;
  if (p >= 1.0) return INFINITY;
  const double lga = lgamma(a);
  double x;
  {
    const double t = 1.0 / (9.0 * a);
    const double wh = 1.0 - t + tppmx_qnorm_approx(p) * sqrt(t);
    x = a * wh * wh * wh;
    const double xs = exp((log(p) + lga + log(a)) / a); /* P(a,x) ~ x^a / Gamma(a+1) as x -> 0 */
    if (!(wh > 0.0) || x < 0.1 * a || xs < 0.2 * a) x = xs;
  }
  double lo = 0.0, hi = INFINITY;
  for (int it = 0; it < 100; it++) {
    const double f = tppmx_gammainc_lga(a, x, lga) - p;
    if (f > 0.0) hi = x; else lo = x;
    const double dens = exp((a - 1.0) * log(x) - x - lga); /* dP/dx */
    double xn = (dens > 0.0) ? x - f / dens : x;
    if (!(xn > lo && xn < hi)) xn = (hi < INFINITY) ? 0.5 * (lo + hi) : 2.0 * x + 1.0;
    if (fabs(xn - x) <= 2e-15 * fabs(xn)) {
      x = xn;
      break;
    }
    x = xn;
  }
  return x;
}

#endif /* TPPMX_INCGAMMA_H */
