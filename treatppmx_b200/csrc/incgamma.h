/* incgamma.h — regularised incomplete gamma function and its inverse (C / CUDA dual header).
 *
 * Stands in for R::pgamma / R::qgamma of libRmath, which the reference's horseshoe update calls
 * once per iteration (up_tau_hs, src/utils_ct.cpp:903-906); libRmath is not available here.
 * Published algorithms: series for x < a+1 and modified-Lentz continued fraction otherwise
 * (Press et al., Numerical Recipes, gser/gcf); the inverse by a bracketed Newton iteration.
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

/* P(a, x) = gamma(a, x) / Gamma(a),  a > 0, x >= 0 */
TPPMX_HD static inline double tppmx_gammainc(double a, double x) {
  if (!(x > 0.0)) return 0.0;
  const double lead = exp(-x + a * log(x) - lgamma(a));
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

/* x with P(a, x) = p, 0 <= p < 1 */
TPPMX_HD static inline double tppmx_gammaincinv(double a, double p) {
  if (!(p > 0.0)) return 0.0;
  if (p >= 1.0) return INFINITY;
  double lo = 0.0, hi = a > 1.0 ? a : 1.0;
  for (int i = 0; i < 1100 && tppmx_gammainc(a, hi) < p; i++) {
    lo = hi;
    hi *= 2.0;
  }
  double x = 0.5 * (lo + hi);
  const double lga = lgamma(a);
  for (int it = 0; it < 200; it++) {
    const double f = tppmx_gammainc(a, x) - p;
    if (f > 0.0) hi = x; else lo = x;
    const double dens = exp((a - 1.0) * log(x) - x - lga); /* dP/dx */
    double xn = (dens > 0.0) ? x - f / dens : 0.5 * (lo + hi);
    if (!(xn > lo && xn < hi)) xn = 0.5 * (lo + hi);
    if (fabs(xn - x) <= 4e-16 * fabs(xn) || hi - lo <= 4e-16 * hi) {
      x = xn;
      break;
    }
    x = xn;
  }
  return x;
}

#endif /* TPPMX_INCGAMMA_H */
