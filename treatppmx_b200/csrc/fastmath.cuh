// fastmath.cuh — branch-free fp64 log / exp / lgamma for the sweep's parallel pre-pass and
// softmax.  CUDA's libdevice versions carry special-case paths (denormals, negative
// arguments, several lgamma ranges) that diverge across lanes and cost 2-4x the instructions;
// here the arguments are known to be positive normal numbers (gamma = exp(eta + beta'z),
// Stirling arguments >= 8, softmax exponents <= 0), so one straight-line path suffices.
// Accuracy (checked on the host by tests/test_fastmath.py, which restates the same formulas):
//   fast_log   <= 4e-16 relative          fast_exp <= 2.3e-16 relative on [-700, 0]
//   lgamma_pos <= 7e-15 (1 + |lgamma|)    -> far inside the 1e-10 bar on allocation probabilities
#pragma once

namespace tppmx {

#define TPPMX_LN2_HI 0.693147180369123816490
#define TPPMX_LN2_LO 1.90821492927058770002e-10

// 1/a, a positive normal: hardware seed (2^-23) + two Newton steps
__device__ __forceinline__ double fast_rcp(double a) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
  double e = fma(-a, r, 1.0);
  r = fma(r, e, r);
  e = fma(-a, r, 1.0);
  r = fma(r, e, r);
  return r;
}

// log(y), y positive normal
__device__ __forceinline__ double fast_log(double y) {
  int hi = __double2hiint(y);
  const int lo = __double2loint(y);
  int e = (hi >> 20) - 1023;
  double m = __hiloint2double((hi & 0x000FFFFF) | 0x3FF00000, lo);
  if (m > 1.4142135623730951) {  // compiles to selects
    m *= 0.5;
    e += 1;
  }
  const double num = m - 1.0, den = m + 1.0;
  const double r = fast_rcp(den);
  double f = num * r;
  f = fma(fma(-f, den, num), r, f);
  const double f2 = f * f;
  double p = 1.0 / 19.0;
  p = fma(p, f2, 1.0 / 17.0);
  p = fma(p, f2, 1.0 / 15.0);
  p = fma(p, f2, 1.0 / 13.0);
  p = fma(p, f2, 1.0 / 11.0);
  p = fma(p, f2, 1.0 / 9.0);
  p = fma(p, f2, 1.0 / 7.0);
  p = fma(p, f2, 1.0 / 5.0);
  p = fma(p, f2, 1.0 / 3.0);
  const double t = f * f2 * p;
  const double de = (double)e;
  return fma(de, TPPMX_LN2_HI, fma(de, TPPMX_LN2_LO, 2.0 * (f + t)));
}

// exp(x), x <= ~1e-3 (softmax exponent after the max shift); x < -700 -> 0
__device__ __forceinline__ double fast_exp(double x) {
  const double t = fma(x, 1.4426950408889634, 6755399441055744.0);
  const double n = t - 6755399441055744.0;
  double r = fma(n, -TPPMX_LN2_HI, x);
  r = fma(n, -TPPMX_LN2_LO, r);
  // exp(r) = A(r^2) + r B(r^2): two half-length Horner chains (depth 8 instead of 14)
  const double r2 = r * r;
  double pa = 1.0 / 479001600.0;  // even coefficients 1/12!, 1/10!, ..., 1
  pa = fma(pa, r2, 1.0 / 3628800.0);
  pa = fma(pa, r2, 1.0 / 40320.0);
  pa = fma(pa, r2, 1.0 / 720.0);
  pa = fma(pa, r2, 1.0 / 24.0);
  pa = fma(pa, r2, 0.5);
  pa = fma(pa, r2, 1.0);
  double pb = 1.0 / 6227020800.0;  // odd coefficients 1/13!, 1/11!, ..., 1
  pb = fma(pb, r2, 1.0 / 39916800.0);
  pb = fma(pb, r2, 1.0 / 362880.0);
  pb = fma(pb, r2, 1.0 / 5040.0);
  pb = fma(pb, r2, 1.0 / 120.0);
  pb = fma(pb, r2, 1.0 / 6.0);
  pb = fma(pb, r2, 1.0);
  const double p = fma(pb, r, pa);
  const int ni = __double2loint(t);  // n as an integer (low word of the magic-number sum)
  const double out = __hiloint2double(__double2hiint(p) + (ni << 20), __double2loint(p));
  return x < -700.0 ? 0.0 : out;
}

// lgamma(x), x > 0: x < 8 is shifted up by 8 with the recurrence, then the Stirling series to
// y^-13 (truncation < 1e-15 at y >= 8)
__device__ __forceinline__ double lgamma_pos(double x) {
  const bool small = x < 8.0;
  double p = x * (x + 1.0);
  p *= (x + 2.0) * (x + 3.0);
  double q = (x + 4.0) * (x + 5.0);
  q *= (x + 6.0) * (x + 7.0);
  p *= q;
  const double y = small ? x + 8.0 : x;
  const double r = fast_rcp(y), r2 = r * r;
  double s = 1.0 / 156.0;
  s = fma(s, r2, -691.0 / 360360.0);
  s = fma(s, r2, 1.0 / 1188.0);
  s = fma(s, r2, -1.0 / 1680.0);
  s = fma(s, r2, 1.0 / 1260.0);
  s = fma(s, r2, -1.0 / 360.0);
  s = fma(s, r2, 1.0 / 12.0);
  const double lg = fma(y - 0.5, fast_log(y), -y) + 0.91893853320467274178 + s * r;
  return lg - fast_log(small ? p : 1.0);
}

}  // namespace tppmx
