// fastmath.cuh — branch-free fp64 log / exp / lgamma for the sweep's parallel pre-pass and
// softmax.  CUDA's libdevice versions carry special-case paths (denormals, negative
// arguments, several lgamma ranges) that diverge across lanes and cost 2-4x the instructions;
// here the arguments are known to be positive normal numbers (gamma = exp(eta + beta'z),
// Stirling arguments >= 8, softmax exponents <= 0), so one straight-line path suffices.
// Accuracy (checked on the host by tests/test_fastmath.py, which restates the same formulas):
//   fast_log   <= 4e-16 relative          fast_exp <= 2.3e-16 relative on [-700, 0]
//   lgamma_pos <= 7e-15 (1 + |lgamma|)    lgamma_fast (table) <= 8e-16 (1 + |lgamma|)
//   -> far inside the 1e-10 bar on allocation probabilities
#pragma once
#include "lgammatab.h"
#include "logtab.h"

namespace tppmx {

#define TPPMX_LN2_HI 0.693147180369123816490
#define TPPMX_LN2_LO 1.90821492927058770002e-10

// Polynomial coefficients live in constant memory: a DFMA takes a constant-bank operand for
// free, whereas a 64-bit literal costs two moves per use (profiles/r1_sweep_fast_*: 8 % of the
// sweep kernel's issued instructions were UMOVs of literals).
enum {
  FC_SQRT2 = 0, FC_LN2_HI, FC_LN2_LO, FC_NLN2_HI, FC_NLN2_LO, FC_LOG2E,
  FC_LOG0,            // 9 entries: 1/19, 1/17, ..., 1/3
  FC_EXPA = FC_LOG0 + 9,   // 7 entries: 1/12!, 1/10!, ..., 1/0!
  FC_EXPB = FC_EXPA + 7,   // 7 entries: 1/13!, 1/11!, ..., 1/1!
  FC_STIR = FC_EXPB + 7,   // 7 entries: Stirling series
  FC_HLN2PI = FC_STIR + 7,
  FC_L1P,             // 4 entries: -1/6, 1/5, -1/4, 1/3 (log1p series of fast_log_tab)
  FC_COUNT = FC_L1P + 4
};
static __constant__ double FC[FC_COUNT] = {
    1.4142135623730951, TPPMX_LN2_HI, TPPMX_LN2_LO, -TPPMX_LN2_HI, -TPPMX_LN2_LO, 1.4426950408889634,
    1.0 / 19.0, 1.0 / 17.0, 1.0 / 15.0, 1.0 / 13.0, 1.0 / 11.0, 1.0 / 9.0, 1.0 / 7.0, 1.0 / 5.0, 1.0 / 3.0,
    1.0 / 479001600.0, 1.0 / 3628800.0, 1.0 / 40320.0, 1.0 / 720.0, 1.0 / 24.0, 0.5, 1.0,
    1.0 / 6227020800.0, 1.0 / 39916800.0, 1.0 / 362880.0, 1.0 / 5040.0, 1.0 / 120.0, 1.0 / 6.0, 1.0,
    1.0 / 156.0, -691.0 / 360360.0, 1.0 / 1188.0, -1.0 / 1680.0, 1.0 / 1260.0, -1.0 / 360.0, 1.0 / 12.0,
    0.91893853320467274178,
    -1.0 / 6.0, 1.0 / 5.0, -1.0 / 4.0, 1.0 / 3.0};
// {1/c_k, log c_k}, c_k = 1 + (k + 0.5)/128 (scripts/gen_logtab.py); 2 KB, read through the L1
static __device__ const double2 LOGTAB[128] = {TPPMX_LOGTAB_INIT};

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
  if (m > FC[FC_SQRT2]) {  // compiles to selects
    m *= 0.5;
    e += 1;
  }
  const double num = m - 1.0, den = m + 1.0;
  const double r = fast_rcp(den);
  double f = num * r;
  f = fma(fma(-f, den, num), r, f);
  const double f2 = f * f;
  double p = FC[FC_LOG0];
#pragma unroll
  for (int i = 1; i < 9; i++) p = fma(p, f2, FC[FC_LOG0 + i]);
  const double t = f * f2 * p;
  const double de = (double)e;
  return fma(de, FC[FC_LN2_HI], fma(de, FC[FC_LN2_LO], 2.0 * (f + t)));
}

// log(y), y positive normal, table-driven: m = mantissa in [1,2), k = its top 7 bits,
// log y = e ln2 + log c_k + log1p(m/c_k - 1), |m/c_k - 1| < 2^-8 so six series terms suffice.
// Absolute error <= 3e-16 max(1, |log y|); unlike fast_log it is NOT relatively accurate for
// y -> 1 (the table's first centre is not 1).  Used where the argument is far from 1 or only the
// absolute error matters (the two logs inside lgamma_pos).  ~20 instructions against ~36.
__device__ __forceinline__ double fast_log_tab(double y) {
  const int hi = __double2hiint(y);
  const int lo = __double2loint(y);
  const int e = (hi >> 20) - 1023;
  const double m = __hiloint2double((hi & 0x000FFFFF) | 0x3FF00000, lo);
  const double2 t = __ldg(&LOGTAB[(hi >> 13) & 0x7F]);
  const double r = fma(m, t.x, -1.0);
  double q = FC[FC_L1P];
  q = fma(q, r, FC[FC_L1P + 1]);
  q = fma(q, r, FC[FC_L1P + 2]);
  q = fma(q, r, FC[FC_L1P + 3]);
  q = fma(q, r, -0.5);
  const double l1p = fma(r * r, q, r);
  const double de = (double)e;
  return fma(de, FC[FC_LN2_HI], t.y) + fma(de, FC[FC_LN2_LO], l1p);
}

// exp(x), x <= ~1e-3 (softmax exponent after the max shift); x < -700 -> 0
__device__ __forceinline__ double fast_exp(double x) {
  const double t = fma(x, FC[FC_LOG2E], 6755399441055744.0);
  const double n = t - 6755399441055744.0;
  double r = fma(n, FC[FC_NLN2_HI], x);
  r = fma(n, FC[FC_NLN2_LO], r);
  // exp(r) = A(r^2) + r B(r^2): two half-length Horner chains (depth 8 instead of 14)
  const double r2 = r * r;
  double pa = FC[FC_EXPA];  // even coefficients 1/12!, 1/10!, ..., 1
  double pb = FC[FC_EXPB];  // odd coefficients 1/13!, 1/11!, ..., 1
#pragma unroll
  for (int i = 1; i < 7; i++) {
    pa = fma(pa, r2, FC[FC_EXPA + i]);
    pb = fma(pb, r2, FC[FC_EXPB + i]);
  }
  const double p = fma(pb, r, pa);
  const int ni = __double2loint(t);  // n as an integer (low word of the magic-number sum)
  const double out = __hiloint2double(__double2hiint(p) + (ni << 20), __double2loint(p));
  return x < -700.0 ? 0.0 : out;
}

// lgamma(x), x > 0: x < 8 is shifted up by 8 with the recurrence, then the Stirling series to
// y^-13 (truncation < 1e-15 at y >= 8)
__device__ __forceinline__ double lgamma_pos(double x) {
  const bool small = x < 8.0;
  // x(x+1)...(x+7) = u (u+6) (u+10) (u+12) with u = x (x+7): 8 operations instead of 15
  const double u = x * (x + 7.0);
  const double p = (u * (u + 6.0)) * ((u + 10.0) * (u + 12.0));
  const double y = small ? x + 8.0 : x;
  // 1/y only scales the Stirling correction (<= 1/96): one Newton step (2^-46) is enough
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(y));
  r = fma(r, fma(-y, r, 1.0), r);
  const double r2 = r * r;
  double s = FC[FC_STIR];
#pragma unroll
  for (int i = 1; i < 7; i++) s = fma(s, r2, FC[FC_STIR + i]);
  const double lg = fma(y - 0.5, fast_log_tab(y), -y) + FC[FC_HLN2PI] + s * r;
  return lg - fast_log_tab(small ? p : 1.0);
}

// lgamma(x), x > 0, table-driven (an OPTION, see lgamma_fast below): on [2^EMIN, 2^EMAX) the top four mantissa bits pick one of 16
// intervals per octave and a degree-7 polynomial in (x - midpoint) (scripts/gen_lgammatab.py;
// 64-byte rows read through the L1).  The midpoint is recovered from the bits of x and x - midpoint
// is exact, so the error is one rounding of the leading coefficient: <= 8e-16 (1 + |lgamma|),
// against <= 7e-15 for the Stirling form above at a fifth of its instructions (no log, no
// reciprocal).  Arguments outside the table (below 2.4e-4, above 4096) take the Stirling form
// through one out-of-line cold call.
static __device__ const double LGTAB[TPPMX_LGT_ROWS * 8] __attribute__((aligned(64))) = {TPPMX_LGAMMATAB_INIT};
static __device__ __noinline__ double lgamma_cold(double x) { return lgamma_pos(x); }

// table evaluation with the row index clamped into the table: always safe to execute; `oor` tells
// the caller that x lies outside the table and the value must be replaced (lgamma_cold)
__device__ __forceinline__ double lgamma_tab_core(double x, bool &oor) {
  const int hi = __double2hiint(x);
  int idx = (hi >> 16) - ((1023 + TPPMX_LGT_EMIN) << 4);
  oor = (unsigned)idx >= (unsigned)TPPMX_LGT_ROWS;
  idx = oor ? 0 : idx;
  const double t = x - __hiloint2double((hi & 0xFFFF0000) | 0x00008000, 0);
  // one 64-byte row = two 256-bit loads (LDG.E.256, sm_100): the lanes of a warp hit different rows,
  // and every load instruction costs the L1 one tag look-up per lane whatever its width
  const double *row = LGTAB + 8 * idx;
  double c0, c1, c2, c3, c4, c5, c6, c7;
#ifdef TPPMX_LGAMMA_LDG128
  {
    const double2 *r2 = reinterpret_cast<const double2 *>(row);
    const double2 c01 = __ldg(r2), c23 = __ldg(r2 + 1), c45 = __ldg(r2 + 2), c67 = __ldg(r2 + 3);
    c0 = c01.x; c1 = c01.y; c2 = c23.x; c3 = c23.y; c4 = c45.x; c5 = c45.y; c6 = c67.x; c7 = c67.y;
  }
#else
  asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(c0), "=d"(c1), "=d"(c2), "=d"(c3) : "l"(row));
  asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(c4), "=d"(c5), "=d"(c6), "=d"(c7) : "l"(row + 4));
#endif
  const double t2 = t * t;
  double pe = fma(c6, t2, c4), po = fma(c7, t2, c5);  // even / odd halves: depth 4, not 7
  pe = fma(pe, t2, c2);
  po = fma(po, t2, c3);
  pe = fma(pe, t2, c0);
  po = fma(po, t2, c1);
  return fma(po, t, pe);
}

__device__ __forceinline__ double lgamma_tab(double x) {
  bool oor;
  double l = lgamma_tab_core(x, oor);
  if (__builtin_expect(oor, 0)) l = lgamma_cold(x);
  return l;
}

// What the kernels call.  Measured on B200 (1024 chains, bench workload, profiles/r2_lgamma_ab.txt):
// the table form executes 12 % fewer instructions in the sweep kernel but is 1.7 % SLOWER
// (1.247e9 against 1.268e9 reallocations/s; with 128-bit row loads 1.203e9): every lane of a warp
// reads a different 64-byte row, each load instruction costs the L1 one tag look-up per lane, and
// the kernel is bound by the latency of its dependent chains, not by issue slots.  The Stirling
// form stays the default; -DTPPMX_LGAMMA_TABLE selects the table.
__device__ __forceinline__ double lgamma_fast(double x) {
#ifdef TPPMX_LGAMMA_TABLE
  return lgamma_tab(x);
#else
  return lgamma_pos(x);
#endif
}

// three at once: the table evaluations are straight-line and interleave (twelve loads in flight,
// three independent fma chains); one rarely taken branch repairs out-of-table arguments
__device__ __forceinline__ void lgamma_fast3(double x0, double x1, double x2, double &l0, double &l1, double &l2) {
#ifndef TPPMX_LGAMMA_TABLE
  l0 = lgamma_pos(x0);
  l1 = lgamma_pos(x1);
  l2 = lgamma_pos(x2);
  return;
#endif
  bool o0, o1, o2;
  l0 = lgamma_tab_core(x0, o0);
  l1 = lgamma_tab_core(x1, o1);
  l2 = lgamma_tab_core(x2, o2);
  if (__builtin_expect(o0 || o1 || o2, 0)) {
    if (o0) l0 = lgamma_cold(x0);
    if (o1) l1 = lgamma_cold(x1);
    if (o2) l2 = lgamma_cold(x2);
  }
}

}  // namespace tppmx
