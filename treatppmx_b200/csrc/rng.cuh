// rng.cuh — device side of the RNG contract in include/tppmx.h.
// Philox4x32-10 keyed by the 64-bit seed; counter = (index, sub|site<<16|arm<<24, iter, chain).
// Replaces the reference's use of R's global RNG (src/RcppExports.cpp:19; R::runif at
// src/dm_ppmx_ct.cpp:805, arma::mvnrnd at :358/:608/:844, R::rgamma at :90/:1045/:1054).
#pragma once
#include <stdint.h>

#include "fastmath.cuh"

namespace tppmx {

struct Rng {
  uint32_t k0, k1, chain;
};

__device__ __forceinline__ Rng make_rng(uint64_t seed, uint32_t chain) {
  Rng r;
  r.k0 = (uint32_t)seed;
  r.k1 = (uint32_t)(seed >> 32);
  r.chain = chain;
  return r;
}

__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                               uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0;
    const uint32_t n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}

// two uniforms in (0,1): (52 random bits + 1/2) * 2^-52
__device__ __forceinline__ void rng_u2(const Rng &g, uint32_t iter, uint32_t site, uint32_t arm,
                                       uint32_t index, uint32_t sub, double &u0, double &u1) {
  const uint4 o = philox4x32_10(index, (sub & 0xFFFFu) | ((site & 0xFFu) << 16) | ((arm & 0xFFu) << 24),
                                iter, g.chain, g.k0, g.k1);
  const unsigned long long a = (((unsigned long long)o.x) << 20) | (o.y >> 12);
  const unsigned long long b = (((unsigned long long)o.z) << 20) | (o.w >> 12);
  u0 = ((double)a + 0.5) * 0x1.0p-52;
  u1 = ((double)b + 0.5) * 0x1.0p-52;
}

__device__ __forceinline__ void rng_n2(const Rng &g, uint32_t iter, uint32_t site, uint32_t arm,
                                       uint32_t index, uint32_t sub, double &n0, double &n1) {
  double u0, u1;
  rng_u2(g, iter, site, arm, index, sub, u0, u1);
  const double r = sqrt(-2.0 * log(u0));
  double s, c;
  sincos(6.283185307179586476925286766559 * u1, &s, &c);
  n0 = r * c;
  n1 = r * s;
}

// the first normal of the block only (what the gamma sampler consumes): no sine, branch-free log
__device__ __forceinline__ double rng_n1(const Rng &g, uint32_t iter, uint32_t site, uint32_t arm,
                                         uint32_t index, uint32_t sub) {
  double u0, u1;
  rng_u2(g, iter, site, arm, index, sub, u0, u1);
  return sqrt(-2.0 * fast_log(u0)) * cos(6.283185307179586476925286766559 * u1);
}

// D standard normals: block b gives normals 2b, 2b+1
__device__ __forceinline__ void rng_normals(const Rng &g, uint32_t iter, uint32_t site,
                                            uint32_t arm, uint32_t index, int D, double *out,
                                            int stride) {
  for (int b = 0; 2 * b < D; b++) {
    double n0, n1;
    rng_n2(g, iter, site, arm, index, (uint32_t)b, n0, n1);
    out[(2 * b) * stride] = n0;
    if (2 * b + 1 < D) out[(2 * b + 1) * stride] = n1;
  }
}

// Gamma(a, 1), Marsaglia & Tsang (2000); same draw order as oracle/philox_ref.h
__device__ inline double rng_gamma(const Rng &g, uint32_t iter, uint32_t site, uint32_t arm,
                                   uint32_t index, double a) {
  double boost = 1.0;
  if (a < 1.0) {
    double u0, u1;
    rng_u2(g, iter, site, arm, index, 0xFF00u, u0, u1);
    boost = fast_exp(fast_log(u1) / a);  // u1 in (0,1): the exponent is <= 0
    a += 1.0;
  }
  const double d = a - 1.0 / 3.0;
  const double c = 1.0 / sqrt(9.0 * d);
  for (uint32_t t = 0; t < 0x7F00u; t++) {
    double u, uu;
    const double x = rng_n1(g, iter, site, arm, index, 2 * t);
    double v = 1.0 + c * x;
    if (v <= 0.0) continue;
    v = v * v * v;
    rng_u2(g, iter, site, arm, index, 2 * t + 1, u, uu);
    const double x2 = x * x;
    if (u < 1.0 - 0.0331 * x2 * x2) return boost * d * v;
    if (fast_log(u) < 0.5 * x2 + d * (1.0 - v + fast_log(v))) return boost * d * v;
  }
  return boost * d;
}

}  // namespace tppmx
