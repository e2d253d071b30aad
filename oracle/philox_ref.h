/*
 * philox_ref.h — CPU statement of the RNG contract in include/tppmx.h (TEST INFRASTRUCTURE).
 *
 * Philox4x32-10 (Salmon, Moraes, Dror, Shaw, "Parallel random numbers: as easy as 1, 2, 3",
 * SC'11) with the published multipliers / Weyl constants.  The reference itself draws from
 * R's global RNG (src/RcppExports.cpp:19, R::runif at src/dm_ppmx_ct.cpp:805, arma::mvnrnd
 * at :358/:608/:844); that stream cannot be reproduced without R, so the oracle and the CUDA
 * library share this counter-based mapping instead and are compared in lock-step.
 */
#ifndef PHILOX_REF_H
#define PHILOX_REF_H
#include <math.h>
#include <stdint.h>

static inline void o_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
  uint32_t k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

typedef struct o_rng {
  uint64_t seed;
  uint32_t chain;
} o_rng;

/* Variate tape (ppmx_oracle.c: oracle_tape_*).  While recording, every variate the oracle
 * CONSUMES is appended in call order as (kind, value, aux): kind 1 = uniform, 2 = standard normal,
 * 3 = Gamma(shape aux, scale 1).  Draws made inside a composite draw (the Box-Muller uniforms, the
 * Marsaglia-Tsang attempts) are not recorded.  oracle/refshim replays the tape into the
 * reference's own sources, which pins the oracle's draw order to theirs.  Single-threaded use. */
extern int o_tape_on, o_tape_depth;
void o_tape_rec(int kind, double val, double aux);
#define O_TAPE_TOP() (o_tape_on && o_tape_depth == 0)

/* two uniforms in (0,1): 52 random bits + 1/2, scaled by 2^-52 -> never 0, never 1 */
static inline void o_rng_u2(const o_rng *g, uint32_t iter, uint32_t site, uint32_t arm,
                            uint32_t index, uint32_t sub, double *u0, double *u1) {
  uint32_t ctr[4] = {index, (sub & 0xFFFFu) | ((site & 0xFFu) << 16) | ((arm & 0xFFu) << 24),
                     iter, g->chain};
  uint32_t key[2] = {(uint32_t)g->seed, (uint32_t)(g->seed >> 32)};
  uint32_t o[4];
  o_philox4x32_10(ctr, key, o);
  uint64_t a = (((uint64_t)o[0]) << 20) | (o[1] >> 12);
  uint64_t b = (((uint64_t)o[2]) << 20) | (o[3] >> 12);
  *u0 = ((double)a + 0.5) * 0x1.0p-52;
  *u1 = ((double)b + 0.5) * 0x1.0p-52;
  if (O_TAPE_TOP()) o_tape_rec(1, *u0, 0.0); /* callers that also consume u1 record it themselves */
}

/* two standard normals by Box-Muller on the same block */
static inline void o_rng_n2(const o_rng *g, uint32_t iter, uint32_t site, uint32_t arm,
                            uint32_t index, uint32_t sub, double *n0, double *n1) {
  double u0, u1;
  if (o_tape_on) o_tape_depth++;
  o_rng_u2(g, iter, site, arm, index, sub, &u0, &u1);
  if (o_tape_on) o_tape_depth--;
  double r = sqrt(-2.0 * log(u0));
  double th = 6.283185307179586476925286766559 * u1;
  *n0 = r * cos(th);
  *n1 = r * sin(th);
  if (O_TAPE_TOP()) o_tape_rec(2, *n0, 0.0); /* direct callers consume n0 only */
}

/* D standard normals: block b gives normals 2b, 2b+1 (sub = sub0 + b) */
static inline void o_rng_normals(const o_rng *g, uint32_t iter, uint32_t site, uint32_t arm,
                                 uint32_t index, uint32_t sub0, int D, double *out) {
  if (o_tape_on) o_tape_depth++;
  for (int b = 0; 2 * b < D; b++) {
    double n0, n1;
    o_rng_n2(g, iter, site, arm, index, sub0 + (uint32_t)b, &n0, &n1);
    out[2 * b] = n0;
    if (2 * b + 1 < D) out[2 * b + 1] = n1;
  }
  if (o_tape_on) o_tape_depth--;
  if (O_TAPE_TOP())
    for (int k = 0; k < D; k++) o_tape_rec(2, out[k], 0.0);
}

/* Gamma(shape a, scale 1) by Marsaglia & Tsang (2000), with the U^(1/a) boost for a<1.
 * Attempt t uses block sub0+2t (normal = first Box-Muller output) and sub0+2t+1 (accept
 * uniform = u0); the boost uniform is u1 of block sub0+0xFF00.  Same draw order on device. */
static inline double o_rng_gamma_raw(const o_rng *g, uint32_t iter, uint32_t site, uint32_t arm,
                                     uint32_t index, uint32_t sub0, double a);
static inline double o_rng_gamma(const o_rng *g, uint32_t iter, uint32_t site, uint32_t arm,
                                 uint32_t index, uint32_t sub0, double a) {
  if (o_tape_on) o_tape_depth++;
  const double v = o_rng_gamma_raw(g, iter, site, arm, index, sub0, a);
  if (o_tape_on) o_tape_depth--;
  if (O_TAPE_TOP()) o_tape_rec(3, v, a);
  return v;
}
static inline double o_rng_gamma_raw(const o_rng *g, uint32_t iter, uint32_t site, uint32_t arm,
                                     uint32_t index, uint32_t sub0, double a) {
  double boost = 1.0;
  if (a < 1.0) {
    double u0, u1;
    o_rng_u2(g, iter, site, arm, index, sub0 + 0xFF00u, &u0, &u1);
    boost = exp(log(u1) / a);
    a += 1.0;
  }
  double d = a - 1.0 / 3.0;
  double c = 1.0 / sqrt(9.0 * d);
  for (uint32_t t = 0; t < 0x7F00u; t++) {
    double x, xx, u, uu;
    o_rng_n2(g, iter, site, arm, index, sub0 + 2 * t, &x, &xx);
    double v = 1.0 + c * x;
    if (v <= 0.0) continue;
    v = v * v * v;
    o_rng_u2(g, iter, site, arm, index, sub0 + 2 * t + 1, &u, &uu);
    double x2 = x * x;
    if (u < 1.0 - 0.0331 * x2 * x2) return boost * d * v;
    if (log(u) < 0.5 * x2 + d * (1.0 - v + log(v))) return boost * d * v;
  }
  return boost * d; /* unreachable in practice */
}

#endif
