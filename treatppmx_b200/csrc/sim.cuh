// sim.cuh — covariate similarity / cohesion / likelihood terms of one cluster score.
//
// Device restatement (not a translation) of the reference's helpers:
//   gsimconNN        src/utils_ct.cpp:382-405     gsimconNNIG + dN_IG + dinvgamma  :427-459,:161-200
//   gsimcatDM        src/utils_ct.cpp:468-495     calculate_gamma                  :497-535
// and of the per-cluster weight assembly src/dm_ppmx_ct.cpp:460-602 (occupied), :612-715 (aux).
//
// Algebra used (all fp64, differences from the reference's evaluation order are O(1e-15)):
//   Normal-Normal, auxiliary (DD=0):   with t = sumx/v2 + m0/s20, s2s(n) = 1/(n/v2 + 1/s20)
//     g(sumx,sumx2,n) = ld1 + ld2 - ld3 = A0[n] - sumx2/(2 v2) + (s2s(n)/2) t^2
//     A0[n] = -(n/2) log(2 pi v2) + dnorm(m0;0,sqrt(s20)) + log sqrt(2 pi) + (1/2) log s2s(n)
//   double dipper (DD=1):  s2ss(n) = 1/(n/v2 + 1/s2s(n))
//     g = ld1 + ld3 - ld4 = B0[n] - sumx2/(2 v2) - (s2s/2) t^2 + (s2ss/2) (sumx/v2 + t)^2
//     B0[n] = -(n/2) log(2 pi v2) - (1/2) log s2s + (1/2) log s2ss
//   so the sum over the P covariates of a cluster needs NO transcendental: the log terms
//   depend on n only and come from a table (NNRow) built once per batch.
#pragma once
#include "layout.h"
#include "fastmath.cuh"

namespace tppmx {

#define TPPMX_LN_SQRT_2PI 0.918938533204672741780329736406
#define TPPMX_LN_2PI 1.837877066409345483560659472811

struct SimConst {  // derived once per kernel from model.similparam (:220-227)
  double m0, s20, v2, k0, nu0, dirw;
  double iv2, hiv2, c0;  // 1/v2, 0.5/v2, m0/s20
  int PPMx, consim, DD, calibration, cohesion, nolik, ncat, P;
  double calscale;       // calibration==2 factor (:588-593)
};

__device__ __forceinline__ SimConst make_simconst(const DevBatch &b) {
  SimConst s;
  s.m0 = b.model.similparam[0]; s.s20 = b.model.similparam[1]; s.v2 = b.model.similparam[2];
  s.k0 = b.model.similparam[3]; s.nu0 = b.model.similparam[4]; s.dirw = b.model.similparam[6];
  s.iv2 = 1.0 / s.v2; s.hiv2 = 0.5 * s.iv2; s.c0 = (1.0 / s.s20) * s.m0;
  s.PPMx = b.model.PPMx; s.consim = b.model.consim; s.DD = (b.model.similarity == 2);
  s.calibration = b.model.calibration; s.cohesion = b.model.cohesion; s.nolik = b.model.nolik;
  s.ncat = b.ncat; s.P = b.P;
  const double pc = (double)b.P + (double)b.ncat;
  s.calscale = (b.model.coardegree == 2) ? 1.0 / sqrt(pc) : 1.0 / pc;
  return s;
}

__device__ __forceinline__ double dnorm_log(double x, double mu, double sd) {
  const double t = (x - mu) / sd;
  return -(TPPMX_LN_SQRT_2PI + 0.5 * t * t + fast_log(sd));
}

// dN_IG(mu, sig2; m, k, a, b) on the log scale, utils_ct.cpp:181-200 (+ dinvgamma :161-176)
__device__ __forceinline__ double dnig_log(double mu, double sig2, double m, double k, double a,
                                           double b) {
  return dnorm_log(mu, m, sqrt(sig2 / k)) + (a * fast_log(b) - lgamma_fast(a) - (a + 1.0) * fast_log(sig2) - b / sig2);
}

// gsimconNNIG, utils_ct.cpp:427-459 (generic path; not the default similarity)
static __device__ __noinline__ double gsim_nnig(const SimConst &c, double sumx, double sumx2, int n, int DD) {
  const double mu = 10.0, v2 = 0.1;
  const double dn = (double)n;
  const double xbar = sumx * (1.0 / dn);
  const double a0 = 0.5 * c.nu0, b0 = 0.5 * c.nu0 * c.s20;
  const double m0s = (c.k0 * c.m0 + dn * xbar) / (c.k0 + dn);
  const double k0s = c.k0 + dn;
  const double a0s = a0 + 0.5 * dn;
  const double dev = sumx2 - dn * xbar * xbar;
  const double b0s = b0 + 0.5 * dev + 0.5 * dn * c.k0 * (xbar - c.m0) * (xbar - c.m0) / (c.k0 + dn);
  const double ld1 = -0.5 * dn * log(2.0 * 3.14159265358979323846 * v2) -
                     0.5 * (1.0 / v2) * (sumx2 - 2.0 * sumx * mu + dn * mu * mu);
  const double ld3 = dnig_log(mu, v2, m0s, k0s, a0s, b0s);
  if (!DD) {
    const double ld2 = dnig_log(mu, v2, c.m0, c.k0, a0, b0);
    return ld1 + ld2 - ld3;
  }
  const double m0ss = (k0s * m0s + dn * xbar) / (k0s + dn);
  const double k0ss = k0s + dn;
  const double a0ss = a0s + 0.5 * dn;
  const double b0ss = b0s + 0.5 * dev + 0.5 * dn * k0s * (xbar - m0s) * (xbar - m0s) / (k0s + dn);
  const double ld4 = dnig_log(mu, v2, m0ss, k0ss, a0ss, b0ss);
  return ld1 + ld3 - ld4;
}

// ---- Normal-Normal-Inverse-Gamma similarity in the reference's own evaluation order (utils_ct.cpp:427-459;
// oracle/ppmx_oracle.c:oracle_gsimconNNIG is compiled with -ffp-contract=off): explicit rounded operations, no
// contraction, and every division by a quantity of n alone done as a correctly rounded quotient from its
// tabulated reciprocal (q = a r; q += (a - q d) r -- Markstein's correction, exact with RN(1/d)).
__device__ __forceinline__ double div_by_tab(double a, double d, double r) {
  const double q = __dmul_rn(a, r);
  return fma(fma(-q, d, a), r, q);
}
// gsimconNNIG(m0, k0, nu0, s20, S, SS, n, DD = 0, log) with row r of the table over n
__device__ __forceinline__ double gsim_nnig_row(const double *__restrict__ r, double S, double SS, double m0, double k0m0,
                                                double b0) {
  const double mu = 10.0;
  const double dn = r[1], kn = r[3], rkn = r[4];
  const double xbar = __dmul_rn(S, r[2]);
  const double nx = __dmul_rn(dn, xbar);
  const double m0s = div_by_tab(__dadd_rn(k0m0, nx), kn, rkn);
  const double dx = __dsub_rn(xbar, m0);
  const double b0s = __dadd_rn(__dadd_rn(b0, __dmul_rn(0.5, __dsub_rn(SS, __dmul_rn(nx, xbar)))),
                               div_by_tab(__dmul_rn(__dmul_rn(r[5], dx), dx), kn, rkn));
  const double ld1 = __dsub_rn(r[0], __dmul_rn(5.0, __dadd_rn(__dsub_rn(SS, __dmul_rn(__dmul_rn(2.0, S), mu)), r[12])));
  const double t = div_by_tab(__dsub_rn(mu, m0s), r[6], r[7]);
  const double dnl = -(__dadd_rn(__dadd_rn(TPPMX_LN_SQRT_2PI, __dmul_rn(__dmul_rn(0.5, t), t)), r[8]));
  const double ig = __dsub_rn(__dsub_rn(__dsub_rn(__dmul_rn(r[9], fast_log(b0s)), r[10]), r[11]), div_by_tab(b0s, 0.1, 10.0));
  return __dsub_rn(__dadd_rn(ld1, r[13]), __dadd_rn(dnl, ig));
}

// gsimcatDM, utils_ct.cpp:468-495, for counts cnt[c*cstride] (+1 at category `add`, -1 = none)
static __device__ __noinline__ double gsim_cat(const SimConst &s, const int *cnt, int cstride, int add, int C,
                                  int DD, bool zero_counts) {
  int sumc = 0;
  double tmp1 = 0.0, tmp2 = 0.0, tmp3 = 0.0, tmp4 = 0.0, tmp5 = 0.0, tmp6 = 0.0;
  for (int c = 0; c < C; c++) {
    int nc = zero_counts ? 0 : cnt[c * cstride];
    if (c == add) nc += 1;
    const double dn = (double)nc;
    sumc += nc;
    tmp1 += s.dirw;
    tmp2 += lgamma_fast(s.dirw);
    tmp3 += dn + s.dirw;
    tmp4 += lgamma_fast(dn + s.dirw);
    if (DD) {
      tmp5 += 2.0 * dn + s.dirw;
      tmp6 += lgamma_fast(2.0 * dn + s.dirw);
    }
  }
  double out = (lgamma_fast(tmp1) - tmp2) + (tmp4 - lgamma_fast(tmp3));
  if (DD) out = (lgamma_fast(tmp3) - tmp4) + (tmp6 - lgamma_fast(tmp5));
  if (sumc == 0) out = 0.0;
  return out;
}

// Similarity of one cluster with and without the candidate patient (src/dm_ppmx_ct.cpp:466-553
// for an occupied cluster, :616-672 for an auxiliary one, which is the "with" term of an empty
// cluster).  sumx/sumx2 point at column j of the [P][ks] SoA arrays; x/xc are the patient's
// covariates; xc_add is the categorical row added in the "with" term.
__device__ inline void sim_terms(const DevBatch &b, const SimConst &s, const double *sumx,
                                 const double *sumx2, const int *njc, int ks, int n, bool occ,
                                 const double *__restrict__ x,
                                 const int *__restrict__ xc_add, double &lgN, double &lgY) {
  lgN = 0.0;
  lgY = 0.0;
  if (!s.PPMx) return;
  const int P = s.P;
  if (s.consim == 1) {
    double tN = 0.0, tY = 0.0, uN = 0.0, uY = 0.0, ssN = 0.0, ssY = 0.0;
#pragma unroll 2
    for (int p = 0; p < P; p++) {
      const double S = occ ? sumx[(size_t)p * ks] : 0.0;
      const double SS = occ ? sumx2[(size_t)p * ks] : 0.0;
      const double xp = __ldg(x + p), x2p = __dmul_rn(xp, xp);
      const double t = fma(S, s.iv2, s.c0);
      const double S2 = S + xp, SS2 = SS + x2p;
      const double t2 = fma(S2, s.iv2, s.c0);
      tN = fma(t, t, tN);
      tY = fma(t2, t2, tY);
      ssN += SS;
      ssY += SS2;
      if (s.DD) {
        const double v = fma(S, s.iv2, t), v2 = fma(S2, s.iv2, t2);
        uN = fma(v, v, uN);
        uY = fma(v2, v2, uY);
      }
    }
    const NNRow rN = b.nn_tab[n], rY = b.nn_tab[n + 1];
    const double dP = (double)P;
    if (!s.DD) {
      lgN = dP * rN.A0 - s.hiv2 * ssN + rN.H * tN;
      lgY = dP * rY.A0 - s.hiv2 * ssY + rY.H * tY;
    } else {
      lgN = dP * rN.B0 - s.hiv2 * ssN - rN.H * tN + rN.H2 * uN;
      lgY = dP * rY.B0 - s.hiv2 * ssY - rY.H * tY + rY.H2 * uY;
    }
  } else {
    if (b.nnig_tab && !s.DD) {  // the per-n table of the fast kernel is there: one log per evaluation, the reference's operation order
      const double *rN = b.nnig_tab + (size_t)TPPMX_NNIG_ROW * n, *rY = rN + TPPMX_NNIG_ROW;
      const double k0m0 = __dmul_rn(s.k0, s.m0), b0 = __dmul_rn(__dmul_rn(0.5, s.nu0), s.s20);
      for (int p = 0; p < P; p++) {
        const double S = occ ? sumx[(size_t)p * ks] : 0.0;
        const double SS = occ ? sumx2[(size_t)p * ks] : 0.0;
        const double xp = __ldg(x + p);
        if (occ) lgN = __dadd_rn(lgN, gsim_nnig_row(rN, S, SS, s.m0, k0m0, b0));
        lgY = __dadd_rn(lgY, gsim_nnig_row(rY, __dadd_rn(S, xp), __dadd_rn(SS, __dmul_rn(xp, xp)), s.m0, k0m0, b0));
      }
    } else
    for (int p = 0; p < P; p++) {
      const double S = occ ? sumx[(size_t)p * ks] : 0.0;
      const double SS = occ ? sumx2[(size_t)p * ks] : 0.0;
      const double xp = __ldg(x + p), x2p = __dmul_rn(xp, xp);
      if (occ) lgN += gsim_nnig(s, S, SS, n, s.DD);
      lgY += gsim_nnig(s, S + xp, SS + x2p, n + 1, s.DD);
    }
  }
  for (int p = 0; p < s.ncat; p++) {
    const int C = b.catvec[p];
    const int *cnt = njc + (size_t)p * b.maxC * ks;
    if (occ) lgN += gsim_cat(s, cnt, ks, -1, C, s.DD, false);
    lgY += gsim_cat(s, cnt, ks, xc_add[p], C, s.DD, !occ);
  }
  if (!occ) lgN = 0.0;
}

// cohesion (:559-565 occupied, :676-681 auxiliary).  dV = logV(n_t,K) - logV(n_t-1,K) at idxs.
// lgamma(n+1-sigma) - lgamma(n-sigma) == log(n - sigma) exactly (Gamma(x+1) = x Gamma(x)).
__device__ __forceinline__ double cohesion_term(const SimConst &s, bool occ, int n, double sigma,
                                                double dV, double log_alpha_cc) {
  if (s.cohesion == 1) return occ ? log((double)n) : log_alpha_cc;
  return occ ? dV + log((double)n - sigma) : dV;
}

// cluster-dependent part of the DM-augmented likelihood (:570-576):
//   sum_k [ gamma_k log JJ_ik - lgamma(gamma_k) ],  gamma_k = exp(eta_k + zb_ik)
__device__ __forceinline__ double lik_term(int D, const double *eta, int estride,
                                           const double *__restrict__ zb,
                                           const double *__restrict__ ljj) {
  // branch-free fp64 exp / lgamma (fastmath.cuh): libdevice's lgamma alone is ~3x the instructions
  // and diverges across lanes; this term is most of what the generic kernel computes per step
  double w = 0.0;
  for (int k = 0; k < D; k++) {
    const double lgk = eta[(size_t)k * estride] + zb[k];
    const double g = lgk > 709.0 ? INFINITY : fast_exp(lgk);
    w += g * ljj[k] - lgamma_fast(g);
  }
  return w;
}

// ---- warp helpers -------------------------------------------------------------------------
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_incl_scan(double v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

}  // namespace tppmx
