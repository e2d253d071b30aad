// predict.cuh — posterior-predictive step for new patients.
//
// B200-first restatement of /root/reference/src/dm_ppmx_ct.cpp:1080-1385.  New patients are
// independent given the chain state, so one THREAD owns one (chain, arm, new patient): it
// streams over the K+CC clusters twice (pass 1: running max / normaliser, pass 2: inverse-CDF
// scan with the same weights), so no per-patient weight vector is ever stored.  Cluster
// statistics are read by all threads of a block at the same address (broadcast loads).
#pragma once
#include "layout.h"
#include "rng.cuh"
#include "sim.cuh"
#include "sweep.cuh"
#include "fastmath.cuh"

namespace tppmx {

struct PredW {  // weight of cluster j for one new patient, plus the raw similarities
  double w, gN, gY;
};

__device__ inline PredW pred_weight(const DevBatch &b, const SimConst &s, const ArmView &A, int j,
                                    const double *x, const int *xc_train,
                                    const int *xc_new, double sigma, double dV,
                                    double log_alpha_cc) {
  const bool occ = j < A.K;
  const int n = occ ? A.nj[j] : 0;
  double lgN, lgY;
  // quirk (:1157): an occupied cluster's "with" count uses the TRAINING xcat row indexed by the
  // prediction index; auxiliary clusters use xcatp (:1238).
  sim_terms(b, s, A.sumx + j, A.sumx2 + j, A.njc + j, A.ks, n, occ, x, occ ? xc_train : xc_new,
            lgN, lgY);
  PredW r;
  r.gN = occ ? lgN : lgY;
  r.gY = lgY;
  r.w = cohesion_term(s, occ, n, sigma, dV, log_alpha_cc);
  if (!(s.calibration == 1 && s.PPMx)) {
    const double d = lgY - lgN;
    r.w += (s.calibration == 2 && s.PPMx) ? s.calscale * d : d;
  }
  return r;
}

// :1366-1382 given the drawn cluster: eta_pred (the cluster's eta*, or N(Theta, Sigma) for an
// auxiliary pick), pi = softmax(eta_pred + beta' zpred) without max-shift as the reference,
// one multinomial draw, outputs in the reference's layouts.
__device__ inline void pred_emit(const DevBatch &b, const Rng &rng, const ArmView &A, int iter, int tt, int pp, int ds,
                                 int newci, int K, const double *beta, const double *Theta, const double *Sigma,
                                 size_t obase, int chain, double *pipred, double *ypred, int *predclass,
                                 double *pisum = nullptr) {
  const int D = b.D, npred = b.npred;
  double eta_pred[TPPMX_MAXD];
  if (newci <= K) {  // :1366-1370
    for (int k = 0; k < D; k++) eta_pred[k] = A.eta[(size_t)k * A.ks + newci - 1];
  } else {
    double zn[TPPMX_MAXD], L[TPPMX_MAXD * TPPMX_MAXD];
    rng_normals(rng, iter, TPPMX_SITE_PRED_ETA, tt, (uint32_t)pp, D, zn, 1);
    for (int a = 0; a < D; a++)
      for (int c = 0; c <= a; c++) {
        double sum = Sigma[a + D * c];
        for (int k = 0; k < c; k++) sum -= L[a + D * k] * L[c + D * k];
        L[a + D * c] = (a == c) ? sqrt(sum) : sum / L[c + D * c];
      }
    for (int k = 0; k < D; k++) {
      double acc = Theta[k];
      for (int r = 0; r <= k; r++) acc += L[k + D * r] * zn[r];
      eta_pred[k] = acc;
    }
  }
  // :1373-1379 softmax(eta_pred + beta' zpred) without max-shift, as the reference
  const double *zp = b.zpred + ((size_t)ds * npred + pp) * b.Q;
  double pi_row[TPPMX_MAXD], sumrow = 0.0;
  for (int k = 0; k < D; k++) {
    double lg = eta_pred[k];
    for (int q = 0; q < b.Q; q++) lg += beta[k + q * D] * zp[q];
    pi_row[k] = exp(lg);
    sumrow += pi_row[k];
  }
  for (int k = 0; k < D; k++) pi_row[k] = pi_row[k] / sumrow;
  int pick = D - 1;
  if (ypred) {  // :1381 rmultinom(1, pi); the draw has its own Philox block, so skipping it moves no other variate
    double uy, uy1, cw = 0.0;
    rng_u2(rng, iter, TPPMX_SITE_PRED_Y, tt, (uint32_t)pp, 0, uy, uy1);
    for (int k = 0; k < D; k++) {
      cw += pi_row[k];
      if (uy < cw) {
        pick = k;
        break;
      }
    }
  }
  for (int k = 0; k < D; k++) {
    const size_t o = obase + pp + (size_t)npred * (k + (size_t)D * tt);
    if (pipred) pipred[o] = pi_row[k];
    if (ypred) ypred[o] = (k == pick) ? 1.0 : 0.0;
    if (pisum) atomicAdd(pisum + (size_t)ds * npred * D * b.T + (o - obase), pi_row[k]);  // summary.cuh
  }
  if (predclass) predclass[(size_t)chain * npred * b.T + pp + (size_t)npred * tt] = newci;
}

// grid = (chains, T); threads stride over new patients
__global__ void __launch_bounds__(128) predict_kernel(DevBatch b, uint64_t seed, int iter,
                                                      double *pipred, double *ypred,
                                                      int *predclass, double *pisum, int kmin) {
  // block = (chain, arm): the arms of a chain (and the chains of a dataset) are adjacent blocks, so a dataset's new
  // patients are read from DRAM once (as grid (chains, T) every arm made its own pass: 930 MB against 250 MB of inputs)
  const int chain = blockIdx.x / b.T, tt = blockIdx.x % b.T;
  const int pp_lo = 0, pp_hi = b.npred;
  const int ds = b.chain_ds[chain];
  const SimConst s = make_simconst(b);
  const Rng rng = make_rng(seed, b.chain_id[chain]);
  const ArmView A = arm_view_global(b, chain, tt, ds);
  const int K = A.K, CC = b.CC, ntot = K + CC, D = b.D, npred = b.npred;
  if (K < kmin) return;  // block-uniform: predict_reg_kernel took this arm
  const double sigma = b.sigma[(size_t)chain * b.T + tt];
  const int idxs = b.idxs[chain];
  const double log_alpha_cc = log(b.model.alpha) - log((double)CC);
  const bool cal1 = (s.calibration == 1) && s.PPMx;
  double dV = 0.0;
  if (s.cohesion == 2 && K >= 1) {
    const size_t base = ((size_t)tt * 2) * b.ntmax;
    dV = b.logV[(base + b.ntmax + (K - 1)) * b.G + idxs] - b.logV[(base + (K - 1)) * b.G + idxs];
  }
  const double *beta = b.beta + (size_t)chain * b.Q * D;
  const double *Theta = b.Theta + ((size_t)chain * b.T + tt) * D;
  const double *Sigma = b.Sigma + ((size_t)chain * b.T + tt) * D * D;
  const size_t obase = (size_t)chain * npred * D * b.T;
  const int nc1 = b.ncat > 0 ? b.ncat : 1;

  for (int pp = pp_lo + threadIdx.x; pp < pp_hi; pp += blockDim.x) {
    const double *x = b.xp + ((size_t)ds * npred + pp) * b.P;
    const int *xc_new = b.xcatp + ((size_t)ds * npred + pp) * nc1;
    // quirk 6: training row `pp` (requires pp < nobs when ncat > 0; checked on the host)
    const int *xc_train = b.xcat + ((size_t)ds * b.nobs + (pp < b.nobs ? pp : 0)) * nc1;

    double mN = -INFINITY, sN = 0.0, mY = -INFINITY, sY = 0.0;
    if (cal1) {  // :1282-1307 log-softmax normalisers of gtilN (all) and gtilY (occupied)
      for (int j = 0; j < ntot; j++) {
        const PredW r = pred_weight(b, s, A, j, x, xc_train, xc_new, sigma, dV, log_alpha_cc);
        double m2 = fmax(mN, r.gN);
        sN = sN * exp(mN - m2) + exp(r.gN - m2);
        mN = m2;
        if (j < K) {
          m2 = fmax(mY, r.gY);
          sY = sY * exp(mY - m2) + exp(r.gY - m2);
          mY = m2;
        }
      }
    }
    const double lsN = log(sN), lsY = log(sY);
    auto weight_of = [&](int j) {
      const PredW r = pred_weight(b, s, A, j, x, xc_train, xc_new, sigma, dV, log_alpha_cc);
      if (!cal1) return r.w;
      const double lgtilNk = (r.gN - mN) - lsN;
      return (j < K) ? r.w + ((r.gY - mY) - lsY) - lgtilNk : r.w + lgtilNk;
    };

    double wmax = -INFINITY, den = 0.0;  // :1337-1345
    for (int j = 0; j < ntot; j++) {
      const double w = weight_of(j);
      const double m2 = fmax(wmax, w);
      den = den * exp(wmax - m2) + exp(w - m2);
      wmax = m2;
    }
    double uu, u1;
    rng_u2(rng, iter, TPPMX_SITE_PRED_U, tt, (uint32_t)pp, 0, uu, u1);
    int newci = ntot;  // reference quirk 7 (:1351-1360): no default; see DESIGN.md
    double cweight = 0.0;
    for (int j = 0; j < ntot; j++) {
      cweight += exp(weight_of(j) - wmax) / den;
      if (uu < cweight) {
        newci = j + 1;
        break;
      }
    }

    pred_emit(b, rng, A, iter, tt, pp, ds, newci, K, beta, Theta, Sigma, obase, chain, pipred, ypred, predclass, pisum);
  }
}

// ---- fast variant for the default switches (PPMx=1, NN auxiliary similarity, calibration 0|2,
// ncat=0), the same closed form as sweep_fast.cuh: for an occupied cluster j and a new patient
//   w_j = C0_j + C1_j * b_j + C2_j * qx,   b_j = sum_p t_jp x_p,  qx = sum_p x_p^2,
// with t_jp = sumx_jp / v2 + m0 / s20 and the three coefficients functions of (n_j, sum_p t_jp^2)
// only.  A block stages t [P][K] and the coefficients in shared memory once; a thread then needs
// P fused multiply-adds per cluster and no transcendental before the softmax.
// dynamic smem: ((P + 3) * KP + P * nth + (KP + 1) * nth) doubles, KP = clusters staged (>= K).
// kmin: arms with fewer occupied clusters were handled by predict_reg_kernel (block-uniform exit).
__global__ void __launch_bounds__(128) predict_fast_kernel(DevBatch b, uint64_t seed, int iter, int KP, int kmin,
                                                           double *pipred, double *ypred, int *predclass, double *pisum) {
  extern __shared__ double psm[];
  const int chain = blockIdx.x / b.T, tt = blockIdx.x % b.T, tid = threadIdx.x, nth = blockDim.x;
  const int pp_lo = 0, pp_hi = b.npred;
  const int ds = b.chain_ds[chain];
  const SimConst s = make_simconst(b);
  const Rng rng = make_rng(seed, b.chain_id[chain]);
  const ArmView A = arm_view_global(b, chain, tt, ds);
  const int K = A.K, CC = b.CC, D = b.D, P = b.P, npred = b.npred;
  if (K < kmin) return;
  double *ts = psm;                  // [P][KP]
  double *c0s = ts + (size_t)P * KP; // [KP] x 3
  double *c1s = c0s + KP, *c2s = c1s + KP;
  double *xs = c2s + KP;             // [P][nth]
  double *ws = xs + (size_t)P * nth; // [KP + 1][nth]
  const double sigma = b.sigma[(size_t)chain * b.T + tt];
  const int idxs = b.idxs[chain];
  const double scale = (s.calibration == 2) ? s.calscale : 1.0;
  double dV = 0.0;
  if (s.cohesion == 2 && K >= 1) {
    const size_t base = ((size_t)tt * 2) * b.ntmax;
    dV = b.logV[(base + b.ntmax + (K - 1)) * b.G + idxs] - b.logV[(base + (K - 1)) * b.G + idxs];
  }
  const double dP = (double)P;
  for (int j = tid; j < K; j += nth) {  // per-cluster coefficients
    const int n = A.nj[j];
    double a = 0.0;
    for (int p = 0; p < P; p++) {
      const double t = fma(A.sumx[(size_t)p * A.ks + j], s.iv2, s.c0);
      ts[(size_t)p * KP + j] = t;
      a = fma(t, t, a);
    }
    const NNRow r0 = b.nn_tab[n], r1 = b.nn_tab[n + 1];
    const double coh = (s.cohesion == 1) ? log((double)n) : dV + log((double)n - sigma);
    c0s[j] = coh + scale * (dP * (r1.A0 - r0.A0) + (r1.H - r0.H) * a);
    c1s[j] = scale * 2.0 * s.iv2 * r1.H;
    c2s[j] = scale * (r1.H * s.iv2 * s.iv2 - s.hiv2);
  }
  // auxiliary clusters: t = c0 for every p, n = 0
  const NNRow q1 = b.nn_tab[1];
  const double a_aux = dP * s.c0 * s.c0;
  const double coh_aux = (s.cohesion == 1) ? log(b.model.alpha) - log((double)CC) : dV;
  const double *beta = b.beta + (size_t)chain * b.Q * D;
  const double *Theta = b.Theta + ((size_t)chain * b.T + tt) * D;
  const double *Sigma = b.Sigma + ((size_t)chain * b.T + tt) * D * D;
  const size_t obase = (size_t)chain * npred * D * b.T;
  __syncthreads();

  for (int pp0 = pp_lo; pp0 < pp_hi; pp0 += nth) {
    const int pp = pp0 + tid;
    const bool on = pp < pp_hi;
    double qx = 0.0, sx = 0.0;
    if (on) {
      const double *x = b.xp + ((size_t)ds * npred + pp) * P;
      for (int p = 0; p < P; p++) {
        const double xv = x[p];
        xs[(size_t)p * nth + tid] = xv;
        qx = fma(xv, xv, qx);
        sx += xv;
      }
    }
    if (on) {
      double wmax = -INFINITY;
      for (int j = 0; j < K; j++) {
        double bj = 0.0;
        for (int p = 0; p < P; p++) bj = fma(ts[(size_t)p * KP + j], xs[(size_t)p * nth + tid], bj);
        const double w = c0s[j] + c1s[j] * bj + c2s[j] * qx;
        ws[(size_t)j * nth + tid] = w;
        wmax = fmax(wmax, w);
      }
      const double tY = a_aux + s.iv2 * (2.0 * s.c0 * sx + s.iv2 * qx);
      const double w_aux = coh_aux + scale * (dP * q1.A0 - s.hiv2 * qx + q1.H * tY);
      wmax = fmax(wmax, w_aux);
      double den = 0.0;
      for (int j = 0; j < K; j++) {
        const double e = exp(ws[(size_t)j * nth + tid] - wmax);
        ws[(size_t)j * nth + tid] = e;
        den += e;
      }
      const double e_aux = exp(w_aux - wmax);
      for (int m = 0; m < CC; m++) den += e_aux;
      double uu, u1;
      rng_u2(rng, iter, TPPMX_SITE_PRED_U, tt, (uint32_t)pp, 0, uu, u1);
      int newci = K + CC;  // quirk 7: see predict_kernel
      double cw = 0.0;
      for (int j = 0; j < K + CC; j++) {
        cw += ((j < K) ? ws[(size_t)j * nth + tid] : e_aux) / den;
        if (uu < cw) {
          newci = j + 1;
          break;
        }
      }
      pred_emit(b, rng, A, iter, tt, pp, ds, newci, K, beta, Theta, Sigma, obase, chain, pipred, ypred, predclass, pisum);
    }
  }
}

// ---- register variant of the fast kernel: the shape the package produces (P <= 16 covariates,
// D <= 3 outcome categories) with at most KR occupied clusters in the arm (decided per block: an
// arm with more is left to predict_fast_kernel, launched behind this one with kmin = KR + 1).
// The new patient's covariates live in registers and its K weights in KR x 128 doubles of shared
// memory (8 KB against the 53 KB predict_fast_kernel needs for kcap + 1 of them), so several
// blocks share an SM;
// everything that does not depend on the patient is hoisted to the block prologue: t_jp, the
// three weight coefficients, eta*, beta, Theta and the Cholesky factor of Sigma (an auxiliary
// pick costs one matrix-vector product instead of a per-thread factorisation that a whole warp
// waits for).  Softmax by fast_exp with the division-free draw of the sweep kernel.
// Config 5 (10 000 new patients x 3 arms x 1024 chains): 5.5 ms -> see DESIGN.md.
#define TPPMX_PRED_KR 16
template <int PR, bool CAT>
__global__ void __launch_bounds__(128, 4) predict_reg_kernel(DevBatch b, uint64_t seed, int iter, double *pipred,
                                                          double *ypred, int *predclass, double *pisum,
                                                          const double *cat_lw, const double *cat_ln) {
  constexpr int KR = TPPMX_PRED_KR, DR = 3;
  // categorical covariates (cat_lw != nullptr): counts n_jc [ncat * maxC][KR] and sizes n_j [KR] in dynamic shared
  // memory; the Dirichlet-multinomial "with minus without" term of covariate p is log(n_jc* + w) - log(n_j + C_p w)
  // from the two tables (as the CAT instantiation of sweep_fast.cuh)
  extern __shared__ int pred_dyn[];
  const int ncat = CAT ? b.ncat : 0, ncC = ncat * b.maxC;
  int *njcs = pred_dyn, *njs = pred_dyn + ncC * KR;
  __shared__ double ts[PR][KR];
  __shared__ double cs[3][KR];
  __shared__ double etas[DR][KR];
  __shared__ double Ls[DR * DR], Ths[DR], betas[DR * TPPMX_MAXQ];
  __shared__ double ws[KR][128];  // the thread's weights, column tid
  const int chain = blockIdx.x / b.T, tt = blockIdx.x % b.T, tid = threadIdx.x, nth = blockDim.x;
  const int pp_lo = 0, pp_hi = b.npred;
  const int ds = b.chain_ds[chain];
  const ArmView A = arm_view_global(b, chain, tt, ds);
  const int K = A.K;
  if (K > KR) return;  // block-uniform: predict_fast_kernel takes this arm
  const SimConst s = make_simconst(b);
  const Rng rng = make_rng(seed, b.chain_id[chain]);
  const int CC = b.CC, D = b.D, P = b.P, Q = b.Q, npred = b.npred;
  const double sigma = b.sigma[(size_t)chain * b.T + tt];
  const int idxs = b.idxs[chain];
  const double scale = (s.calibration == 2) ? s.calscale : 1.0;
  double dV = 0.0;
  if (s.cohesion == 2 && K >= 1) {
    const size_t base = ((size_t)tt * 2) * b.ntmax;
    dV = b.logV[(base + b.ntmax + (K - 1)) * b.G + idxs] - b.logV[(base + (K - 1)) * b.G + idxs];
  }
  const double dP = (double)P;
  for (int j = tid; j < KR; j += nth) {  // per-cluster coefficients (as predict_fast_kernel)
    double a = 0.0;
    for (int p = 0; p < PR; p++) {
      const double t = (j < K && p < P) ? fma(A.sumx[(size_t)p * A.ks + j], s.iv2, s.c0) : 0.0;
      ts[p][j] = t;
      a = fma(t, t, a);
    }
    if (j < K) {
      const int n = A.nj[j];
      const NNRow r0 = b.nn_tab[n], r1 = b.nn_tab[n + 1];
      const double coh = (s.cohesion == 1) ? log((double)n) : dV + log((double)n - sigma);
      cs[0][j] = coh + scale * (dP * (r1.A0 - r0.A0) + (r1.H - r0.H) * a);
      cs[1][j] = scale * 2.0 * s.iv2 * r1.H;
      cs[2][j] = scale * (r1.H * s.iv2 * s.iv2 - s.hiv2);
      for (int k = 0; k < D; k++) etas[k][j] = A.eta[(size_t)k * A.ks + j];
    }
  }
  if constexpr (CAT) {
    for (int e = tid; e < ncC * KR; e += nth) {
      const int pc = e / KR, j = e - pc * KR;
      njcs[e] = (j < K) ? A.njc[(size_t)pc * A.ks + j] : 0;
    }
    for (int j = tid; j < KR; j += nth) njs[j] = (j < K) ? A.nj[j] : 0;
  }
  double cat_aux = 0.0;  // an auxiliary cluster: counts zero, sum_p [log w - log(C_p w)]
  if constexpr (CAT)
    for (int p = 0; p < ncat; p++) cat_aux += cat_lw[0] - cat_ln[(size_t)p * (b.ntmax + 2)];
  if (tid == 0) {  // chol_lower(Sigma) as pred_emit forms it, Theta
    const double *Sigma = b.Sigma + ((size_t)chain * b.T + tt) * D * D;
    for (int a = 0; a < D; a++)
      for (int c = 0; c <= a; c++) {
        double sum = Sigma[a + D * c];
        for (int k = 0; k < c; k++) sum -= Ls[a + D * k] * Ls[c + D * k];
        Ls[a + D * c] = (a == c) ? sqrt(sum) : sum / Ls[c + D * c];
      }
    for (int k = 0; k < D; k++) Ths[k] = b.Theta[((size_t)chain * b.T + tt) * D + k];
  }
  for (int e = tid; e < Q * D; e += nth) betas[e] = b.beta[(size_t)chain * Q * D + e];
  const NNRow q1 = b.nn_tab[1];
  const double a_aux = dP * s.c0 * s.c0;
  const double coh_aux = (s.cohesion == 1) ? log(b.model.alpha) - log((double)CC) : dV;
  const size_t obase = (size_t)chain * npred * D * b.T;
  __syncthreads();

  for (int pp = pp_lo + tid; pp < pp_hi; pp += nth) {
    double x[PR], qx = 0.0, sx = 0.0;
    // the first prognostic covariates travel with the predictive ones (their loads would otherwise sit, one
    // L2 round trip each, on the tail of the thread's dependent chain)
    const double *zp = b.zpred + ((size_t)ds * npred + pp) * Q;
    double zq[4];
#pragma unroll
    for (int q = 0; q < 4; q++) zq[q] = (q < Q) ? zp[q] : 0.0;
    {
      const double *xr = b.xp + ((size_t)ds * npred + pp) * P;
#pragma unroll
      for (int p = 0; p < PR; p++) {
        x[p] = (p < P) ? xr[p] : 0.0;
        qx = fma(x[p], x[p], qx);
        sx += x[p];
      }
    }
    double wmax;
    {
      const double tY = a_aux + s.iv2 * (2.0 * s.c0 * sx + s.iv2 * qx);
      wmax = coh_aux + scale * (dP * q1.A0 - s.hiv2 * qx + q1.H * tY + cat_aux);  // the auxiliary clusters' weight
    }
    // quirk (:1157): an occupied cluster's categorical "with" count uses the TRAINING row indexed by the prediction index
    const int *xct = b.xcat + ((size_t)ds * b.nobs + (pp < b.nobs ? pp : 0)) * (ncat > 0 ? ncat : 1);
    const double w_aux = wmax;
#pragma unroll 1
    for (int j = 0; j < K; j++) {
      double bj = 0.0;
#pragma unroll
      for (int p = 0; p < PR; p++) bj = fma(ts[p][j], x[p], bj);
      double w = cs[0][j] + cs[1][j] * bj + cs[2][j] * qx;
      if constexpr (CAT) {
        double cj = 0.0;
        for (int p = 0; p < ncat; p++)
          cj += cat_lw[njcs[(p * b.maxC + xct[p]) * KR + j]] - cat_ln[(size_t)p * (b.ntmax + 2) + njs[j]];
        w += scale * cj;
      }
      ws[j][tid] = w;
      wmax = fmax(wmax, w);
    }
    double den = 0.0;
#pragma unroll 1
    for (int j = 0; j < K; j++) {
      const double e = fast_exp(ws[j][tid] - wmax);
      ws[j][tid] = e;
      den += e;
    }
    const double e_aux = fast_exp(w_aux - wmax);
    for (int m = 0; m < CC; m++) den += e_aux;
    double uu, u1;
    rng_u2(rng, iter, TPPMX_SITE_PRED_U, tt, (uint32_t)pp, 0, uu, u1);
    // :1351-1360 first j with uu < cumsum(e)/den, i.e. uu*den < cumsum(e); default K + CC (quirk 7)
    const double thr = uu * den;
    int newci = K + CC;
    bool found = false;
    double cw = 0.0;
#pragma unroll 1
    for (int j = 0; j < K; j++) {
      cw += ws[j][tid];
      if (!found && thr < cw) {
        newci = j + 1;
        found = true;
      }
    }
    for (int m = 0; m < CC; m++) {
      cw += e_aux;
      if (!found && thr < cw) {
        newci = K + m + 1;
        found = true;
      }
    }
    // :1366-1382 eta_pred, pi = softmax(eta_pred + beta' zpred) without max-shift, one multinomial draw
    double ep[DR];
    if (newci <= K) {
#pragma unroll
      for (int k = 0; k < DR; k++) ep[k] = (k < D) ? etas[k][newci - 1] : 0.0;
    } else {
      // the D <= 3 standard normals of rng_normals (block b -> normals 2b, 2b+1), with the branch-free log and
      // sincospi (no large-argument path): most warps enter here for some lane, so the branch is on every
      // warp's path (profiles/r2_predict_reg_*: a quarter of the kernel's instructions were these draws)
      double zn[4];
      {
        double u0, u1, sn, cs;
        rng_u2(rng, iter, TPPMX_SITE_PRED_ETA, tt, (uint32_t)pp, 0, u0, u1);
        const double r = sqrt(-2.0 * fast_log(u0));
        sincospi(2.0 * u1, &sn, &cs);
        zn[0] = r * cs;
        zn[1] = r * sn;
        zn[2] = 0.0;
        if (D > 2) {
          rng_u2(rng, iter, TPPMX_SITE_PRED_ETA, tt, (uint32_t)pp, 1, u0, u1);
          zn[2] = sqrt(-2.0 * fast_log(u0)) * cospi(2.0 * u1);
        }
      }
#pragma unroll
      for (int k = 0; k < DR; k++) {
        double acc = (k < D) ? Ths[k] : 0.0;
        for (int r = 0; r <= k && k < D; r++) acc += Ls[k + D * r] * zn[r];
        ep[k] = acc;
      }
    }
    double pr[DR], sumrow = 0.0;
#pragma unroll
    for (int k = 0; k < DR; k++) {
      pr[k] = 0.0;
      if (k < D) {
        double lg = ep[k];
#pragma unroll
        for (int q = 0; q < 4; q++)
          if (q < Q) lg += betas[k + q * D] * zq[q];
        for (int q = 4; q < Q; q++) lg += betas[k + q * D] * zp[q];
        pr[k] = fast_exp(lg);
        sumrow += pr[k];
      }
    }
    const double rs = fast_rcp(sumrow);
#pragma unroll
    for (int k = 0; k < DR; k++) pr[k] *= rs;
    int pick = D - 1;
    if (ypred) {  // block-uniform; the outcome draw has its own Philox block (skipping it moves no other variate)
      double uy, uy1, cwy = 0.0;
      rng_u2(rng, iter, TPPMX_SITE_PRED_Y, tt, (uint32_t)pp, 0, uy, uy1);
      bool got = false;
#pragma unroll
      for (int k = 0; k < DR; k++)
        if (k < D) {
          cwy += pr[k];
          if (!got && uy < cwy) {
            pick = k;
            got = true;
          }
        }
    }
#pragma unroll
    for (int k = 0; k < DR; k++)
      if (k < D) {
        const size_t o = obase + pp + (size_t)npred * (k + (size_t)D * tt);
        if (pipred) pipred[o] = pr[k];
        if (ypred) ypred[o] = (k == pick) ? 1.0 : 0.0;
        // summary.cuh: posterior-predictive sums per replicate dataset, accumulated here (one RED.F64 per value)
        // instead of a per-chain cube written out and read back by a second kernel
        if (pisum) atomicAdd(pisum + (size_t)ds * npred * D * b.T + (o - obase), pr[k]);
      }
    if (predclass) predclass[(size_t)chain * npred * b.T + pp + (size_t)npred * tt] = newci;
  }
}

}  // namespace tppmx
