// sweep.cuh — the collapsed-Gibbs reallocation step and sweep kernels.
//
// B200-first restatement of /root/reference/src/dm_ppmx_ct.cpp:347-860.  One warp owns one
// (chain, arm): arms of a chain never touch each other's state inside the sweep (:349-860
// only reads/writes curr_clu(tt,.), nj_curr(tt,.), sumx(tt,.), eta_star_curr.slice(tt),
// eta_empty.slice(tt) and the rows of loggamma of arm tt), so the T arms of a chain run
// concurrently; the patient loop inside an arm is the true sequential dependence and stays
// sequential.  Within a step lane j scores cluster j (occupied clusters first, then the CC
// auxiliary ones), reading column j of the structure-of-arrays statistics; max / sum / prefix
// scan over clusters are warp shuffles; the categorical draw is a ballot on the scanned CDF.
#pragma once
#include "layout.h"
#include "rng.cuh"
#include "sim.cuh"

namespace tppmx {

// pointers of one (chain, arm); may point to global memory or to a shared-memory stage
struct ArmView {
  double *sumx, *sumx2;  // [P][ks]
  int *njc;              // [ncat][maxC][ks]
  int *nj;               // [ks]
  double *eta;           // [D][ks]
  double *etae;          // [D][CC]
  int *labels;           // [nt] 1-based
  double *w, *gN, *gY;   // scratch [ks+CC]
  int ks;                // cluster stride (capacity)
  int K;                 // occupied clusters (nclu_curr)
  int nt;                // patients in the arm
};

struct PatientIn {  // read-only inputs of the patient being reallocated
  const double *x;       // [P]
  const int *xc;         // [ncat]
  const double *zb;      // [D]
  const double *ljj;     // [D]
  double cpat;           // -sum_k JJ_ik (ss_i+1) - sum_k 1{y_ik != 1} log JJ_ik   (:572-575)
};

__device__ __forceinline__ ArmView arm_view_global(const DevBatch &b, int chain, int tt, int ds) {
  ArmView A;
  const size_t P1 = b.P > 0 ? b.P : 1;
  const size_t nc1 = b.ncat > 0 ? (size_t)b.ncat * b.maxC : 1;
  A.ks = b.kcap;
  A.sumx = b.sumx + ((size_t)chain * b.T + tt) * P1 * b.kcap;
  A.sumx2 = b.sumx2 + ((size_t)chain * b.T + tt) * P1 * b.kcap;
  A.njc = b.njc + ((size_t)chain * b.T + tt) * nc1 * b.kcap;
  A.nj = b.nj + ((size_t)chain * b.T + tt) * b.kcap;
  A.eta = b.eta + ((size_t)chain * b.T + tt) * b.D * b.kcap;
  A.etae = b.etae + ((size_t)chain * b.T + tt) * b.D * b.CC;
  A.labels = b.labels + ((size_t)chain * b.T + tt) * b.ntmax;
  double *wb = b.wbuf + ((size_t)chain * b.T + tt) * 3 * (b.kcap + b.CC);
  A.w = wb;
  A.gN = wb + (b.kcap + b.CC);
  A.gY = wb + 2 * (b.kcap + b.CC);
  A.K = b.nclu[(size_t)chain * b.T + tt];
  A.nt = b.arm_n[(size_t)ds * b.T + tt];
  return A;
}

// ---- :368-457  take patient out of its cluster ----------------------------------------------
__device__ inline void remove_patient(const DevBatch &b, const SimConst &s, ArmView &A, int it,
                                      const PatientIn &pt, int lane) {
  const int ks = A.ks;
  const int zi = A.labels[it] - 1;
  const int nzi = A.nj[zi];
  __syncwarp();
  if (nzi > 1) {  // :370-383
    if (lane == 0) A.nj[zi] = nzi - 1;
    if (s.PPMx) {
      for (int p = lane; p < s.P; p += 32) {
        A.sumx[(size_t)p * ks + zi] = __dsub_rn(A.sumx[(size_t)p * ks + zi], pt.x[p]);
        A.sumx2[(size_t)p * ks + zi] = __dsub_rn(A.sumx2[(size_t)p * ks + zi], __dmul_rn(pt.x[p], pt.x[p]));
      }
      for (int p = lane; p < s.ncat; p += 32) A.njc[((size_t)p * b.maxC + pt.xc[p]) * ks + zi] -= 1;
    }
  } else {  // singleton :384-457
    const int iaux = zi + 1;
    const int K = A.K;
    if (iaux < K) {  // swap with the last cluster :387-423
      for (int ii = lane; ii < A.nt; ii += 32)
        if (A.labels[ii] == K) A.labels[ii] = iaux;
      __syncwarp();
      if (lane == 0) {
        A.labels[it] = K;
        A.nj[iaux - 1] = A.nj[K - 1];
        A.nj[K - 1] = 1;
      }
      for (int k = lane; k < b.D; k += 32) {
        const double t = A.eta[(size_t)k * ks + iaux - 1];
        A.eta[(size_t)k * ks + iaux - 1] = A.eta[(size_t)k * ks + K - 1];
        A.eta[(size_t)k * ks + K - 1] = t;
      }
      if (s.PPMx) {
        for (int p = lane; p < s.P; p += 32) {
          double t = A.sumx[(size_t)p * ks + iaux - 1];
          A.sumx[(size_t)p * ks + iaux - 1] = A.sumx[(size_t)p * ks + K - 1];
          A.sumx[(size_t)p * ks + K - 1] = t;
          t = A.sumx2[(size_t)p * ks + iaux - 1];
          A.sumx2[(size_t)p * ks + iaux - 1] = A.sumx2[(size_t)p * ks + K - 1];
          A.sumx2[(size_t)p * ks + K - 1] = t;
        }
        for (int pc = lane; pc < s.ncat * b.maxC; pc += 32) {
          const int t = A.njc[(size_t)pc * ks + iaux - 1];
          A.njc[(size_t)pc * ks + iaux - 1] = A.njc[(size_t)pc * ks + K - 1];
          A.njc[(size_t)pc * ks + K - 1] = t;
        }
      }
      __syncwarp();
    }
    // :425-440 the singleton now sits in slot K-1: empty it.  (The reference leaves the
    // rounding residue sumx - x there and later adds onto it, :850; so do we.)
    if (lane == 0) A.nj[K - 1] -= 1;
    if (s.PPMx) {
      for (int p = lane; p < s.P; p += 32) {
        A.sumx[(size_t)p * ks + K - 1] = __dsub_rn(A.sumx[(size_t)p * ks + K - 1], pt.x[p]);
        A.sumx2[(size_t)p * ks + K - 1] = __dsub_rn(A.sumx2[(size_t)p * ks + K - 1], __dmul_rn(pt.x[p], pt.x[p]));
      }
      for (int p = lane; p < s.ncat; p += 32) A.njc[((size_t)p * b.maxC + pt.xc[p]) * ks + K - 1] -= 1;
    }
    A.K = K - 1;
    // :442-454 recomputes loggamma of the arm from (eta, beta); labels and eta columns were
    // swapped together, so the values are unchanged -> nothing to do on device.
  }
  __syncwarp();
}

// ---- :460-802  score K+CC clusters, normalise -------------------------------------------------
// Leaves exp(w - max) in A.w[0..K+CC) when more than one 32-wide chunk is needed; returns this
// lane's value of the first chunk, the max and the normaliser.  logw_out/prob_out optional.
struct ScoreOut {
  double e0;     // exp(w_j - max) of cluster j = lane (first chunk)
  double den;    // sum_j exp(w_j - max)
  int ntot;
};

__device__ inline ScoreOut score_clusters(const DevBatch &b, const SimConst &s, const ArmView &A,
                                          const PatientIn &pt, const int *xc_occ, double sigma,
                                          double dV, bool with_lik, int lane, double *logw_out,
                                          double *prob_out) {
  const int K = A.K, CC = b.CC, ntot = K + CC, ks = A.ks;
  const double log_alpha_cc = log(b.model.alpha) - log((double)CC);
  const bool multi = ntot > 32;
  const bool cal1 = (s.calibration == 1) && s.PPMx;
  double w0 = -INFINITY, wmax = -INFINITY;
  double mN = -INFINITY, mY = -INFINITY;

  for (int c0 = 0; c0 < ntot; c0 += 32) {
    const int j = c0 + lane;
    const bool act = j < ntot, occ = j < K;
    double w = -INFINITY, lgN = 0.0, lgY = 0.0;
    if (act) {
      const int n = occ ? A.nj[j] : 0;
      sim_terms(b, s, A.sumx + j, A.sumx2 + j, A.njc + j, ks, n, occ, pt.x,
                occ ? xc_occ : pt.xc, lgN, lgY);
      w = cohesion_term(s, occ, n, sigma, dV, log_alpha_cc);
      if (!cal1) {
        const double d = lgY - lgN;  // == lgcondraw + lgcatdraw for an auxiliary cluster
        w += (s.calibration == 2 && s.PPMx) ? s.calscale * d : d;
      }
      if (with_lik) {
        const double *ecol = occ ? (A.eta + j) : (A.etae + (j - K));
        w += lik_term(b.D, ecol, occ ? ks : CC, pt.zb, pt.ljj) + pt.cpat;
      }
    }
    if (cal1) {  // :717-788 needs softmax-normalised similarities over all clusters first
      if (act) {
        A.gN[j] = occ ? lgN : lgY;  // gtilN == gtilY == lgcondraw+lgcatdraw for aux (:670-671)
        A.gY[j] = lgY;
        A.w[j] = w;
      }
      mN = fmax(mN, act ? (occ ? lgN : lgY) : -INFINITY);
      mY = fmax(mY, occ ? lgY : -INFINITY);
    } else {
      if (multi && act) A.w[j] = w;
      if (c0 == 0) w0 = w;
      wmax = fmax(wmax, w);
    }
  }

  if (cal1) {
    __syncwarp();
    mN = warp_max(mN);
    mY = warp_max(mY);
    double sgN = 0.0, sgY = 0.0;
    for (int c0 = 0; c0 < ntot; c0 += 32) {
      const int j = c0 + lane;
      if (j < ntot) sgN += fast_exp(A.gN[j] - mN);
      if (j < K) sgY += fast_exp(A.gY[j] - mY);
    }
    sgN = warp_sum(sgN);
    sgY = warp_sum(sgY);
    const double lsN = log(sgN), lsY = log(sgY);
    for (int c0 = 0; c0 < ntot; c0 += 32) {
      const int j = c0 + lane;
      double w = -INFINITY;
      if (j < ntot) {
        const double lgtilNk = (A.gN[j] - mN) - lsN;
        if (j < K) w = A.w[j] + ((A.gY[j] - mY) - lsY) - lgtilNk;
        else w = A.w[j] + lgtilNk;
        A.w[j] = w;
      }
      if (c0 == 0) w0 = w;
      wmax = fmax(wmax, w);
    }
    __syncwarp();
  }

  wmax = warp_max(wmax);
  ScoreOut o;
  o.ntot = ntot;
  double den = 0.0;
  if (!multi) {
    if (logw_out && lane < ntot) logw_out[lane] = w0;
    o.e0 = (lane < ntot) ? fast_exp(w0 - wmax) : 0.0;
    den = o.e0;
  } else {
    __syncwarp();
    o.e0 = 0.0;
    for (int c0 = 0; c0 < ntot; c0 += 32) {
      const int j = c0 + lane;
      if (j < ntot) {
        const double w = A.w[j];
        if (logw_out) logw_out[j] = w;
        const double e = fast_exp(w - wmax);
        A.w[j] = e;
        den += e;
      }
    }
    __syncwarp();
  }
  den = warp_sum(den);
  o.den = den;
  if (prob_out) {
    if (!multi) {
      if (lane < ntot) prob_out[lane] = o.e0 / den;
    } else {
      for (int j = lane; j < ntot; j += 32) prob_out[j] = A.w[j] / den;
    }
  }
  return o;
}

// inverse-CDF draw on the normalised weights (:805-814); returns newci in 1..ntot
__device__ inline int draw_cluster(const ArmView &A, const ScoreOut &o, double uu, int lane) {
  int newci = o.ntot;  // :807 default when rounding leaves uu >= total
  double carry = 0.0;
  for (int c0 = 0; c0 < o.ntot; c0 += 32) {
    const int j = c0 + lane;
    double p = 0.0;
    if (j < o.ntot) p = ((o.ntot > 32) ? A.w[j] : o.e0) / o.den;
    const double cum = carry + warp_incl_scan(p, lane);
    const unsigned hit = __ballot_sync(0xffffffffu, (j < o.ntot) && (uu < cum));
    if (hit) {
      newci = c0 + __ffs(hit);
      break;
    }
    carry = __shfl_sync(0xffffffffu, cum, 31);
  }
  return newci;
}

// ---- one full reallocation (:364-858) ---------------------------------------------------------
// returns false if the arm ran out of cluster capacity
__device__ inline bool realloc_step(const DevBatch &b, const SimConst &s, ArmView &A,
                                    const Rng &rng, int chain, int tt, int it, int i, int iter,
                                    const PatientIn &pt, double sigma, int idxs, int lane,
                                    double *lg_row) {
  remove_patient(b, s, A, it, pt, lane);

  if (b.model.reuse == 0) {  // :605-610
    for (int mm = lane; mm < b.CC; mm += 32)
      rng_normals(rng, iter, TPPMX_SITE_AUX_PATIENT, tt, (uint32_t)(i * b.CC + mm), b.D, A.etae + mm, b.CC);
    __syncwarp();
  }

  double dV = 0.0;
  if (s.cohesion == 2 && A.K >= 1) {
    const size_t base = ((size_t)tt * 2) * b.ntmax;
    dV = b.logV[(base + b.ntmax + (A.K - 1)) * b.G + idxs] - b.logV[(base + (A.K - 1)) * b.G + idxs];
  }
  const ScoreOut so = score_clusters(b, s, A, pt, pt.xc, sigma, dV, !s.nolik, lane, nullptr, nullptr);

  double uu, u1;
  rng_u2(rng, iter, TPPMX_SITE_ALLOC_U, tt, (uint32_t)i, 0, uu, u1);
  const int newci = draw_cluster(A, so, uu, lane);

  const int ks = A.ks;
  int lab;
  if (newci <= A.K) {  // :820-828
    lab = newci;
    if (lane == 0) {
      A.labels[it] = lab;
      A.nj[lab - 1] += 1;
    }
  } else {  // :829-845
    if (A.K + 1 > ks) return false;
    const int id_empty = newci - A.K - 1;
    A.K += 1;
    lab = A.K;
    if (lane == 0) {
      A.labels[it] = lab;
      A.nj[lab - 1] = 1;
    }
    for (int k = lane; k < b.D; k += 32) A.eta[(size_t)k * ks + lab - 1] = A.etae[k * b.CC + id_empty];
    __syncwarp();
    if (lane == 0)
      rng_normals(rng, iter, TPPMX_SITE_AUX_REFRESH, tt, (uint32_t)i, b.D, A.etae + id_empty, b.CC);
  }
  for (int k = lane; k < b.D; k += 32) lg_row[k] = A.eta[(size_t)k * ks + lab - 1] + pt.zb[k];

  if (s.PPMx) {  // :847-856
    for (int p = lane; p < s.P; p += 32) {
      A.sumx[(size_t)p * ks + lab - 1] = __dadd_rn(A.sumx[(size_t)p * ks + lab - 1], pt.x[p]);
      A.sumx2[(size_t)p * ks + lab - 1] = __dadd_rn(A.sumx2[(size_t)p * ks + lab - 1], __dmul_rn(pt.x[p], pt.x[p]));
    }
    for (int p = lane; p < s.ncat; p += 32) A.njc[((size_t)p * b.maxC + pt.xc[p]) * ks + lab - 1] += 1;
  }
  __syncwarp();
  return true;
}

__device__ __forceinline__ PatientIn patient_in(const DevBatch &b, int chain, int ds, int i) {
  PatientIn pt;
  pt.x = b.x + ((size_t)ds * b.nobs + i) * b.P;
  pt.xc = b.xcat + ((size_t)ds * b.nobs + i) * (b.ncat > 0 ? b.ncat : 1);
  pt.zb = b.zb + ((size_t)chain * b.nobs + i) * b.D;
  pt.ljj = b.ljj + ((size_t)chain * b.nobs + i) * b.D;
  pt.cpat = b.cpat[(size_t)chain * b.nobs + i];
  return pt;
}

// per-patient quantities that are constant during a sweep: zb = beta' z, log JJ, and the
// cluster-independent part of the likelihood term.  All threads of the CTA.
__device__ inline void prep_patients(const DevBatch &b, int chain, int ds) {
  const double *beta = b.beta + (size_t)chain * b.Q * b.D;
  for (int i = threadIdx.x; i < b.nobs; i += blockDim.x) {
    const double *zi = b.z + ((size_t)ds * b.nobs + i) * b.Q;
    const double *JJ = b.JJ + ((size_t)chain * b.nobs + i) * b.D;
    const double *yi = b.y + ((size_t)ds * b.nobs + i) * b.D;
    const double ss1 = b.ss[(size_t)chain * b.nobs + i] + 1.0;
    double c = 0.0;
    for (int k = 0; k < b.D; k++) {
      double acc = 0.0;
      for (int q = 0; q < b.Q; q++) acc += beta[k + q * b.D] * zi[q];
      const double lj = log(JJ[k]);
      b.zb[((size_t)chain * b.nobs + i) * b.D + k] = acc;
      b.ljj[((size_t)chain * b.nobs + i) * b.D + k] = lj;
      c -= JJ[k] * ss1;
      if (yi[k] != 1.0) c -= lj;
    }
    b.cpat[(size_t)chain * b.nobs + i] = c;
  }
}

// the sweep of one arm for one iteration (:349-860 body for fixed tt), starting at patient it0
// (it0 > 0 only when a unit is taken over from the fast kernel mid-sweep)
__device__ inline bool sweep_arm(const DevBatch &b, const SimConst &s, ArmView &A, const Rng &rng,
                                 int chain, int ds, int tt, int iter, double sigma, int idxs,
                                 int lane, int it0) {
  if (b.model.reuse == 1 && it0 == 0) {  // :351-360
    for (int mm = lane; mm < b.CC; mm += 32)
      rng_normals(rng, iter, TPPMX_SITE_AUX_SWEEP, tt, (uint32_t)mm, b.D, A.etae + mm, b.CC);
    __syncwarp();
  }
  const int *pat = b.arm_pat + ((size_t)ds * b.T + tt) * b.ntmax;
  for (int it = it0; it < A.nt; it++) {
    const int i = pat[it];
    const PatientIn pt = patient_in(b, chain, ds, i);
    if (!realloc_step(b, s, A, rng, chain, tt, it, i, iter, pt, sigma, idxs, lane,
                      b.lg + ((size_t)chain * b.nobs + i) * b.D))
      return false;
  }
  return true;
}

// ---- kernels ----------------------------------------------------------------------------------
// grid = chains, block = 32*T threads: warp tt sweeps arm tt.  resume_l / resume_it (nullable):
// per (chain, arm) position at which the fast kernel handed the unit over; -1 = nothing to do.
__global__ void __launch_bounds__(512, 1) sweep_kernel(DevBatch b, uint64_t seed, int iter0,
                                                    int n_sweeps, const int *resume_l,
                                                    const int *resume_it) {
  const int chain = blockIdx.x;
  const int tt = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ds = b.chain_ds[chain];
  const SimConst s = make_simconst(b);
  const Rng rng = make_rng(seed, b.chain_id[chain]);
  if (resume_l) {  // nothing to do for this chain?
    bool any = false;
    for (int t = 0; t < b.T; t++) any |= resume_l[(size_t)chain * b.T + t] >= 0;
    if (!any) return;
  }
  prep_patients(b, chain, ds);
  __syncthreads();
  if (tt >= b.T) return;
  int l0 = iter0, it0 = 0;
  if (resume_l) {
    l0 = resume_l[(size_t)chain * b.T + tt];
    it0 = resume_it[(size_t)chain * b.T + tt];
    if (l0 < 0) return;
  }
  ArmView A = arm_view_global(b, chain, tt, ds);
  const double sigma = b.sigma[(size_t)chain * b.T + tt];
  const int idxs = b.idxs[chain];
  bool ok = true;
  for (int l = l0; l < iter0 + n_sweeps && ok; l++) {
    ok = sweep_arm(b, s, A, rng, chain, ds, tt, l, sigma, idxs, lane, it0);
    it0 = 0;
  }
  if (lane == 0) {
    b.nclu[(size_t)chain * b.T + tt] = A.K;
    if (!ok) atomicExch(&b.status[chain], TPPMX_ERR_CAPACITY);
  }
}

// parity hook: remove + score patient (arm, it) of the scratch chain slot, no draw
__global__ void score_kernel(DevBatch b, int chain, int tt, int it, uint64_t seed, int iter,
                             double *logw, double *prob, int *K_out) {
  const int lane = threadIdx.x & 31;
  const int ds = b.chain_ds[chain];
  const SimConst s = make_simconst(b);
  const Rng rng = make_rng(seed, b.chain_id[chain]);
  prep_patients(b, chain, ds);
  __syncthreads();
  ArmView A = arm_view_global(b, chain, tt, ds);
  const int i = b.arm_pat[((size_t)ds * b.T + tt) * b.ntmax + it];
  const PatientIn pt = patient_in(b, chain, ds, i);
  remove_patient(b, s, A, it, pt, lane);
  if (b.model.reuse == 0) {  // :605-610
    for (int mm = lane; mm < b.CC; mm += 32)
      rng_normals(rng, iter, TPPMX_SITE_AUX_PATIENT, tt, (uint32_t)(i * b.CC + mm), b.D, A.etae + mm, b.CC);
    __syncwarp();
  }
  double dV = 0.0;
  const int idxs = b.idxs[chain];
  if (s.cohesion == 2 && A.K >= 1) {
    const size_t base = ((size_t)tt * 2) * b.ntmax;
    dV = b.logV[(base + b.ntmax + (A.K - 1)) * b.G + idxs] - b.logV[(base + (A.K - 1)) * b.G + idxs];
  }
  score_clusters(b, s, A, pt, pt.xc, b.sigma[(size_t)chain * b.T + tt], dV, !s.nolik, lane, logw, prob);
  if (lane == 0) *K_out = A.K;
}

}  // namespace tppmx
