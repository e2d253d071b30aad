// update.cuh — everything of an MCMC iteration that follows the reallocation sweep, on device:
//   eta* Metropolis step per cluster     reference src/utils_ct.cpp:579-702 (src/dm_ppmx_ct.cpp:866-878)
//   (Theta, Sigma) hierarchy             src/dm_ppmx_ct.cpp:880-902
//   (sigma, kappa) grid draw             src/dm_ppmx_ct.cpp:907-1007
//   beta Metropolis step + horseshoe     src/utils_ct.cpp:710-800, 876-915 (src/dm_ppmx_ct.cpp:1009-1034)
//   JJ / ss gamma draws                  src/dm_ppmx_ct.cpp:1042-1055
// with these on the device, batched chains never return to the host.
// One CTA per chain; phases separated by __syncthreads.  The reference's sequential loops over
// clusters / patients carry no dependence inside a phase (clusters are disjoint; the beta steps
// are sequential over (k, q) and stay so).  Same Philox sites and draw order as the CPU oracle
// (oracle/ppmx_chain.c), which restates the reference statement by statement.
#pragma once
#include "fastmath.cuh"
#include "incgamma.h"
#include "layout.h"
#include "rng.cuh"
#include "sim.cuh"

namespace tppmx {

struct UpdArgs {
  double *eta2;  // [nc][T][2][D][kcap]  old / proposed eta*
  double *pat2;  // [nc][2][T][ntmax]    per-patient current / proposed likelihood sums
  double *os;    // [nc][T][G]           (sigma,kappa) grid weights
  int *acc;      // [nc][T][kcap]        accept flags of this iteration
};

// lower Cholesky factor, column-major n x n
__device__ inline void d_chol_lower(const double *A, int n, double *L) {
  for (int i = 0; i < n * n; i++) L[i] = 0.0;
  for (int i = 0; i < n; i++)
    for (int j = 0; j <= i; j++) {
      double sum = A[i + n * j];
      for (int k = 0; k < j; k++) sum -= L[i + n * k] * L[j + n * k];
      L[i + n * j] = (i == j) ? sqrt(sum) : sum / L[j + n * j];
    }
}
__device__ inline void d_inv_spd(const double *A, int n, double *Ainv) {
  double L[TPPMX_MAXD * TPPMX_MAXD], Li[TPPMX_MAXD * TPPMX_MAXD];
  d_chol_lower(A, n, L);
  for (int i = 0; i < n * n; i++) Li[i] = 0.0;
  for (int c = 0; c < n; c++)
    for (int r = 0; r < n; r++) {
      double s = (r == c) ? 1.0 : 0.0;
      for (int k = 0; k < r; k++) s -= L[r + n * k] * Li[k + n * c];
      Li[r + n * c] = s / L[r + n * r];
    }
  for (int r = 0; r < n; r++)
    for (int c = 0; c < n; c++) {
      double s = 0.0;
      for (int k = 0; k < n; k++) s += Li[k + n * r] * Li[k + n * c];
      Ainv[r + n * c] = s;
    }
}
// dmvnrm(x; mean, sigma) on the log scale given L = chol_lower(sigma)  (utils_ct.cpp:242-266)
__device__ inline double d_dmvn_log(const double *x, int xstride, const double *mean, const double *L, int n) {
  double z[TPPMX_MAXD], rootisum = 0.0, zz = 0.0;
  for (int r = 0; r < n; r++) {
    double s = x[(size_t)r * xstride] - mean[r];
    for (int k = 0; k < r; k++) s -= L[r + n * k] * z[k];
    z[r] = s / L[r + n * r];
    zz += z[r] * z[r];
    rootisum += log(1.0 / L[r + n * r]);
  }
  return -(double)n / 2.0 * 1.8378770664093453 - 0.5 * zz + rootisum;
}

// sum over a block of (a, b) pairs, deterministic tree; result valid in every thread
__device__ inline void block_sum2(double &a, double &b, double *red) {
  const int tid = threadIdx.x, nth = blockDim.x;
  red[tid] = a;
  red[nth + tid] = b;
  __syncthreads();
  for (int o = nth >> 1; o > 0; o >>= 1) {
    if (tid < o) {
      red[tid] += red[tid + o];
      red[nth + tid] += red[nth + tid + o];
    }
    __syncthreads();
  }
  a = red[0];
  b = red[nth];
  __syncthreads();
}

// blockDim.x must be a power of two; dynamic smem: 2 * blockDim doubles + T*D*D + 64
__global__ void __launch_bounds__(256) update_kernel(DevBatch b, UpdArgs ua, uint64_t seed, int l) {
  extern __shared__ double sm[];
  const int chain = blockIdx.x, tid = threadIdx.x, nth = blockDim.x;
  const int ds = b.chain_ds[chain];
  const Rng rng = make_rng(seed, b.chain_id[chain]);
  const int nobs = b.nobs, T = b.T, D = b.D, Q = b.Q, P = b.P, ks = b.kcap, G = b.G, ntm = b.ntmax;
  double *red = sm;                       // [2 * nth]
  double *Lsig = sm + 2 * nth;            // [T][D*D]
  double *bc = Lsig + T * D * D;          // broadcast slots [16]
  int *pick = reinterpret_cast<int *>(bc + 16);  // [T]

  const int *labels = b.labels + (size_t)chain * st_labels(b);
  const int *njv = b.nj + (size_t)chain * st_nj(b);
  const int *nclu = b.nclu + (size_t)chain * T;
  double *eta = b.eta + (size_t)chain * st_eta(b);
  double *lg = b.lg + (size_t)chain * st_pat(b);
  double *JJ = b.JJ + (size_t)chain * st_pat(b);
  double *ss = b.ss + (size_t)chain * nobs;
  double *beta = b.beta + (size_t)chain * Q * D;
  double *Theta = b.Theta + (size_t)chain * T * D;
  double *Sigma = b.Sigma + (size_t)chain * T * D * D;
  double *eold = ua.eta2 + (size_t)chain * T * 2 * D * ks;
  double *cur = ua.pat2 + (size_t)chain * 2 * T * ntm;
  double *prop = cur + (size_t)T * ntm;
  int *acc = ua.acc + (size_t)chain * T * ks;
  const int *arm_n = b.arm_n + (size_t)ds * T;
  const int *arm_pat = b.arm_pat + (size_t)ds * T * ntm;
  const double *z = b.z + (size_t)ds * nobs * Q;
  const double *y = b.y + (size_t)ds * nobs * D;
  const double mh_eta = b.model.mhtune_eta, mh_beta = b.model.mhtune_beta;

  // ================= eta* Metropolis step, utils_ct.cpp:579-702 =================
  for (int t = tid; t < T; t += nth) d_chol_lower(Sigma + t * D * D, D, Lsig + t * D * D);
  for (int e = tid; e < T * ks; e += nth) {
    const int t = e / ks, j = e - t * ks;
    if (j < nclu[t]) {
      double nrm[TPPMX_MAXD];
      rng_normals(rng, l, TPPMX_SITE_ETA_PROP, t, (uint32_t)j, D, nrm, 1);
      for (int k = 0; k < D; k++) {
        const double ev = eta[((size_t)t * D + k) * ks + j];
        eold[(((size_t)t * 2 + 0) * D + k) * ks + j] = ev;
        eold[(((size_t)t * 2 + 1) * D + k) * ks + j] = ev + mh_eta * nrm[k];  // :643
      }
    }
  }
  __syncthreads();
  for (int e = tid; e < T * ntm; e += nth) {
    const int t = e / ntm, it = e - t * ntm;
    if (it < arm_n[t]) {
      const int i = arm_pat[e], j = labels[e] - 1;
      double c = 0.0, p = 0.0;
      for (int k = 0; k < D; k++) {
        const double lgv = lg[(size_t)i * D + k], lj = fast_log(JJ[(size_t)i * D + k]);
        const double e0 = eold[(((size_t)t * 2 + 0) * D + k) * ks + j], e1 = eold[(((size_t)t * 2 + 1) * D + k) * ks + j];
        const double g0 = fast_exp(lgv), g1 = fast_exp(__dadd_rn(__dsub_rn(lgv, e0), e1));  // :651
        c += lgamma_pos(g0) + g0 * lj;
        p += lgamma_pos(g1) + g1 * lj;
      }
      cur[e] = c;
      prop[e] = p;
    }
  }
  __syncthreads();
  for (int e = tid; e < T * ks; e += nth) {
    const int t = e / ks, j = e - t * ks;
    if (j < nclu[t]) {
      double log_den = 0.0, log_num = 0.0;
      const int nt = arm_n[t];
      for (int it = 0; it < nt; it++)
        if (labels[t * ntm + it] == j + 1) {
          log_den -= cur[t * ntm + it];
          log_num -= prop[t * ntm + it];
        }
      const double *L = Lsig + t * D * D;
      log_den += d_dmvn_log(eold + (((size_t)t * 2 + 0) * D) * ks + j, ks, Theta + t * D, L, D);
      log_den += d_dmvn_log(eold + (((size_t)t * 2 + 1) * D) * ks + j, ks, Theta + t * D, L, D);  // quirk 9 (:671)
      double u0, u1;
      rng_u2(rng, l, TPPMX_SITE_ETA_PROP, t, (uint32_t)j, 0x100u, u0, u1);
      const int ok = log(u0) < log_num - log_den;
      acc[e] = ok;
      if (ok) {
        b.eta_flag[(size_t)chain * T * ks + e] += 1;
        for (int k = 0; k < D; k++) eta[((size_t)t * D + k) * ks + j] = eold[(((size_t)t * 2 + 1) * D + k) * ks + j];
      }
    }
  }
  __syncthreads();
  for (int e = tid; e < T * ntm; e += nth) {
    const int t = e / ntm, it = e - t * ntm;
    if (it < arm_n[t]) {
      const int i = arm_pat[e], j = labels[e] - 1;
      if (acc[t * ks + j])
        for (int k = 0; k < D; k++)
          lg[(size_t)i * D + k] = __dadd_rn(__dsub_rn(lg[(size_t)i * D + k], eold[(((size_t)t * 2 + 0) * D + k) * ks + j]),
                                            eold[(((size_t)t * 2 + 1) * D + k) * ks + j]);
    }
  }
  __syncthreads();

  // ================= (Theta, Sigma) hierarchy, src/dm_ppmx_ct.cpp:880-902 =================
  if (b.model.upd_hier == 1) {
    for (int t = tid; t < T; t += nth) {
      const int nt = arm_n[t];
      const double dn = (double)nt, nu0 = b.model.hP0_nu0;
      double em[TPPMX_MAXD], ecov[TPPMX_MAXD * TPPMX_MAXD], S[TPPMX_MAXD * TPPMX_MAXD];
      for (int k = 0; k < D; k++) {
        double a = 0.0;
        for (int i = 0; i < nt; i++) a += eta[((size_t)t * D + k) * ks + labels[t * ntm + i] - 1];
        em[k] = a / dn;
      }
      for (int k = 0; k < D; k++)
        for (int r = 0; r < D; r++) {
          double a = 0.0;
          for (int i = 0; i < nt; i++) {
            const int j = labels[t * ntm + i] - 1;
            a += (eta[((size_t)t * D + k) * ks + j] - em[k]) * (eta[((size_t)t * D + r) * ks + j] - em[r]);
          }
          ecov[k + D * r] = a / dn;
        }
      const double fplp = b.model.hP0_s0 + (dn * .5);
      for (int k = 0; k < D; k++)
        for (int r = 0; r < D; r++)
          S[k + D * r] = ((k == r) ? b.model.hP0_Lambda0 : 0.0) +
                         (dn * .5) * (ecov[k + D * r] + (nu0 / (nu0 + dn)) * ((em[k] - b.model.hP0_mu0[k]) * (em[r] - b.model.hP0_mu0[r])));
      // Lambda ~ Wishart(S, fplp): Bartlett
      double L[TPPMX_MAXD * TPPMX_MAXD], A[TPPMX_MAXD * TPPMX_MAXD], LA[TPPMX_MAXD * TPPMX_MAXD], Lam[TPPMX_MAXD * TPPMX_MAXD];
      d_chol_lower(S, D, L);
      for (int i = 0; i < D * D; i++) A[i] = 0.0;
      for (int k = 0; k < D; k++) {
        A[k + D * k] = sqrt(2.0 * rng_gamma(rng, l, TPPMX_SITE_HIER_W, t, (uint32_t)k, 0.5 * (fplp - (double)k)));
        for (int r = 0; r < k; r++) {
          double n0, n1;
          rng_n2(rng, l, TPPMX_SITE_HIER_W, t, (uint32_t)(1000 + k * D + r), 0, n0, n1);
          A[k + D * r] = n0;
        }
      }
      for (int k = 0; k < D; k++)
        for (int r = 0; r < D; r++) {
          double a = 0.0;
          for (int i = 0; i < D; i++) a += L[k + D * i] * A[i + D * r];
          LA[k + D * r] = a;
        }
      for (int k = 0; k < D; k++)
        for (int r = 0; r < D; r++) {
          double a = 0.0;
          for (int i = 0; i < D; i++) a += LA[k + D * i] * LA[r + D * i];
          Lam[k + D * r] = a;
        }
      double fptp[TPPMX_MAXD], tmp[TPPMX_MAXD * TPPMX_MAXD], sptp[TPPMX_MAXD * TPPMX_MAXD], zn[TPPMX_MAXD];
      for (int k = 0; k < D; k++) fptp[k] = (b.model.hP0_mu0[k] * nu0 + dn * em[k]) / (nu0 + dn);
      for (int k = 0; k < D * D; k++) tmp[k] = Lam[k] * (dn + nu0);
      d_inv_spd(tmp, D, sptp);
      d_chol_lower(sptp, D, L);
      rng_normals(rng, l, TPPMX_SITE_HIER_THETA, t, 0, D, zn, 1);
      for (int k = 0; k < D; k++) {
        double a = fptp[k];
        for (int r = 0; r <= k; r++) a += L[k + D * r] * zn[r];
        Theta[t * D + k] = a;
      }
      d_inv_spd(Lam, D, Sigma + t * D * D);
    }
    __syncthreads();
  }

  // ================= (sigma, kappa) grid draw, src/dm_ppmx_ct.cpp:907-1007 =================
  // omegasigma (cluster similarities) is the same for every grid point and cancels in the
  // normalisation, so it is not evaluated (oracle keeps it; the draw is the same law).
  if (b.model.cohesion == 2) {
    double *os = ua.os + (size_t)chain * T * G;
    for (int e = tid; e < T * G; e += nth) {
      const int t = e / G, g = e - t * G;
      const double sg = b.grid[2 * g];
      const int K = nclu[t];
      double v = 0.0;
      const double l1 = lgamma(1.0 - sg);
      for (int j = 0; j < K; j++) v += lgamma((double)njv[t * ks + j] - sg) - l1;
      v += b.logV[(((size_t)t * 2 + 1) * ntm + (K - 1)) * G + g];
      os[e] = v;
    }
    __syncthreads();
    for (int t = tid; t < T; t += nth) {
      double mx = os[t * G];
      for (int g = 1; g < G; g++) mx = fmax(mx, os[t * G + g]);
      double den = 0.0;
      for (int g = 0; g < G; g++) {
        const double e = exp(os[t * G + g] - mx);
        os[t * G + g] = e;
        den += e;
      }
      double u0, u1, cw = 0.0;
      rng_u2(rng, l, TPPMX_SITE_SIGKAP, t, 0, 0, u0, u1);
      int pk = G - 1;
      for (int g = 0; g < G; g++) {
        cw += os[t * G + g] / den;
        if (u0 < cw) {
          pk = g;
          break;
        }
      }
      pick[t] = pk;
      b.sigma[(size_t)chain * T + t] = b.grid[2 * pk];
      b.kappa[(size_t)chain * T + t] = b.grid[2 * pk + 1];
    }
    __syncthreads();
    if (tid == 0) b.idxs[chain] = pick[T - 1];  // quirk 2: one shared index, the last arm's draw
    __syncthreads();
  }

  // ================= beta Metropolis steps + horseshoe =================
  if (b.model.noprog != 1) {
    double *sigma_beta = b.sigma_beta + (size_t)chain * Q * D;
    for (int kk = 0; kk < D; kk++)
      for (int q = 0; q < Q; q++) {  // utils_ct.cpp:750-791
        const int h = kk + q * D;
        const double bh = beta[h];
        double n0, n1;
        rng_n2(rng, l, TPPMX_SITE_BETA, 0, (uint32_t)(kk * Q + q), 0, n0, n1);
        const double beta_p = bh + mh_beta * n0;
        double sd = 0.0, sn = 0.0;
        for (int i = tid; i < nobs; i += nth) {
          const double lgv = lg[(size_t)i * D + kk], lj = fast_log(JJ[(size_t)i * D + kk]);
          const double zq = z[(size_t)i * Q + q];
          const double lgp = __dadd_rn(__dsub_rn(lgv, __dmul_rn(bh, zq)), __dmul_rn(beta_p, zq));  // :764
          const double g0 = fast_exp(lgv), g1 = fast_exp(lgp);
          sd += g0 * lj - lgamma_pos(g0);
          sn += g1 * lj - lgamma_pos(g1);
        }
        block_sum2(sd, sn, red);
        const double sb = sigma_beta[h];
        const double log_den = sd + dnorm_log(bh, 0.0, sb), log_num = sn + dnorm_log(beta_p, 0.0, sb);
        double u0, u1;
        rng_u2(rng, l, TPPMX_SITE_BETA, 0, (uint32_t)(kk * Q + q), 1, u0, u1);
        if (log(u0) < log_num - log_den) {  // uniform over the block: same inputs in every thread
          for (int i = tid; i < nobs; i += nth) {
            const double zq = z[(size_t)i * Q + q];
            lg[(size_t)i * D + kk] = __dadd_rn(__dsub_rn(lg[(size_t)i * D + kk], __dmul_rn(bh, zq)), __dmul_rn(beta_p, zq));
          }
          if (tid == 0) {
            beta[h] = beta_p;
            b.beta_flag[(size_t)chain * Q * D + q + Q * kk] += 1;
          }
        }
        __syncthreads();
      }
    if (b.model.hsp == 1) {  // utils_ct.cpp:876-915
      if (tid == 0) {
        const int len = Q * D;
        double *lambda = b.lambda + (size_t)chain * len;
        double tau = b.tau[chain];
        for (int k = 0; k < len; k++) {
          double u0, u1;
          rng_u2(rng, l, TPPMX_SITE_HS_LAMBDA, 0, (uint32_t)k, 0, u0, u1);
          double et = 1.0 / (lambda[k] * lambda[k]);
          const double upsi = u0 * (1.0 / (1.0 + et));
          const double tempps = (beta[k] * beta[k]) / (2.0 * tau * tau);
          const double ub = (1.0 - upsi) / upsi;
          double Fub = 1.0 - exp(-tempps * ub);
          if (Fub < 1e-4) Fub = 1e-4;
          const double up = u1 * Fub;
          et = -log(1.0 - up) / tempps;
          lambda[k] = 1.0 / sqrt(et);
        }
        double tempt = 0.0, u0, u1;
        for (int k = 0; k < len; k++) {
          const double r = beta[k] / lambda[k];
          tempt += r * r;
        }
        tempt = tempt / 2.0;
        double et = 1.0 / (tau * tau);
        rng_u2(rng, l, TPPMX_SITE_HS_TAU, 0, 0, 0, u0, u1);
        const double utau = u0 * (1.0 / (1.0 + et));
        const double ubt = (1.0 - utau) / utau;
        const double shape = ((double)len + 1.0) / 2.0;
        double Fubt = tppmx_gammainc(shape, ubt * tempt);
        if (Fubt < 1e-4) Fubt = 1e-4;
        const double ut = u1 * Fubt;
        et = tppmx_gammaincinv(shape, ut) / tempt;
        if (et < .0001) et = .0001;
        tau = 1.0 / sqrt(et);
        if (tau < .0001) tau = .0001;
        b.tau[chain] = tau;
        for (int k = 0; k < len; k++) sigma_beta[k] = lambda[k] * tau;
      }
      __syncthreads();
    }
  }

  // ================= JJ, ss gamma draws, src/dm_ppmx_ct.cpp:1042-1055 =================
  for (int e = tid; e < nobs * D; e += nth) {
    const int i = e / D;
    const double shape = y[e] + exp(lg[e]);
    double v = rng_gamma(rng, l, TPPMX_SITE_JJ, 0, (uint32_t)e, shape) * (1.0 / (ss[i] + 1.0));
    if (v < 1e-100) v = 1e-100;
    JJ[e] = v;
  }
  __syncthreads();
  for (int i = tid; i < nobs; i += nth) {
    double TT = 0.0;
    for (int k = 0; k < D; k++) TT += JJ[(size_t)i * D + k];
    ss[i] = rng_gamma(rng, l, TPPMX_SITE_SS, 0, (uint32_t)i, 1.0) * (1.0 / TT);
  }
  (void)P;
}

}  // namespace tppmx
