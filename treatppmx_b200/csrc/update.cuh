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
  double *pat2;  // [nc][3][T][ntmax]    per-patient current / proposed likelihood sums (+ log JJ)
  double *os;    // [nc][T][G]           (sigma,kappa) grid weights
  int *acc;      // [nc][T][kcap]        accept flags of this iteration
  const double *lgtab;  // [G][ntmax+1]    lgamma(n - sigma_g) - lgamma(1 - sigma_g) (cohesion 2)
};

// lower Cholesky factor, column-major n x n (n <= DM).  The loops are deliberately NOT unrolled:
// this code runs once per chain and iteration on one lane, so its cost is instruction fetch,
// not arithmetic (profiles/r1_update_kernel_*: stall_no_instruction dominates).
template <int DM>
__device__ __forceinline__ void d_chol_lower(const double *A, int n, double *L) {
  #pragma unroll 1
  for (int i = 0; i < n * n; i++) L[i] = 0.0;
  #pragma unroll 1
  for (int i = 0; i < n; i++)
    #pragma unroll 1
    for (int j = 0; j <= i; j++) {
      double sum = A[i + n * j];
      #pragma unroll 1
      for (int k = 0; k < j; k++) sum -= L[i + n * k] * L[j + n * k];
      L[i + n * j] = (i == j) ? sqrt(sum) : sum / L[j + n * j];
    }
}
template <int DM>
__device__ __forceinline__ void d_inv_spd(const double *A, int n, double *Ainv) {
  double L[DM * DM], Li[DM * DM];
  d_chol_lower<DM>(A, n, L);
  #pragma unroll 1
  for (int i = 0; i < n * n; i++) Li[i] = 0.0;
  #pragma unroll 1
  for (int c = 0; c < n; c++)
    #pragma unroll 1
    for (int r = 0; r < n; r++) {
      double s = (r == c) ? 1.0 : 0.0;
      #pragma unroll 1
      for (int k = 0; k < r; k++) s -= L[r + n * k] * Li[k + n * c];
      Li[r + n * c] = s / L[r + n * r];
    }
  #pragma unroll 1
  for (int r = 0; r < n; r++)
    #pragma unroll 1
    for (int c = 0; c < n; c++) {
      double s = 0.0;
      #pragma unroll 1
      for (int k = 0; k < n; k++) s += Li[k + n * r] * Li[k + n * c];
      Ainv[r + n * c] = s;
    }
}
// dmvnrm(x; mean, sigma) on the log scale given L = chol_lower(sigma)  (utils_ct.cpp:242-266)
template <int DM>
__device__ __forceinline__ double d_dmvn_log(const double *x, int xstride, const double *mean, const double *L, int n) {
  double z[DM], rootisum = 0.0, zz = 0.0;
  #pragma unroll 1
  for (int r = 0; r < n; r++) {
    double s = x[(size_t)r * xstride] - mean[r];
    #pragma unroll 1
    for (int k = 0; k < r; k++) s -= L[r + n * k] * z[k];
    z[r] = s / L[r + n * r];
    zz += z[r] * z[r];
    rootisum += fast_log(1.0 / L[r + n * r]);
  }
  return -(double)n / 2.0 * 1.8378770664093453 - 0.5 * zz + rootisum;
}

// Out-of-line copies of the heavy helpers: the kernel has ~15 call sites of Philox / gamma /
// lgamma; inlined, the code grows to ~270 KB and the instruction cache, not the math, bounds it
// (profiles/r1_update_kernel_*: stall_no_instruction 3.8 per issue).
__device__ __noinline__ double upd_lgamma(double x) { return lgamma_fast(x); }
__device__ __noinline__ double upd_gamma(Rng g, uint32_t iter, uint32_t site, uint32_t arm, uint32_t index, double a) {
  return rng_gamma(g, iter, site, arm, index, a);
}
__device__ __noinline__ double2 upd_u2(Rng g, uint32_t iter, uint32_t site, uint32_t arm, uint32_t index, uint32_t sub) {
  double u0, u1;
  rng_u2(g, iter, site, arm, index, sub, u0, u1);
  return make_double2(u0, u1);
}
__device__ __noinline__ double2 upd_n2(Rng g, uint32_t iter, uint32_t site, uint32_t arm, uint32_t index, uint32_t sub) {
  double n0, n1;
  rng_n2(g, iter, site, arm, index, sub, n0, n1);
  return make_double2(n0, n1);
}
__device__ __noinline__ double upd_gammainc(double a, double x) { return tppmx_gammainc(a, x); }
__device__ __noinline__ double upd_gammaincinv(double a, double p) { return tppmx_gammaincinv(a, p); }
__device__ __noinline__ double upd_log(double x) { return fast_log(x); }
__device__ __noinline__ double upd_exp(double x) { return fast_exp(x); }
// D standard normals (block b gives normals 2b, 2b+1), as rng_normals
__device__ __forceinline__ void upd_normals(const Rng &g, uint32_t iter, uint32_t site, uint32_t arm, uint32_t index,
                                            int D, double *out) {
  for (int bq = 0; 2 * bq < D; bq++) {
    const double2 n = upd_n2(g, iter, site, arm, index, (uint32_t)bq);
    out[2 * bq] = n.x;
    if (2 * bq + 1 < D) out[2 * bq + 1] = n.y;
  }
}

// named barriers: 1 = the worker warps among themselves, 2 = workers -> serial warp hand-off
__device__ __forceinline__ void bar_workers(int nw) { asm volatile("bar.sync 1, %0;" ::"r"(nw) : "memory"); }

// sum of (a, b) over the nw worker threads (warps 0 .. nw/32-1), deterministic order; the result
// is valid in every worker thread.  red: 2 * (nw/32) doubles of shared memory.
__device__ inline void workers_sum2(double &a, double &b, double *red, int nw) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  const int w = threadIdx.x >> 5, nwarp = nw >> 5;
  if ((threadIdx.x & 31) == 0) {
    red[2 * w] = a;
    red[2 * w + 1] = b;
  }
  bar_workers(nw);
  a = 0.0;
  b = 0.0;
  for (int i = 0; i < nwarp; i++) {
    a += red[2 * i];
    b += red[2 * i + 1];
  }
  bar_workers(nw);
}

// blockDim.x = 128: warps 0-2 are workers, warp 3 is the SERIAL warp that runs the two phases
// with no parallelism to speak of ((Theta,Sigma) hierarchy: one lane per arm; horseshoe: one
// lane) concurrently with the workers' grid draw, beta steps and gamma draws.
// DD = 3 (the package's outcome dimension: loops unroll, the small matrices live in registers)
// or 0 (generic, D <= TPPMX_MAXD).  dynamic smem: (64 + T*D*D + 16) doubles + T ints
template <int DD>
__global__ void __launch_bounds__(128, 7) update_kernel(DevBatch b, UpdArgs ua, uint64_t seed, int l) {
  extern __shared__ double sm[];
  const int chain = blockIdx.x, tid = threadIdx.x, nth = blockDim.x;
  const int nw = nth - 32;           // worker threads
  const bool serial = tid >= nw;     // the serial warp
  const int ds = b.chain_ds[chain];
  const Rng rng = make_rng(seed, b.chain_id[chain]);
  const int nobs = b.nobs, T = b.T, D = DD ? DD : b.D, Q = b.Q, P = b.P, ks = b.kcap, G = b.G, ntm = b.ntmax;
  constexpr int DM = DD ? DD : TPPMX_MAXD;  // extent of the small local arrays
  double *red = sm;                       // [64]
  double *Lsig = sm + 64;                 // [T][D*D]

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
  double *cur = ua.pat2 + (size_t)chain * 3 * T * ntm;
  double *prop = cur + (size_t)T * ntm;
  int *acc = ua.acc + (size_t)chain * T * ks;
  const int *arm_n = b.arm_n + (size_t)ds * T;
  const int *arm_pat = b.arm_pat + (size_t)ds * T * ntm;
  const double *z = b.z + (size_t)ds * nobs * Q;
  const double *y = b.y + (size_t)ds * nobs * D;
  const double mh_eta = b.model.mhtune_eta, mh_beta = b.model.mhtune_beta;

  // ================= eta* Metropolis step, utils_ct.cpp:579-702 =================
  for (int t = tid; t < T; t += nth) d_chol_lower<DM>(Sigma + t * D * D, D, Lsig + t * D * D);
  for (int e = tid; e < T * ks; e += nth) {
    const int t = e / ks, j = e - t * ks;
    if (j < nclu[t]) {
      double nrm[DM];
      upd_normals(rng, l, TPPMX_SITE_ETA_PROP, t, (uint32_t)j, D, nrm);
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
        const double lgv = lg[(size_t)i * D + k], lj = upd_log(JJ[(size_t)i * D + k]);
        const double e0 = eold[(((size_t)t * 2 + 0) * D + k) * ks + j], e1 = eold[(((size_t)t * 2 + 1) * D + k) * ks + j];
        const double g0 = upd_exp(lgv), g1 = upd_exp(__dadd_rn(__dsub_rn(lgv, e0), e1));  // :651
        c += upd_lgamma(g0) + g0 * lj;
        p += upd_lgamma(g1) + g1 * lj;
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
      log_den += d_dmvn_log<DM>(eold + (((size_t)t * 2 + 0) * D) * ks + j, ks, Theta + t * D, L, D);
      log_den += d_dmvn_log<DM>(eold + (((size_t)t * 2 + 1) * D) * ks + j, ks, Theta + t * D, L, D);  // quirk 9 (:671)
      double u0, u1;
      { const double2 r2_ = upd_u2(rng, l, TPPMX_SITE_ETA_PROP, t, (uint32_t)j, 0x100u); u0 = r2_.x; u1 = r2_.y; }
      const int ok = upd_log(u0) < log_num - log_den;
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

  // From here the block splits: the SERIAL warp runs the hierarchy (one lane per arm) and, once
  // the workers have finished the beta steps, the horseshoe update (one lane); the WORKERS run
  // the (sigma,kappa) grid draw, the beta Metropolis steps and the JJ / ss gamma draws.  The two
  // sides touch disjoint state (Theta,Sigma | lambda,tau,sigma_beta  vs  sigma,kappa,idxs,
  // beta,loggamma,JJ,ss); sigma_beta is read by the beta steps BEFORE the horseshoe rewrites it,
  // as in the reference (beta_update :1017 precedes up_lambda_hs :1030).
  if (serial) {
    const int lane = tid - nw;
    // ================= (Theta, Sigma) hierarchy, src/dm_ppmx_ct.cpp:880-902 =================
    if (b.model.upd_hier == 1) {
      #pragma unroll 1
      for (int t = lane; t < T; t += 32) {
        const int nt = arm_n[t];
        const double dn = (double)nt, nu0 = b.model.hP0_nu0;
        double em[DM], ecov[DM * DM], S[DM * DM];
        // eta_i = eta*[label_i] (:885): mean and covariance over patients are sums over the K
        // clusters weighted by their sizes (same value, K instead of n_t terms)
        const int K = nclu[t];
        #pragma unroll 1
        for (int k = 0; k < D; k++) {
          double a = 0.0;
          #pragma unroll 1
          for (int j = 0; j < K; j++) a += (double)njv[t * ks + j] * eta[((size_t)t * D + k) * ks + j];
          em[k] = a / dn;
        }
        #pragma unroll 1
        for (int k = 0; k < D; k++)
          #pragma unroll 1
          for (int r = 0; r < D; r++) {
            double a = 0.0;
            #pragma unroll 1
            for (int j = 0; j < K; j++)
              a += (double)njv[t * ks + j] * ((eta[((size_t)t * D + k) * ks + j] - em[k]) * (eta[((size_t)t * D + r) * ks + j] - em[r]));
            ecov[k + D * r] = a / dn;
          }
        const double fplp = b.model.hP0_s0 + (dn * .5);
        #pragma unroll 1
        for (int k = 0; k < D; k++)
          #pragma unroll 1
          for (int r = 0; r < D; r++)
            S[k + D * r] = ((k == r) ? b.model.hP0_Lambda0 : 0.0) +
                           (dn * .5) * (ecov[k + D * r] + (nu0 / (nu0 + dn)) * ((em[k] - b.model.hP0_mu0[k]) * (em[r] - b.model.hP0_mu0[r])));
        // Lambda ~ Wishart(S, fplp): Bartlett
        double L[DM * DM], A[DM * DM], LA[DM * DM], Lam[DM * DM];
        d_chol_lower<DM>(S, D, L);
        #pragma unroll 1
        for (int i = 0; i < D * D; i++) A[i] = 0.0;
        #pragma unroll 1
        for (int k = 0; k < D; k++) {
          A[k + D * k] = sqrt(2.0 * upd_gamma(rng, l, TPPMX_SITE_HIER_W, t, (uint32_t)k, 0.5 * (fplp - (double)k)));
          #pragma unroll 1
          for (int r = 0; r < k; r++) {
            double n0, n1;
            { const double2 r2_ = upd_n2(rng, l, TPPMX_SITE_HIER_W, t, (uint32_t)(1000 + k * D + r), 0); n0 = r2_.x; n1 = r2_.y; }
            A[k + D * r] = n0;
          }
        }
        #pragma unroll 1
        for (int k = 0; k < D; k++)
          #pragma unroll 1
          for (int r = 0; r < D; r++) {
            double a = 0.0;
            #pragma unroll 1
            for (int i = 0; i < D; i++) a += L[k + D * i] * A[i + D * r];
            LA[k + D * r] = a;
          }
        #pragma unroll 1
        for (int k = 0; k < D; k++)
          #pragma unroll 1
          for (int r = 0; r < D; r++) {
            double a = 0.0;
            #pragma unroll 1
            for (int i = 0; i < D; i++) a += LA[k + D * i] * LA[r + D * i];
            Lam[k + D * r] = a;
          }
        double fptp[DM], tmp[DM * DM], sptp[DM * DM], zn[DM];
        #pragma unroll 1
        for (int k = 0; k < D; k++) fptp[k] = (b.model.hP0_mu0[k] * nu0 + dn * em[k]) / (nu0 + dn);
        #pragma unroll 1
        for (int k = 0; k < D * D; k++) tmp[k] = Lam[k] * (dn + nu0);
        d_inv_spd<DM>(tmp, D, sptp);
        d_chol_lower<DM>(sptp, D, L);
        upd_normals(rng, l, TPPMX_SITE_HIER_THETA, t, 0, D, zn);
        #pragma unroll 1
        for (int k = 0; k < D; k++) {
          double a = fptp[k];
          #pragma unroll 1
          for (int r = 0; r <= k; r++) a += L[k + D * r] * zn[r];
          Theta[t * D + k] = a;
        }
        d_inv_spd<DM>(Lam, D, tmp);
        #pragma unroll 1
        for (int k = 0; k < D * D; k++) Sigma[t * D * D + k] = tmp[k];
      }
    }
    return;
  }

  // ------------------------------------ WORKER warps ------------------------------------
  // ================= (sigma, kappa) grid draw, src/dm_ppmx_ct.cpp:907-1007 =================
  // omegasigma (cluster similarities) is the same for every grid point and cancels in the
  // normalisation, so it is not evaluated (oracle keeps it; the draw is the same law).
  if (b.model.cohesion == 2) {
    double *os = ua.os + (size_t)chain * T * G;
    for (int e = tid; e < T * G; e += nw) {
      const int t = e / G, g = e - t * G;
      const int K = nclu[t];
      double v = 0.0;
      const double *lt = ua.lgtab + (size_t)g * (ntm + 1);  // lgamma(n - sigma_g) - lgamma(1 - sigma_g)
      for (int j = 0; j < K; j++) v += lt[njv[t * ks + j]];
      v += b.logV[(((size_t)t * 2 + 1) * ntm + (K - 1)) * G + g];
      os[e] = v;
    }
    bar_workers(nw);
    double *mxs = red + 16;  // [T] maxima (red[0..15] is the reduction scratch)
    for (int t = tid; t < T; t += nw) {
      double mx = os[t * G];
      for (int g = 1; g < G; g++) mx = fmax(mx, os[t * G + g]);
      mxs[t] = mx;
    }
    bar_workers(nw);
    for (int e = tid; e < T * G; e += nw) os[e] = upd_exp(os[e] - mxs[e / G]);
    bar_workers(nw);
    for (int t = tid; t < T; t += nw) {
      double den = 0.0;
      for (int g = 0; g < G; g++) den += os[t * G + g];
      double u0, u1, cw = 0.0;
      { const double2 r2_ = upd_u2(rng, l, TPPMX_SITE_SIGKAP, t, 0, 0); u0 = r2_.x; u1 = r2_.y; }
      (void)u1;
      int pk = G - 1;
      const double thr = u0 * den;  // u0 < cumsum(e)/den  <=>  u0*den < cumsum(e)
      for (int g = 0; g < G; g++) {
        cw += os[t * G + g];
        if (thr < cw) {
          pk = g;
          break;
        }
      }
      b.sigma[(size_t)chain * T + t] = b.grid[2 * pk];
      b.kappa[(size_t)chain * T + t] = b.grid[2 * pk + 1];
      if (t == T - 1) b.idxs[chain] = pk;  // quirk 2: one shared index, the last arm's draw
    }
  }

  // ================= beta Metropolis steps, utils_ct.cpp:750-791 =================
  // The "current" sum of category kk (log_den's data part) is the same for q = 0..Q-1 until a
  // proposal is accepted, and then equals the accepted proposal's sum: the per-patient terms are
  // kept in scratch (ljs, cu, pr) so that each (kk, q) step costs one exp + one lgamma per patient.
  if (b.model.noprog != 1) {
    const double *sigma_beta = b.sigma_beta + (size_t)chain * Q * D;
    double *ljs = cur, *cu = prop, *pr = prop + (size_t)T * ntm;  // >= nobs entries each
    // the proposal normal and the acceptance uniform of every (k, q) step, drawn up front by
    // Q*D threads in parallel (they do not depend on the chain's evolving state)
    double *bn = Lsig + T * D * D + 16, *bu = bn + TPPMX_MAXD * TPPMX_MAXQ;
    for (int e = tid; e < Q * D; e += nw) {
      bn[e] = upd_n2(rng, l, TPPMX_SITE_BETA, 0, (uint32_t)e, 0).x;
      bu[e] = upd_log(upd_u2(rng, l, TPPMX_SITE_BETA, 0, (uint32_t)e, 1).x);
    }
    bar_workers(nw);
    for (int kk = 0; kk < D; kk++) {
      double sd = 0.0, dummy = 0.0;
      for (int i = tid; i < nobs; i += nw) {
        const double lj = upd_log(JJ[(size_t)i * D + kk]);
        const double g0 = upd_exp(lg[(size_t)i * D + kk]);
        const double c = g0 * lj - upd_lgamma(g0);
        ljs[i] = lj;
        cu[i] = c;
        sd += c;
      }
      workers_sum2(sd, dummy, red, nw);
      for (int q = 0; q < Q; q++) {
        const int h = kk + q * D;
        const double bh = beta[h];
        const double beta_p = bh + mh_beta * bn[kk * Q + q];
        double sn = 0.0;
        dummy = 0.0;
        for (int i = tid; i < nobs; i += nw) {
          const double zq = z[(size_t)i * Q + q];
          const double lgp = __dadd_rn(__dsub_rn(lg[(size_t)i * D + kk], __dmul_rn(bh, zq)), __dmul_rn(beta_p, zq));  // :764
          const double g1 = upd_exp(lgp);
          const double c = g1 * ljs[i] - upd_lgamma(g1);
          pr[i] = c;
          sn += c;
        }
        workers_sum2(sn, dummy, red, nw);
        // dnorm(beta_p; 0, sb) - dnorm(beta; 0, sb) on the log scale: the log(sb) terms cancel (:758, :771)
        const double isb = 1.0 / sigma_beta[h];
        const double tp = beta_p * isb, tb = bh * isb;
        const double ln_acp = (sn - sd) - 0.5 * (tp * tp - tb * tb);
        if (bu[kk * Q + q] < ln_acp) {  // same inputs in every worker thread: uniform branch
          for (int i = tid; i < nobs; i += nw) {
            const double zq = z[(size_t)i * Q + q];
            lg[(size_t)i * D + kk] = __dadd_rn(__dsub_rn(lg[(size_t)i * D + kk], __dmul_rn(bh, zq)), __dmul_rn(beta_p, zq));
            cu[i] = pr[i];
          }
          sd = sn;
          if (tid == 0) {
            beta[h] = beta_p;
            b.beta_flag[(size_t)chain * Q * D + q + Q * kk] += 1;
          }
        }
        bar_workers(nw);
      }
    }
  }

  (void)P;
  (void)y;
  (void)ss;
}

// ================= horseshoe, utils_ct.cpp:876-915 =================
// lambda_h | beta, tau (slice step), tau | beta, lambda (truncated-gamma inverse CDF), one thread
// per chain.  A strictly serial chain of ~10 transcendental calls and an incomplete-gamma
// inversion: inside update_kernel it sat on the chain's critical path with one active lane;
// here the chains fill whole warps and run beside the gamma draws (nothing reads
// lambda / tau / sigma_beta before the next iteration's beta step).
__device__ __noinline__ void horseshoe_chain(const DevBatch &b, const Rng &rng, int chain, int l) {
  const int Q = b.Q, D = b.D;
  const double *beta = b.beta + (size_t)chain * Q * D;
  if (b.model.noprog != 1 && b.model.hsp == 1) {
    double *sigma_beta = b.sigma_beta + (size_t)chain * Q * D;
    const int len = Q * D;
    double *lambda = b.lambda + (size_t)chain * len;
    double tau = b.tau[chain];
    #pragma unroll 1
    for (int k = 0; k < len; k++) {
      double u0, u1;
      { const double2 r2_ = upd_u2(rng, l, TPPMX_SITE_HS_LAMBDA, 0, (uint32_t)k, 0); u0 = r2_.x; u1 = r2_.y; }
      double et = 1.0 / (lambda[k] * lambda[k]);
      const double upsi = u0 * (1.0 / (1.0 + et));
      const double tempps = (beta[k] * beta[k]) / (2.0 * tau * tau);
      const double ub = (1.0 - upsi) / upsi;
      double Fub = 1.0 - upd_exp(-tempps * ub);
      if (Fub < 1e-4) Fub = 1e-4;
      const double up = u1 * Fub;
      et = -upd_log(1.0 - up) / tempps;
      lambda[k] = 1.0 / sqrt(et);
    }
    double tempt = 0.0, u0, u1;
    #pragma unroll 1
    for (int k = 0; k < len; k++) {
      const double r = beta[k] / lambda[k];
      tempt += r * r;
    }
    tempt = tempt / 2.0;
    double et = 1.0 / (tau * tau);
    { const double2 r2_ = upd_u2(rng, l, TPPMX_SITE_HS_TAU, 0, 0, 0); u0 = r2_.x; u1 = r2_.y; }
    const double utau = u0 * (1.0 / (1.0 + et));
    const double ubt = (1.0 - utau) / utau;
    const double shape = ((double)len + 1.0) / 2.0;
    double Fubt = upd_gammainc(shape, ubt * tempt);
    if (Fubt < 1e-4) Fubt = 1e-4;
    const double ut = u1 * Fubt;
    et = upd_gammaincinv(shape, ut) / tempt;
    if (et < .0001) et = .0001;
    tau = 1.0 / sqrt(et);
    if (tau < .0001) tau = .0001;
    b.tau[chain] = tau;
    #pragma unroll 1
    for (int k = 0; k < len; k++) sigma_beta[k] = lambda[k] * tau;
  }
}

// ================= JJ, ss gamma draws, src/dm_ppmx_ct.cpp:1042-1055 =================
// One thread per (chain, patient), flattened over the whole batch: the draws of a patient only
// read its own loggamma row and ss, so nothing ties them to the chain's CTA — a separate small
// kernel keeps the update kernel's instruction footprint down and fills the machine.
// The first nhs = ceil(n_chains / 128) blocks run the horseshoe update instead (one thread per
// chain): they are the long-latency blocks, so they are scheduled first.  nhs = 0 when the
// horseshoe runs as its own kernel on another stream (tppmx_batch_run).
__global__ void __launch_bounds__(128) horseshoe_kernel(DevBatch b, uint64_t seed, int l) {
  const int chain = blockIdx.x * 128 + threadIdx.x;
  if (chain < b.n_chains) horseshoe_chain(b, make_rng(seed, b.chain_id[chain]), chain, l);
}
__global__ void __launch_bounds__(128) jjss_kernel(DevBatch b, uint64_t seed, int l, int nhs) {
  if ((int)blockIdx.x < nhs) {
    const int chain = blockIdx.x * 128 + threadIdx.x;
    if (chain < b.n_chains) horseshoe_chain(b, make_rng(seed, b.chain_id[chain]), chain, l);
    return;
  }
  const size_t gi = (size_t)(blockIdx.x - nhs) * blockDim.x + threadIdx.x;
  if (gi >= (size_t)b.n_chains * b.nobs) return;
  const int chain = (int)(gi / b.nobs), i = (int)(gi - (size_t)chain * b.nobs);
  const int ds = b.chain_ds[chain], D = b.D;
  const Rng rng = make_rng(seed, b.chain_id[chain]);
  const double *lg = b.lg + ((size_t)chain * b.nobs + i) * D;
  const double *y = b.y + ((size_t)ds * b.nobs + i) * D;
  double *JJ = b.JJ + ((size_t)chain * b.nobs + i) * D;
  double *ss = b.ss + (size_t)chain * b.nobs + i;
  const double sc = 1.0 / (*ss + 1.0);
  double TT = 0.0;
  for (int k = 0; k < D; k++) {
    const double shape = y[k] + fast_exp(lg[k]);
    double v = rng_gamma(rng, l, TPPMX_SITE_JJ, 0, (uint32_t)(i * D + k), shape) * sc;
    if (v < 1e-100) v = 1e-100;
    JJ[k] = v;
    TT += v;
  }
  *ss = rng_gamma(rng, l, TPPMX_SITE_SS, 0, (uint32_t)i, 1.0) * (1.0 / TT);
}

}  // namespace tppmx
