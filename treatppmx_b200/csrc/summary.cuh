// summary.cuh — posterior summaries accumulated on the device over saved iterations and chains:
//   * co-clustering (posterior similarity) counts per replicate dataset and arm:
//       psm[ds][t][a][b] += 1{ label_a == label_b }      (what mcclust::comp.psm computes from the
//       reference's cl_lab output, src/dm_ppmx_ct.cpp:1402-1404; imported R/treatppmx.R:29)
//   * posterior-predictive probability sums  pisum[ds][t][k][pp] += pipred  (:1373-1379), from which
//     the mean, the utility sum_k w_k E[pi_k] and the arg-max arm follow (SURVEY §8a "treatment utility").
// These are what config 5 gathers across GPUs; nothing else of a chain has to leave the device.
#pragma once
#include "layout.h"

namespace tppmx {

// grid = (chains, T), threads stride over the n_t x n_t pairs of the arm
__global__ void psm_accumulate_kernel(DevBatch b, int *psm) {
  const int chain = blockIdx.x, tt = blockIdx.y;
  const int ds = b.chain_ds[chain];
  const int nt = b.arm_n[(size_t)ds * b.T + tt];
  const int ld = b.ntmax;
  const int *lab = b.labels + ((size_t)chain * b.T + tt) * b.ntmax;
  int *out = psm + ((size_t)ds * b.T + tt) * ld * ld;
  for (int e = threadIdx.x; e < nt * nt; e += blockDim.x) {
    const int a = e / nt, c = e - a * nt;
    if (lab[a] == lab[c]) atomicAdd(out + (size_t)a * ld + c, 1);
  }
}

// pisum[ds][(t*D+k)*npred + pp] += pipred[chain][...]   (pipred is the col-major npred x D x T cube)
__global__ void pi_accumulate_kernel(DevBatch b, const double *pipred, double *pisum) {
  const int chain = blockIdx.x;
  const int ds = b.chain_ds[chain];
  const size_t n1 = (size_t)b.npred * b.D * b.T;
  for (size_t e = threadIdx.x; e < n1; e += blockDim.x) atomicAdd(pisum + (size_t)ds * n1 + e, pipred[(size_t)chain * n1 + e]);
}

// grid = datasets: means, utilities, chosen arm
__global__ void summarise_kernel(DevBatch b, const int *psm, const int *psm_n, const double *pisum,
                                 const int *pi_n, const double *util_w, double *psm_mean,
                                 double *pi_mean, double *utility, int *best_arm) {
  const int ds = blockIdx.x;
  const int T = b.T, D = b.D, npred = b.npred, ld = b.ntmax;
  if (psm_mean) {
    const double inv = psm_n[ds] > 0 ? 1.0 / (double)psm_n[ds] : 0.0;
    const size_t n = (size_t)T * ld * ld;
    for (size_t e = threadIdx.x; e < n; e += blockDim.x) psm_mean[(size_t)ds * n + e] = inv * (double)psm[(size_t)ds * n + e];
  }
  if (pi_mean && npred > 0) {
    const double inv = pi_n[ds] > 0 ? 1.0 / (double)pi_n[ds] : 0.0;
    const size_t n1 = (size_t)npred * D * T;
    for (size_t e = threadIdx.x; e < n1; e += blockDim.x) pi_mean[(size_t)ds * n1 + e] = inv * pisum[(size_t)ds * n1 + e];
    __syncthreads();
    for (int pp = threadIdx.x; pp < npred; pp += blockDim.x) {
      int best = 0;
      double ub = -INFINITY;
      for (int t = 0; t < T; t++) {
        double u = 0.0;
        for (int k = 0; k < D; k++) u += util_w[k] * pi_mean[(size_t)ds * n1 + pp + (size_t)npred * (k + (size_t)D * t)];
        utility[((size_t)ds * T + t) * npred + pp] = u;
        if (u > ub) {
          ub = u;
          best = t;
        }
      }
      best_arm[(size_t)ds * npred + pp] = best;
    }
  }
}

// npc (R/countUT.R:26-53; Ma, Stingo & Hobbs 2019): per replicate dataset, the number of new
// patients whose most probable outcome category under the arm they were assigned to by design
// (which.max over the D predictive means, first maximum on ties, :37-38) equals the outcome that
// was observed.  pi_mean is the [npred x D x T] col-major posterior mean of pipred.
__global__ void npc_kernel(DevBatch b, const double *pi_mean, const int *trtsgn, const int *outcome, int *npc) {
  const int ds = blockIdx.x, T = b.T, D = b.D, npred = b.npred;
  __shared__ int cnt;
  if (threadIdx.x == 0) cnt = 0;
  __syncthreads();
  const size_t n1 = (size_t)npred * D * T;
  int mine = 0;
  for (int pp = threadIdx.x; pp < npred; pp += blockDim.x) {
    const int t = trtsgn[(size_t)ds * npred + pp];
    if (t < 0 || t >= T) continue;
    int best = 0;
    double pb = -INFINITY;
    for (int k = 0; k < D; k++) {
      const double v = pi_mean[(size_t)ds * n1 + pp + (size_t)npred * (k + (size_t)D * t)];
      if (v > pb) {
        pb = v;
        best = k;
      }
    }
    mine += (best == outcome[(size_t)ds * npred + pp]);
  }
  atomicAdd(&cnt, mine);
  __syncthreads();
  if (threadIdx.x == 0) npc[ds] = cnt;
}

// per-dataset save counters: how many (chain, save) pairs went into the sums
__global__ void count_saves_kernel(DevBatch b, int *psm_n, int *pi_n, int do_psm, int do_pi) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= b.n_chains) return;
  const int ds = b.chain_ds[c];
  if (do_psm) atomicAdd(psm_n + ds, 1);
  if (do_pi) atomicAdd(pi_n + ds, 1);
}

}  // namespace tppmx
