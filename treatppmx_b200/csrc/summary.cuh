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

// pisum[ds][(t*D+k)*npred + pp] += pipred of every chain of the dataset: added by the predictive kernels themselves
// (predict.cuh, one RED.F64 per value) -- no per-chain cube is written out and read back.

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

// Point-estimate partition from the posterior similarity matrix: what the package's
// users compute next with mcclust.ext::minVI / minbinder.ext (imported R/treatppmx.R:30-31; the
// package is un-vendored and unpinned, GitHub sarawade/mcclust.ext, so its PUBLISHED algorithm is
// restated: Wade & Ghahramani 2018, "Bayesian cluster analysis: point estimation and credible
// balls", Bayesian Analysis 13(2), section 4, and the package's default method "avg" / "comp"):
//   1. agglomerative clustering of the n_t patients of an arm on the dissimilarity 1 - psm, average
//      (UPGMA) or complete linkage, Lance-Williams updates;
//   2. for every cut with k = 1..max_k clusters (default max_k = ceiling(n_t / 8)) the loss
//        VI lower bound  (1/n) sum_i [ log2 #{j: c_j = c_i} + log2 sum_j psm_ij - 2 log2 sum_j 1{c_j = c_i} psm_ij ]
//        Binder          sum_{i<j} | 1{c_i = c_j} - psm_ij |
//   3. the cut with the smallest loss (the smallest k among equals, as which.min over k = 1..max_k).
// Ties between equal dissimilarities are broken towards the smallest (i, j), i < j, in index order
// (R's hclust leaves them to its scan order; exact ties only occur between patients that were never
// separated or never together).  One CTA per (dataset, arm); the n_t x n_t dissimilarities live in
// shared memory (n_t <= TPPMX_PE_MAXN).  Labels are numbered in order of first appearance.
#define TPPMX_PE_MAXN 144
__global__ void point_estimate_kernel(DevBatch b, const double *psm_mean, int loss_kind, int linkage, int max_k_in,
                                      int *labels_out, int *k_out, double *loss_out) {
  extern __shared__ __align__(16) char pe_raw[];
  const int ds = blockIdx.x, tt = blockIdx.y, tid = threadIdx.x, nth = blockDim.x;
  const int n = b.arm_n[(size_t)ds * b.T + tt], ld = b.ntmax;
  const double *psm = psm_mean + ((size_t)ds * b.T + tt) * ld * ld;
  double *dist = reinterpret_cast<double *>(pe_raw);             // [n][n]
  double *rowsum = dist + (size_t)n * n;                          // [n] sum_j psm_ij
  double *red = rowsum + n;                                       // [nth] reduction scratch
  int *redi = reinterpret_cast<int *>(red + nth);                 // [nth]
  int *cl = redi + nth;                                           // [n] current cluster id of item i (= smallest member)
  int *size = cl + n;                                             // [n] size of the cluster whose id is i
  int *best = size + n;                                           // [n] best partition so far
  __shared__ double s_bestloss;
  __shared__ int s_bestk, s_mi, s_mj;
  int *out_lab = labels_out + ((size_t)ds * b.T + tt) * ld;
  if (n <= 0) {
    if (tid == 0) {
      k_out[(size_t)ds * b.T + tt] = 0;
      loss_out[(size_t)ds * b.T + tt] = 0.0;
    }
    return;
  }
  const int max_k = max_k_in > 0 ? (max_k_in < n ? max_k_in : n) : (n + 7) / 8;
  for (int e = tid; e < n * n; e += nth) {
    const int i = e / n, j = e - i * n;
    dist[e] = 1.0 - psm[(size_t)i * ld + j];
  }
  for (int i = tid; i < n; i += nth) {
    double s = 0.0;
    for (int j = 0; j < n; j++) s += psm[(size_t)i * ld + j];
    rowsum[i] = s;
    cl[i] = i;
    size[i] = 1;
    best[i] = 0;
  }
  if (tid == 0) {
    s_bestloss = INFINITY;
    s_bestk = 0;
  }
  __syncthreads();

  for (int k = n; k >= 1; k--) {
    // ---- loss of the current partition (k clusters), if it is a candidate cut
    if (k <= max_k) {
      double part = 0.0;
      for (int i = tid; i < n; i += nth) {
        const int ci = cl[i];
        if (loss_kind == 1) {
          double within = 0.0;
          for (int j = 0; j < n; j++)
            if (cl[j] == ci) within += psm[(size_t)i * ld + j];
          part += (log2((double)size[ci]) + log2(rowsum[i]) - 2.0 * log2(within)) / (double)n;
        } else {
          for (int j = i + 1; j < n; j++) part += fabs((cl[j] == ci ? 1.0 : 0.0) - psm[(size_t)i * ld + j]);
        }
      }
      red[tid] = part;
      __syncthreads();
      if (tid == 0) {  // fixed summation order: the value does not depend on the launch shape beyond nth
        double tot = 0.0;
        for (int r = 0; r < nth; r++) tot += red[r];
        if (tot <= s_bestloss) {   // going down in k: the smaller k wins ties
          s_bestloss = tot;
          s_bestk = k;
        }
      }
      __syncthreads();
      if (s_bestk == k)
        for (int i = tid; i < n; i += nth) best[i] = cl[i];
      __syncthreads();
    }
    if (k == 1) break;
    // ---- closest pair of active clusters (ids = smallest member), smallest (i, j) on ties
    double dmin = INFINITY;
    int emin = 0x7fffffff;
    for (int e = tid; e < n * n; e += nth) {
      const int i = e / n, j = e - i * n;
      if (i < j && cl[i] == i && cl[j] == j) {
        const double d = dist[e];
        if (d < dmin || (d == dmin && e < emin)) {
          dmin = d;
          emin = e;
        }
      }
    }
    red[tid] = dmin;
    redi[tid] = emin;
    __syncthreads();
    if (tid == 0) {
      double d0 = red[0];
      int e0 = redi[0];
      for (int r = 1; r < nth; r++)
        if (red[r] < d0 || (red[r] == d0 && redi[r] < e0)) {
          d0 = red[r];
          e0 = redi[r];
        }
      s_mi = e0 / n;
      s_mj = e0 - (e0 / n) * n;
    }
    __syncthreads();
    const int mi = s_mi, mj = s_mj, ni = size[mi], nj = size[mj];
    // ---- Lance-Williams update of the merged cluster (keeps id mi), then relabel
    for (int c = tid; c < n; c += nth) {
      if (cl[c] != c || c == mi || c == mj) continue;
      const double di = dist[(size_t)(c < mi ? c : mi) * n + (c < mi ? mi : c)];
      const double dj = dist[(size_t)(c < mj ? c : mj) * n + (c < mj ? mj : c)];
      // separately rounded operations (no fused multiply-add): merge order must not depend on contraction
      const double dn = linkage == 1 ? fmax(di, dj)
                                     : __ddiv_rn(__dadd_rn(__dmul_rn((double)ni, di), __dmul_rn((double)nj, dj)), (double)(ni + nj));
      dist[(size_t)(c < mi ? c : mi) * n + (c < mi ? mi : c)] = dn;
    }
    __syncthreads();
    for (int i = tid; i < n; i += nth)
      if (cl[i] == mj) cl[i] = mi;
    if (tid == 0) size[mi] = ni + nj;
    __syncthreads();
  }
  // ---- labels in order of first appearance (1-based)
  if (tid == 0) {
    int next = 0;
    for (int i = 0; i < n; i++) {
      const int c = best[i];
      int lab = 0;
      for (int j = 0; j < i; j++)
        if (best[j] == c) {
          lab = out_lab[j];
          break;
        }
      if (lab == 0) lab = ++next;
      out_lab[i] = lab;
    }
    for (int i = n; i < ld; i++) out_lab[i] = 0;
    k_out[(size_t)ds * b.T + tt] = s_bestk;
    loss_out[(size_t)ds * b.T + tt] = s_bestloss;
  }
}
static inline size_t point_estimate_smem(int n, int nth) {
  return sizeof(double) * ((size_t)n * n + n + nth) + sizeof(int) * ((size_t)nth + 3 * (size_t)n) + 16;
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
