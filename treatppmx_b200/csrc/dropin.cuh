// dropin.cuh — per-saved-iteration work of the single-chain drop-in (tppmx_dm_ppmx_ct):
//   in-sample fit            reference src/dm_ppmx_ct.cpp:1060-1075 (+ dmultinom_rcpp, utils_ct.cpp:842-874)
//   save of the iterate      reference src/dm_ppmx_ct.cpp:1391-1429
// The traces of all saved iterations stay on the device, in the reference's output layouts, and
// are downloaded once at the end of the run.
#pragma once
#include "layout.h"
#include "rng.cuh"

namespace tppmx {

struct TraceBufs {  // device, sized for nout saved iterations (col-major, see include/tppmx.h)
  double *eta, *beta, *sigma, *kappa, *cl_lab, *pi, *pipred, *yispred, *ypred, *clupred, *nclus, *lpml;
  double *mnlike, *mnllike;  // [nobs] running means for WAIC (:1068-1069)
  double *sumtotclu;         // [T]  (:877)
  int nout;
};

// after every iteration: sumtotclu(tt) += nclu_curr(tt)   (:877)
__global__ void sumtotclu_kernel(DevBatch b, TraceBufs tr) {
  const int t = threadIdx.x;
  if (t < b.T) tr.sumtotclu[t] += (double)b.nclu[t];
}

// chain 0 only; one block.  ll = index of the saved iteration.
__global__ void insample_save_kernel(DevBatch b, TraceBufs tr, uint64_t seed, int l, int ll,
                                     const double *pipred_it, const double *ypred_it, const int *predclass_it) {
  __shared__ double red[256];
  const int tid = threadIdx.x, nth = blockDim.x;
  const int nobs = b.nobs, T = b.T, D = b.D, Q = b.Q, ntm = b.ntmax, ks = b.kcap, npred = b.npred, nout = tr.nout;
  const Rng rng = make_rng(seed, b.chain_id[0]);
  double lp = 0.0;
  for (int i = tid; i < nobs; i += nth) {  // :1061-1073
    double TT = 0.0;
    for (int k = 0; k < D; k++) TT += b.JJ[(size_t)i * D + k];
    double u0, u1, cw = 0.0, like = 0.0, ssum = 0.0;
    rng_u2(rng, l, TPPMX_SITE_ISPRED, 0, (uint32_t)i, 0, u0, u1);
    int pick = D - 1;
    bool found = false;
    for (int k = 0; k < D; k++) ssum += b.JJ[(size_t)i * D + k] / TT;  // dmultinom_rcpp renormalises (:844-845)
    double r = 0.0;  // lgamma(size + 1) = 0
    for (int k = 0; k < D; k++) {
      const double pk = b.JJ[(size_t)i * D + k] / TT;
      tr.pi[i + (size_t)nobs * (k + (size_t)D * ll)] = pk;  // pigreco (:1396)
      cw += pk;
      if (!found && u0 < cw) {
        pick = k;
        found = true;
      }
      const double x = floor(b.y[((size_t)0 * nobs + i) * D + k] + .5);
      if (x != 0.0) r += x * log(pk / ssum) - lgamma(x + 1.0);  // :867 (0 * log(p) terms vanish)
    }
    like = exp(r);
    for (int k = 0; k < D; k++) tr.yispred[i + (size_t)nobs * (k + (size_t)D * ll)] = (k == pick) ? 1.0 : 0.0;
    tr.mnlike[i] += like / (double)nout;
    tr.mnllike[i] += log(like) / (double)nout;
    lp += log(like);  // CPOinv(i) = log(like_iter(i)); lpml(ll) = sum (:1072-1074)
  }
  red[tid] = lp;
  __syncthreads();
  for (int o = nth >> 1; o > 0; o >>= 1) {
    if (tid < o) red[tid] += red[tid + o];
    __syncthreads();
  }
  if (tid == 0) tr.lpml[ll] = red[0];
  // ---- save of the iterate (:1391-1429)
  for (int e = tid; e < T; e += nth) {
    tr.sigma[e + (size_t)T * ll] = b.sigma[e];
    tr.kappa[e + (size_t)T * ll] = b.kappa[e];
    tr.nclus[e + (size_t)T * ll] = (double)b.nclu[e];
  }
  for (int e = tid; e < T * ntm; e += nth) {  // Clui(tt, i, ll) for i < num_treat(tt), else 0
    const int t = e / ntm, it = e - t * ntm;
    tr.cl_lab[t + (size_t)T * (it + (size_t)ntm * ll)] = (it < b.arm_n[t]) ? (double)b.labels[e] : 0.0;
  }
  for (int e = tid; e < Q * D; e += nth) tr.beta[e + (size_t)Q * D * ll] = b.beta[e];
  // eta_out(ll, tt) (:1421-1427): ONE scratch matrix is zeroed per saved iteration and reused for
  // every arm, only its columns j < nclu(tt) being overwritten, so column j >= nclu(tt) of arm tt
  // still holds the eta of the last earlier arm that had more than j clusters (0 if none).
  for (int e = tid; e < T * D * ntm; e += nth) {
    const int t = e / (D * ntm), k = (e / ntm) % D, j = e % ntm;
    double v = 0.0;
    for (int ts = t; ts >= 0; ts--)
      if (j < b.nclu[ts] && j < ks) {
        v = b.eta[((size_t)ts * D + k) * ks + j];
        break;
      }
    tr.eta[((size_t)ll + (size_t)nout * t) * D * ntm + k + (size_t)D * j] = v;
  }
  if (npred > 0) {
    const size_t n1 = (size_t)npred * D * T;
    for (size_t e = tid; e < n1; e += nth) {
      tr.pipred[n1 * ll + e] = pipred_it[e];
      tr.ypred[n1 * ll + e] = ypred_it[e];
    }
    for (int e = tid; e < npred * T; e += nth) {  // predclass_out(ll*npred + pp, tt)
      const int pp = e % npred, t = e / npred;
      tr.clupred[(size_t)ll * npred + pp + (size_t)nout * npred * t] = (double)predclass_it[e];
    }
  }
}

}  // namespace tppmx
