// layout.h — device-side data layout of a batch of chains (private to the library).
//
// HBM layout (all arrays chain-major, one contiguous slab per field):
//   datasets : x [nd][nobs][P]   z [nd][nobs][Q]   y [nd][nobs][D]   (patient-major rows,
//              read-only, shared by every chain of the dataset -> L2 resident)
//              arm_pat [nd][T][ntmax] patient index of the it-th patient of arm t
//   chains   : cluster sufficient statistics in structure-of-arrays form
//              sumx, sumx2 [nc][T][P][kcap]   njc [nc][T][ncat][maxC][kcap]
//              eta [nc][T][D][kcap]  nj [nc][T][kcap]   -> lane j of a warp reads cluster j:
//              consecutive addresses, coalesced in HBM and conflict-free when staged in smem.
//              The reference stores them cluster-major (sumx(t, j*P+p), src/dm_ppmx_ct.cpp:179).
//   labels   : [nc][T][ntmax] int32, 1-based like the reference's curr_clu (:113)
//   patient  : JJ, loggamma [nc][nobs][D]   ss [nc][nobs]
#pragma once
#include <stdint.h>

#include "../../include/tppmx.h"

namespace tppmx {

#define TPPMX_NNIG_ROW 16  // doubles per n of the NNIG table (tppmx.cu:fast_reserve, sim.cuh:gsim_nnig_row)

struct NNRow {  // per-n constants of gsimconNN (utils_ct.cpp:382-405), see sim.cuh
  double A0, H, B0, H2;
};

struct DevBatch {
  tppmx_model model;
  int nobs, T, D, Q, P, ncat, maxC, npred, G, ntmax, kcap, CC;
  int n_chains, n_datasets;
  // shared
  const int *catvec;
  const double *grid;
  const double *logV;
  const NNRow *nn_tab;  // [ntmax+2]
  const double *nnig_tab;  // [ntmax+2][TPPMX_NNIG_ROW] per-n terms of gsimconNNIG (consim == 2, tppmx.cu:fast_reserve) or nullptr
  // datasets
  const int *arm_n;     // [nd][T]
  const int *arm_pat;   // [nd][T][ntmax]
  const int *treat;     // [nd][nobs]
  const double *x;       // [nd][nobs][P]  (x^2 is formed on the device with a separately rounded multiply)
  const int *xcat;      // [nd][nobs][ncat]
  const double *z;      // [nd][nobs][Q]
  const double *y;      // [nd][nobs][D]
  const double *zpred;  // [nd][npred][Q]
  const double *xp;      // [nd][npred][P]
  const int *xcatp;     // [nd][npred][ncat]
  // chains (n_chains + 1 slots; the last one is the scratch slot of tppmx_score_patient)
  const int *chain_ds;
  const uint32_t *chain_id;
  int *labels, *nj, *nclu, *njc, *idxs, *status;
  double *sumx, *sumx2, *eta, *etae, *JJ, *ss, *lg, *beta, *sigma, *kappa, *Theta, *Sigma;
  double *lambda, *tau, *sigma_beta;
  int *eta_flag;   // [nc][T][kcap]  eta* acceptances per cluster slot (reference eta_flag, :145)
  int *beta_flag;  // [nc][Q*D]      beta acceptances, index q + Q*k (reference beta_flag(q,k), :150)
  // per-chain scratch
  double *zb;    // [nc][nobs][D]   sum_q beta[k+qD] z_iq
  double *ljj;   // [nc][nobs][D]   log JJ_ik
  double *cpat;  // [nc][nobs]      cluster-independent part of the likelihood term (:572-575)
  double *wbuf;  // [nc][T][3][kcap+CC]  weights / gtilN / gtilY
};

// strides (elements per chain)
__host__ __device__ inline size_t st_labels(const DevBatch &b) { return (size_t)b.T * b.ntmax; }
__host__ __device__ inline size_t st_nj(const DevBatch &b) { return (size_t)b.T * b.kcap; }
__host__ __device__ inline size_t st_sumx(const DevBatch &b) { return (size_t)b.T * (b.P > 0 ? b.P : 1) * b.kcap; }
__host__ __device__ inline size_t st_njc(const DevBatch &b) {
  return (size_t)b.T * (b.ncat > 0 ? b.ncat * b.maxC : 1) * b.kcap;
}
__host__ __device__ inline size_t st_eta(const DevBatch &b) { return (size_t)b.T * b.D * b.kcap; }
__host__ __device__ inline size_t st_etae(const DevBatch &b) { return (size_t)b.T * b.D * b.CC; }
__host__ __device__ inline size_t st_pat(const DevBatch &b) { return (size_t)b.nobs * b.D; }
__host__ __device__ inline size_t st_wbuf(const DevBatch &b) { return (size_t)b.T * 3 * (b.kcap + b.CC); }

}  // namespace tppmx
