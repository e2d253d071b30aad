// sweep_fast_args.h — host-visible part of the fast sweep kernel (sweep_fast.cuh): launch
// arguments, shared-memory sizing and the per-LPU launchers.  The kernel templates are
// instantiated in their own translation units (sweep_fast_lpu{8,16,32}.cu) so the library builds
// in parallel; tppmx.cu only needs this header.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "layout.h"

namespace tppmx {

#define TPPMX_FAST_KSMAX 32
#define TPPMX_FAST_KSTR 33  // smem stride of the [P][cluster] arrays; column 32 is an all-zero cluster

struct FastArgs {
  double *lik;     // [units][ncol][ntp] (+ntp pad)  likelihood table (global, L2 resident)
  double *uq;      // [units][3][ntp+1]    uniforms, sum_p x_ip^2, x_{it-1} . x_it (the long-arm kernel uses two rows)
  int *resume_l;   // [units] sweep at which the unit handed over to the generic kernel, -1 = done
  int *resume_it;  // [units] patient position of the hand-over
  int *bail_count; // [1]
  int KS, ntp, ncol, n_units;
  int unit_bytes, tab_bytes;
  // big arms (BASELINE config 4: n_t in the thousands): everything indexed by patient position or
  // by cluster size would not fit in shared memory, so it lives in global memory (L2-resident):
  int big;          // 1: the three tables over n and labels / patient order / cohesion table are global
  double *tabg;     // [ntmax + 2][4]  (dA, HN, HY, -) per cluster size n (filled once at batch creation)
  double *logng;    // [units][ntp + 1]
  int *patpad;      // [units][ntp + 1] patient order of the unit, zero-padded
  // parity probe (tppmx_batch_sweep_probe): the PROBE instantiations record, for the units
  // probe_unit0 .. probe_unit0 + probe_nunits - 1, every step's normalised allocation probabilities
  // (exactly the e / den the draw is made from) and the number of occupied clusters after removal
  double *probe;    // [probe_nunits][ntp][probe_stride], nullptr = production instantiation
  int *probe_k;     // [probe_nunits][ntp], -1 = step not visited by the fast kernel (hand-over)
  int probe_stride, probe_unit0, probe_nunits;
  // long-arm kernel (sweep_big.cuh): the staged capacity KS is chosen per launch, the likelihood table
  // keeps the stride of the largest one; kmax_dev receives the largest K any unit reached
  // categorical covariates (Dirichlet-multinomial similarity, utils_ct.cpp:468-495) in the fast kernel: with the
  // patient's category c* in covariate p, lgY - lgN = log(n_jc* + w) - log(n_j + C_p w) -- two table entries
  const double *cat_lw;    // [ntmax + 2]          log(m + dirw)
  const double *cat_ln;    // [ncat][ntmax + 2]    log(n + C_p dirw)
  size_t lik_unit_stride;  // doubles per unit of `lik`
  int *kmax_dev;
  int resume_mode;         // 1: run only the units with resume_l >= 0, from (resume_l, resume_it) (second stage)
};

__host__ __device__ inline int fast_pp(int P) { return (P + 1) & ~1; }
// bytes of shared memory of one unit
// ncat / maxC: categorical covariates staged by the CAT instantiation (counts [ncat * maxC][KSTR] + a ring of four rows)
__host__ __device__ inline size_t fast_unit_bytes(int P, int KS, int ntp, int CC, int D, int big, int ncat = 0, int maxC = 0) {
  size_t dbl = (size_t)2 * P * TPPMX_FAST_KSTR + (big ? 0 : (ntp + 1)) + 4 * fast_pp(P) + (KS + CC) + (size_t)(KS + CC) * D;
  size_t in = (size_t)KS + 2 * (KS + CC) + (big ? 0 : 2 * (ntp + 1)) + 1 + (size_t)ncat * maxC * TPPMX_FAST_KSTR + 4 * (size_t)ncat;
  return ((dbl * 8 + in * 4) + 15) & ~(size_t)15;
}
// per-n table (dA, HN, HY, P A0) of the Normal-Normal similarity, staged in shared memory for short arms.  The
// Normal-Normal-Inverse-Gamma similarity (consim == 2) reads its 16-entry rows from global memory (L1).
__host__ __device__ inline size_t fast_tab_bytes(int ntmax, int big, int nnig = 0) {
  return (big || nnig) ? 16 : (size_t)4 * (ntmax + 2) * 8;
}

// Launch for LPU lanes per (chain, arm) on `stream`; picks the instantiation from the batch shape
// (P -> XR, D -> DR, model.nolik, fa.big).  fa.big requires LPU = 32.
cudaError_t fast_launch_lpu8(const DevBatch &d, const FastArgs &fa, size_t smem, uint64_t seed, int iter0, int n_sweeps, cudaStream_t stream);
cudaError_t fast_launch_lpu16(const DevBatch &d, const FastArgs &fa, size_t smem, uint64_t seed, int iter0, int n_sweeps, cudaStream_t stream);
cudaError_t fast_launch_lpu32(const DevBatch &d, const FastArgs &fa, size_t smem, uint64_t seed, int iter0, int n_sweeps, cudaStream_t stream);

// Long arms (fa.big): sweep_big_kernel (sweep_big.cuh), one CTA of nth = 128 | 256 threads per (chain, arm),
// up to TPPMX_BIG_KSMAX staged clusters.
#define TPPMX_BIG_KSMAX 256
cudaError_t big_launch(const DevBatch &d, const FastArgs &fa, int nth, size_t smem, uint64_t seed, int iter0, int n_sweeps, cudaStream_t stream);
// row stride (doubles) of the long-arm kernel's t matrix [cluster][covariate]: even (16-byte aligned rows for
// 128-bit loads) with an odd half (rows of consecutive clusters start in different 16-byte bank groups)
__host__ __device__ inline int big_pstr(int P) { const int h = (P + 1) >> 1; return 2 * (h | 1); }
// bytes of dynamic shared memory of one CTA of the long-arm kernel (layout: sweep_big.cuh)
__host__ __device__ inline size_t big_smem_bytes(int P, int KS, int CC, int D, int nth) {
  (void)nth;
  const size_t NC = (size_t)KS + CC, NO = 16;
  const size_t dbl = (size_t)(KS + 1) * big_pstr(P) + ((NC * D + 1) & ~(size_t)1) + 2 * fast_pp(P) + 2 * NC + NC + NC + (KS + 1) + 4 + NO + 4;
  const size_t in = (size_t)KS + 2 * NC + 8 + 2 + 4 + 2;
  return ((dbl * 8 + in * 4) + 15) & ~(size_t)15;
}

}  // namespace tppmx
