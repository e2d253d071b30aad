// sweep_big.cuh — the reallocation-sweep kernel for LONG arms (BASELINE config 4: n_t in the
// thousands, P = 50 covariates, K_t from tens to hundreds of clusters): reference
// src/dm_ppmx_ct.cpp:347-860 with consim == 1, similarity == 1, the same switches and the same
// arithmetic per score as sweep_fast.cuh, laid out for a different regime.
//
// With thousands of steps per sweep and K * P in the thousands, a step is a matrix-vector product
// (clusters x covariates) followed by a softmax draw, and the cluster statistics of one (chain, arm)
// do not fit next to eleven other units in an SM's shared memory.  So:
//  * unit = one (chain, arm) = one CTA of NTH threads.  NTH = 32 when there are more units than an
//    SM can hold anyway (1024 chains x 3 arms: residency, not threads per unit, sets the rate --
//    measured, profiles/r2_*), NTH = 128 when units are few.  Thread t scores clusters t, t + NTH, ...:
//    ANY number of candidates up to the staged capacity KS (<= 256) is scored in one pass, no
//    32-cluster cap and no hand-over below KS.  KS is chosen PER LAUNCH from the largest K the
//    previous sweep left behind (read back without a sync), so the shared-memory footprint follows
//    the partition and 16 warps per SM stay resident at K ~ 20 as well as at K ~ 100.
//  * shared memory holds ONE matrix: t_jp = sumx_jp / v2 + m0 / s20, [P][KS | 1] (odd stride: column
//    reads by cluster and row updates by covariate are both conflict free), plus a_j = sum_p t_jp^2
//    per cluster.  A score needs only b_j = sum_p t_jp x_p (one LDS + one DFMA per covariate); a_j is
//    carried incrementally (removal: a -= iv2 (2 b + iv2 q), arrival: a = t_Y of the draw's score)
//    and t by rank-one updates; both are REBUILT from the exact statistics at every sweep start, so
//    rounding drift is bounded by one sweep (<= 1e-12 on a log-weight at n_t = 6667; every
//    probability is tested <= 1e-10 against the oracle through the probe instantiation).
//  * the exact sufficient statistics sumx / sumx2 stay in global memory (L2) and are updated by the
//    covariate owners with the reference's own operation sequence (bit-identical to the oracle);
//    their read-modify-writes are issued at the top of a step and retired at its end, one step
//    deferred for the arrival, so no L2 round trip sits on the step's dependent chain.
//  * the operands of step it+1 (covariate row, likelihood entries of every candidate, uniform,
//    label, patient index of it+2) are fetched by cp.async (LDGSTS) into a second shared-memory
//    buffer while step it computes: no register staging, no in-order stall on a dependent address.
//  * a unit that reaches the staged capacity hands over to the generic kernel at the exact (sweep,
//    patient) position, like the small-arm kernel.
#pragma once
#include "sweep_fast.cuh"

namespace tppmx {

// asynchronous global -> shared copies (LDGSTS): issued at the top of a step, waited for at its end
__device__ __forceinline__ void cp_async8(void *smem, const void *gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async4(void *smem, const void *gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

template <int NTH, int XR, int DR, bool NOLIK, bool PROBE>
__global__ void __launch_bounds__(NTH, 512 / NTH)   // 128 registers per thread: 16 warps per SM
sweep_big_kernel(DevBatch b, FastArgs fa, uint64_t seed, int iter0, int n_sweeps) {
  constexpr int NW = NTH / 32;
  constexpr int NO = 16;  // (chunk, warp) slices of the cross-slice prefix: ceil((256 + CC) / NTH) * NW <= 16
  extern __shared__ __align__(16) char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int unit = blockIdx.x;
  const int P = b.P, D = b.D, Q = b.Q, CC = b.CC, KS = fa.KS, ntp = fa.ntp, kcap = b.kcap;
  const int pstr = big_pstr(P), Pp = fast_pp(P), NC = KS + CC;
  const SimConst s = make_simconst(b);

  // ---- shared memory (big_smem_bytes)
  double *T = reinterpret_cast<double *>(smem_raw);  // [KS + 1][pstr] cluster-major rows, row KS = the empty cluster (t = c0)
  double *E = T + (size_t)(KS + 1) * pstr;           // [NC][D] exp(eta) by likelihood column
  double *XS = E + (((size_t)NC * D + 1) & ~(size_t)1);  // [2][Pp] covariate row of the current / next patient (16-byte aligned)
  double *s_lik = XS + 2 * Pp;                       // [2][NC] likelihood entries of the current / next patient
  double *s_w = s_lik + 2 * NC;                      // [NC] log-weights, then partial sums of e
  double *s_tY = s_w + NC;                           // [NC] t_Y of every candidate (the arrival value of a_j)
  double *s_a = s_tY + NC;                           // [KS + 1] a_j; [KS] = the empty cluster
  double *s_uq = s_a + KS + 1;                       // [2][2] uniform, sum_p x^2 of the current / next patient
  double *s_tot = s_uq + 4;                          // [NO] sum of e per (chunk, warp)
  float *s_max = reinterpret_cast<float *>(s_tot + NO);  // [NW <= 8] per-warp max of w (float image)
  int *s_nj = reinterpret_cast<int *>(s_tot + NO + 4);   // [KS]
  int *s_col = s_nj + KS;                            // [NC] cluster / auxiliary -> likelihood column
  int *s_free = s_col + NC;                          // [NC] free likelihood columns
  int *s_hit = s_free + NC;                          // [8] first hit of each warp
  int *s_lab = s_hit + 8;                            // [2] label of the current / next patient
  int *s_pat = s_lab + 2;                            // [4] ring of patient indices (it .. it+2)

  const int chain = unit / b.T, tt = unit % b.T;
  const int ds = b.chain_ds[chain];
  const int nt = b.arm_n[(size_t)ds * b.T + tt];
  const Rng rng = make_rng(seed, b.chain_id[chain]);

  double *g_sumx = b.sumx + ((size_t)chain * b.T + tt) * P * kcap;
  double *g_sumx2 = b.sumx2 + ((size_t)chain * b.T + tt) * P * kcap;
  int *g_nj = b.nj + ((size_t)chain * b.T + tt) * kcap;
  int *labp = b.labels + ((size_t)chain * b.T + tt) * b.ntmax;
  double *g_eta = b.eta + ((size_t)chain * b.T + tt) * D * kcap;
  double *g_etae = b.etae + ((size_t)chain * b.T + tt) * D * CC;
  const int *g_pat = b.arm_pat + ((size_t)ds * b.T + tt) * b.ntmax;
  const double *g_x = b.x + (size_t)ds * b.nobs * P;
  double *lik = fa.lik + (size_t)unit * fa.lik_unit_stride;
  const int uqs = ntp + 1;
  double *uq = fa.uq + (size_t)unit * 2 * uqs;
  double *g_ez = b.zb + (size_t)chain * b.nobs * D;  // exp(beta'z) (scratch, rebuilt per sweep)
  double *g_ljj = b.ljj + (size_t)chain * b.nobs * D;
  const double *beta = b.beta + (size_t)chain * Q * D;
  int *patp = fa.patpad + (size_t)unit * (ntp + 1);
  double *lognp = fa.logng + (size_t)unit * (ntp + 1);
  const double2 *tab2 = reinterpret_cast<const double2 *>(fa.tabg);

  // second stage of a long-arm sweep: only the units the first stage (sweep_fast_kernel, <= 32 clusters) handed
  // over, from their exact (sweep, patient) position; the other units are complete
  int l_begin = iter0, it_begin = 0;
  if (fa.resume_mode) {
    const int rl = fa.resume_l[unit];
    if (rl < 0) return;  // CTA-uniform
    l_begin = rl;
    it_begin = fa.resume_it[unit];
  }
  int K = b.nclu[(size_t)chain * b.T + tt];
  int Kpeak = K;
  const double sigma = b.sigma[(size_t)chain * b.T + tt];
  const int idxs = b.idxs[chain];
  const double iv2 = s.iv2, c0 = s.c0, hiv2 = s.hiv2;
  const double simscale = (s.calibration == 2) ? s.calscale : 1.0;

  // ---- stage what lives on chip for the whole launch
  for (int j = tid; j < KS; j += NTH) s_nj[j] = g_nj[j];
  for (int e = tid; e < 2 * Pp; e += NTH) XS[e] = 0.0;  // the pad entry of an odd P stays zero
  for (int it = tid; it <= ntp; it += NTH) patp[it] = (it < nt) ? g_pat[it] : 0;
  int nfree;
  {
    const int Kc = K < KS ? K : KS;
    for (int j = tid; j < Kc; j += NTH) s_col[j] = j;
    for (int m = tid; m < CC; m += NTH) s_col[KS + m] = Kc + m;
    nfree = NC - Kc - CC;
    for (int f = tid; f < nfree; f += NTH) s_free[f] = Kc + CC + f;
  }
  if (tid == 0) lognp[0] = (s.cohesion == 1) ? log(b.model.alpha) - log((double)CC) : 0.0;
  for (int n = 1 + tid; n <= nt; n += NTH) lognp[n] = fast_log((s.cohesion == 1) ? (double)n : (double)n - sigma);
  __syncthreads();

  auto dV_of = [&](int Kcur) -> double {
    if (s.cohesion != 2 || Kcur < 1) return 0.0;
    const size_t base = ((size_t)tt * 2) * b.ntmax;
    return b.logV[(base + b.ntmax + (Kcur - 1)) * b.G + idxs] - b.logV[(base + (Kcur - 1)) * b.G + idxs];
  };
  auto colmap = [&](int j, int Kcur) -> int { return (j < Kcur) ? s_col[j] : s_col[KS + (j - Kcur)]; };
  // a_j = sum_p t_jp^2 from the staged matrix (two accumulators, fixed order)
  auto a_of = [&](int col) -> double {
    const double *tc = T + (size_t)col * pstr;
    double a0 = 0.0, a1 = 0.0;
    int p = 0;
    for (; p + 1 < P; p += 2) {
      a0 = fma(tc[p], tc[p], a0);
      a1 = fma(tc[p + 1], tc[p + 1], a1);
    }
    if (p < P) a0 = fma(tc[p], tc[p], a0);
    return a0 + a1;
  };

  int bail_l = -1, bail_it = 0;
  bool live = true;  // CTA-uniform

  // exact statistics in global memory: the read-modify-write of an arrival is retired one step late
  int pend_c = -1;      // cluster whose sumx / sumx2 still miss the patient with covariates pend_x
  double pend_x[XR];
#pragma unroll
  for (int r = 0; r < XR; r++) pend_x[r] = 0.0;
  auto flush_pending = [&]() {
    if (pend_c >= 0) {
#pragma unroll
      for (int r = 0; r < XR; r++) {
        const int p = tid + r * NTH;
        if (p < P) {
          const size_t e = (size_t)p * kcap + pend_c;
          g_sumx[e] = __dadd_rn(g_sumx[e], pend_x[r]);
          g_sumx2[e] = __dadd_rn(g_sumx2[e], __dmul_rn(pend_x[r], pend_x[r]));
        }
      }
    }
    pend_c = -1;
  };

  for (int l = l_begin; l < iter0 + n_sweeps && live; l++) {
    const int itb = (l == l_begin) ? it_begin : 0;  // a resumed sweep starts mid-arm
    if (K > KS || (itb > 0 && K >= KS)) {
      bail_l = l;
      bail_it = itb;
      live = false;
      break;
    }
    // =========================== parallel pre-pass of the sweep ===========================
    if (!(fa.resume_mode && l == l_begin && itb > 0))  // a sweep resumed mid-arm keeps its (partly refreshed) auxiliary eta
      for (int mm = tid; mm < CC; mm += NTH)           // :351-360 auxiliary eta ~ N(0, I)
        rng_normals(rng, l, TPPMX_SITE_AUX_SWEEP, tt, (uint32_t)mm, D, g_etae + mm, CC);
    // t = sumx / v2 + m0 / s20 from the exact statistics (drift of the rank-one updates ends here)
    for (int e = tid; e < (KS + 1) * pstr; e += NTH) {
      const int j = e / pstr, p = e - j * pstr;
      T[e] = (p < P) ? fma((j < K) ? g_sumx[(size_t)p * kcap + j] : 0.0, iv2, c0) : 0.0;   // padding: t = 0 against x = 0
    }
    __syncthreads();
    for (int j = tid; j <= KS; j += NTH) s_a[j] = (j < K || j == KS) ? a_of(j) : 0.0;
    if (!NOLIK) {
      const int ntot = K + CC;
      for (int e = tid; e < ntot * D; e += NTH) {
        const int c = e / D, k = e - c * D;
        const double eta = (c < K) ? g_eta[(size_t)k * kcap + c] : g_etae[k * CC + (c - K)];
        E[colmap(c, K) * D + k] = fast_exp(eta);
      }
    }
    __syncthreads();
    {  // one thread per patient: uniform, sum_p x^2, exp(beta'z), log JJ in registers, then every candidate cluster
      const int ntot = K + CC;
      for (int it = tid; it < nt; it += NTH) {
        const int i = patp[it];
        const double *xi = g_x + (size_t)i * P;
        double qx = 0.0;
        for (int p = 0; p < P; p++) qx = fma(xi[p], xi[p], qx);
        double uu, u1;
        rng_u2(rng, l, TPPMX_SITE_ALLOC_U, tt, (uint32_t)i, 0, uu, u1);
        uq[it] = uu;
        uq[uqs + it] = qx;
        if (!NOLIK) {
          const double *zi = b.z + ((size_t)ds * b.nobs + i) * Q;
          const double *JJ = b.JJ + ((size_t)chain * b.nobs + i) * D;
          double Z[DR], ljj[DR];
#pragma unroll
          for (int k = 0; k < DR; k++) {
            Z[k] = 1.0;
            ljj[k] = 0.0;
            if (k < D) {
              double acc = 0.0;
              for (int q = 0; q < Q; q++) acc += beta[k + q * D] * zi[q];
              Z[k] = fast_exp(acc);
              ljj[k] = fast_log(JJ[k]);
              g_ez[(size_t)i * D + k] = Z[k];
              g_ljj[(size_t)i * D + k] = ljj[k];
            }
          }
          for (int c = 0; c < ntot; c++) {
            const int cl = colmap(c, K);
            lik[(size_t)cl * ntp + it] = fast_lik_reg<DR>(E + cl * D, Z, ljj, D);
          }
        }
      }
    }
    __syncthreads();

    // =========================== sequential reallocation steps ===========================
    double dV = dV_of(K);
    // operands of step 0 (synchronously), patient indices of steps 0..1
    {
      const int cb = itb & 1;
      if (tid == 0) {
        s_pat[itb & 3] = patp[itb];
        s_pat[(itb + 1) & 3] = patp[itb + 1];
        s_uq[2 * cb] = uq[itb];
        s_uq[2 * cb + 1] = uq[uqs + itb];
        s_lab[cb] = labp[itb];
      }
      for (int p = tid; p < P; p += NTH) XS[cb * Pp + p] = g_x[(size_t)patp[itb] * P + p];
      for (int j = tid; j < K + CC; j += NTH) s_lik[cb * NC + j] = NOLIK ? 0.0 : lik[(size_t)colmap(j, K) * ntp + itb];
    }
    bool resync = false;  // the candidates' likelihood entries of the NEXT step must be re-read (structural change)
    __syncthreads();

    for (int it = itb; it < nt; it++) {
      if (K >= KS) {  // a birth could overflow the staged capacity: hand the unit over (state is consistent)
        bail_l = l;
        bail_it = it;
        live = false;
        break;
      }
      const int cur = it & 1, nxt = cur ^ 1;
      const double *xs = XS + cur * Pp;
      double *likc = s_lik + cur * NC;
      // ---- asynchronous fetch of the operands of step it+1 into the other buffer
      {
        const int i1 = s_pat[(it + 1) & 3];
#pragma unroll
        for (int r = 0; r < XR; r++) {
          const int p = tid + r * NTH;
          if (p < P) cp_async8(XS + nxt * Pp + p, g_x + (size_t)i1 * P + p);
        }
        if (tid == 0) {
          cp_async8(s_uq + 2 * nxt, uq + it + 1);
          cp_async8(s_uq + 2 * nxt + 1, uq + uqs + it + 1);
          cp_async4(s_lab + nxt, labp + it + 1);
          if (it + 2 <= ntp) cp_async4(s_pat + ((it + 2) & 3), patp + it + 2);
        }
        if (!NOLIK)
          for (int j = tid; j < K + CC; j += NTH) cp_async8(s_lik + nxt * NC + j, lik + (size_t)colmap(j, K) * ntp + it + 1);
      }
      const double uu = s_uq[2 * cur], qx = s_uq[2 * cur + 1];
      double xv[XR];
#pragma unroll
      for (int r = 0; r < XR; r++) xv[r] = (tid + r * NTH < P) ? xs[tid + r * NTH] : 0.0;

      // ---- :368-457 take the patient out of its cluster
      const int zi = s_lab[cur] - 1;
      const int nzi = s_nj[zi];
      TPPMX_ASSERT(zi >= 0 && zi < K && nzi >= 1, 21);
      const bool single = nzi <= 1;
      bool structural = false;
      double vR1[XR], vR2[XR], vA1[XR], vA2[XR];  // exact statistics in flight (covariate owners)
      if (single) {  // singleton (:384-457), rare: everything synchronously on the exact statistics
        cp_async_wait_all();
        flush_pending();
        __syncthreads();
        const int iaux = zi + 1, Kc = K;
        if (iaux < Kc) {  // swap with the last cluster
          for (int ii = tid; ii < nt; ii += NTH)
            if (labp[ii] == Kc) labp[ii] = iaux;
          __syncthreads();
          if (tid == 0) {
            labp[it] = Kc;
            s_nj[iaux - 1] = s_nj[Kc - 1];
            s_nj[Kc - 1] = 1;
            const int c = s_col[iaux - 1];
            s_col[iaux - 1] = s_col[Kc - 1];
            s_col[Kc - 1] = c;
          }
          for (int k = tid; k < D; k += NTH) {
            const double t = g_eta[(size_t)k * kcap + iaux - 1];
            g_eta[(size_t)k * kcap + iaux - 1] = g_eta[(size_t)k * kcap + Kc - 1];
            g_eta[(size_t)k * kcap + Kc - 1] = t;
          }
#pragma unroll
          for (int r = 0; r < XR; r++) {
            const int p = tid + r * NTH;
            if (p < P) {
              const size_t e1 = (size_t)p * kcap + iaux - 1, e2 = (size_t)p * kcap + Kc - 1;
              double t = g_sumx[e1];
              g_sumx[e1] = g_sumx[e2];
              g_sumx[e2] = t;
              t = g_sumx2[e1];
              g_sumx2[e1] = g_sumx2[e2];
              g_sumx2[e2] = t;
              T[(iaux - 1) * pstr + p] = fma(g_sumx[e1], iv2, c0);
            }
          }
        }
        __syncthreads();
        if (tid == 0) {  // empty slot Kc-1 (keeps the rounding residue, as the reference does)
          s_nj[Kc - 1] -= 1;
          s_free[nfree] = s_col[Kc - 1];
          if (iaux < Kc) s_a[iaux - 1] = a_of(iaux - 1);
          s_lab[nxt] = labp[it + 1];  // a relabelling may have touched the next patient
        }
        nfree += 1;
#pragma unroll
        for (int r = 0; r < XR; r++) {
          const int p = tid + r * NTH;
          if (p < P) {
            const size_t e2 = (size_t)p * kcap + Kc - 1;
            g_sumx[e2] = __dsub_rn(g_sumx[e2], xv[r]);
            g_sumx2[e2] = __dsub_rn(g_sumx2[e2], __dmul_rn(xv[r], xv[r]));
          }
        }
        K = Kc - 1;
        dV = dV_of(K);
        structural = true;
        resync = true;
        __syncthreads();
        if (!NOLIK)  // the column map changed: this step's entries again, by the new cluster numbering
          for (int j = tid; j < K + CC; j += NTH) likc[j] = lik[(size_t)colmap(j, K) * ntp + it];
      } else {
#pragma unroll
        for (int r = 0; r < XR; r++) {
          const int p = tid + r * NTH;
          vR1[r] = vR2[r] = vA1[r] = vA2[r] = 0.0;
          if (p < P) {
            T[zi * pstr + p] = fma(-iv2, xv[r], T[zi * pstr + p]);
            if (pend_c >= 0) {
              const size_t eA = (size_t)p * kcap + pend_c;
              vA1[r] = g_sumx[eA];
              vA2[r] = g_sumx2[eA];
            }
            if (pend_c != zi) {
              const size_t eR = (size_t)p * kcap + zi;
              vR1[r] = g_sumx[eR];
              vR2[r] = g_sumx2[eR];
            }
          }
        }
      }
      __syncthreads();  // (A) removal visible

      // ---- :460-814 score the K occupied + CC auxiliary clusters, normalise, draw
      const int ntot = K + CC;
      const int nch = (ntot + NTH - 1) / NTH;
      int wimg = (int)0x80000000;  // running max of the float image of w
      for (int c = 0; c < nch; c++) {
        const int j = tid + c * NTH;
        if (c * NTH + warp * 32 >= ntot) break;  // warp-uniform: no candidate in this warp's slice
        const bool actj = j < ntot, occj = j < K;
        const int col = occj ? j : KS;
        double b0 = 0.0, b1 = 0.0;
        {  // b_j = sum_p t_jp x_p: 128-bit loads (row stride 2 x odd doubles: conflict free, 16-byte aligned),
           // eight covariates per trip with the loads issued ahead of the fused multiply-adds
          const double2 *tc = reinterpret_cast<const double2 *>(T + (size_t)col * pstr);
          const double2 *xc = reinterpret_cast<const double2 *>(xs);
          const int np2 = Pp >> 1;  // pairs; the pad entry of xs / t is zero
          int q = 0;
          for (; q + 4 <= np2; q += 4) {
            const double2 t0 = tc[q], t1 = tc[q + 1], t2 = tc[q + 2], t3 = tc[q + 3];
            const double2 x0 = xc[q], x1 = xc[q + 1], x2 = xc[q + 2], x3 = xc[q + 3];
            b0 = fma(t0.x, x0.x, b0);
            b1 = fma(t0.y, x0.y, b1);
            b0 = fma(t1.x, x1.x, b0);
            b1 = fma(t1.y, x1.y, b1);
            b0 = fma(t2.x, x2.x, b0);
            b1 = fma(t2.y, x2.y, b1);
            b0 = fma(t3.x, x3.x, b0);
            b1 = fma(t3.y, x3.y, b1);
          }
          for (; q < np2; q++) {
            const double2 t0 = tc[q], x0 = xc[q];
            b0 = fma(t0.x, x0.x, b0);
            b1 = fma(t0.y, x0.y, b1);
          }
        }
        const double incr = iv2 * (2.0 * (b0 + b1) + iv2 * qx);
        int n = occj ? s_nj[j] : 0;
        double a = s_a[col];
        if (!structural && j == zi) {  // the cluster the patient just left: its size and a_j follow the removal
          n -= 1;
          a -= incr;
          s_a[j] = a;
        }
        const double tY = a + incr;
        const double2 t01 = tab2[2 * n], t23 = tab2[2 * n + 1];  // (dA, HN), (HY, -)
        const double d = t01.x - hiv2 * qx + (t23.x * tY - t01.y * a);
        const double w = actj ? dV + lognp[n] + simscale * d + (NOLIK ? 0.0 : likc[actj ? j : 0]) : -INFINITY;
        if (actj) {
          s_w[j] = w;
          s_tY[j] = tY;
        }
        int i = __float_as_int((float)w);
        i ^= (i >> 31) & 0x7fffffff;
        wimg = max(wimg, i);
      }
      wimg = __reduce_max_sync(TPPMX_FULL, wimg);
      float wmaxf;
      if (NW > 1) {
        if (lane == 0) s_max[warp] = __int_as_float(wimg ^ ((wimg >> 31) & 0x7fffffff));
        __syncthreads();  // (B)
        wmaxf = s_max[0];
#pragma unroll
        for (int r = 1; r < NW; r++) wmaxf = fmaxf(wmaxf, s_max[r]);
      } else {
        wmaxf = __int_as_float(wimg ^ ((wimg >> 31) & 0x7fffffff));
      }
      // e = exp(w - max), inclusive sums inside each (chunk, warp) slice; slice totals for the cross-slice prefix
      for (int c = 0; c < nch; c++) {
        const int j = tid + c * NTH;
        const int o = c * NW + warp;
        if (c * NTH + warp * 32 >= ntot) {
          if (lane == 0 && o < NO) s_tot[o] = 0.0;
          continue;
        }
        const bool actj = j < ntot;
        const double e0 = actj ? fast_exp(s_w[j] - (double)wmaxf) : 0.0;
        if constexpr (PROBE) {
          const int pu = unit - fa.probe_unit0;
          if (pu >= 0 && pu < fa.probe_nunits && actj && j < fa.probe_stride)
            fa.probe[((size_t)pu * ntp + it) * fa.probe_stride + j] = e0;  // normalised below, once den is known
        }
        double cum = e0;
#pragma unroll
        for (int sh = 1; sh < 32; sh <<= 1) {
          const double t = __shfl_up_sync(TPPMX_FULL, cum, sh);
          if (lane >= sh) cum += t;
        }
        if (actj) s_w[j] = cum;
        if (lane == 31) s_tot[o] = cum;
      }
      __syncthreads();  // (C)
      const int nslice = nch * NW;
      double den = 0.0;
      for (int o = 0; o < nslice; o++) den += s_tot[o];
      const double thr = uu * den;  // uu < cumsum(e)/den  <=>  uu*den < cumsum(e)   (:805-814)
      int myhit = 0x7fffffff;
      {
        double pre = 0.0;
        int o0 = 0;
        for (int c = 0; c < nch; c++) {
          const int j = tid + c * NTH;
          const int o = c * NW + warp;
          for (; o0 < o; o0++) pre += s_tot[o0];
          if (c * NTH + warp * 32 >= ntot) break;
          const bool actj = j < ntot;
          const unsigned hit = __ballot_sync(TPPMX_FULL, actj && (thr < pre + s_w[actj ? j : 0]));
          if (hit && myhit == 0x7fffffff) myhit = c * NTH + warp * 32 + __ffs(hit);
        }
      }
      int newci;
      if (NW > 1) {
        if (lane == 0) s_hit[warp] = myhit;
        __syncthreads();  // (D)
        newci = s_hit[0];
#pragma unroll
        for (int r = 1; r < NW; r++) newci = min(newci, s_hit[r]);
      } else {
        newci = myhit;
      }
      if (newci == 0x7fffffff) newci = ntot;  // :807 default when rounding leaves uu >= total
      TPPMX_ASSERT(newci >= 1 && newci <= ntot, 22);
      if constexpr (PROBE) {
        const int pu = unit - fa.probe_unit0;
        if (pu >= 0 && pu < fa.probe_nunits) {
          for (int j = tid; j < ntot && j < fa.probe_stride; j += NTH) fa.probe[((size_t)pu * ntp + it) * fa.probe_stride + j] /= den;
          if (tid == 0) fa.probe_k[(size_t)pu * ntp + it] = K;
        }
      }

      // ---- :820-856 assign, update sufficient statistics
      const bool born = newci > K;
      int lab = newci;
      if (!born) {
#pragma unroll
        for (int r = 0; r < XR; r++) {
          const int p = tid + r * NTH;
          if (p < P) {
            T[(lab - 1) * pstr + p] = fma(iv2, xv[r], T[(lab - 1) * pstr + p]);
            if (!single) {  // retire the reads issued at the top of the step
              const size_t eR = (size_t)p * kcap + zi;
              if (pend_c >= 0) {
                const size_t eA = (size_t)p * kcap + pend_c;
                double s1 = __dadd_rn(vA1[r], pend_x[r]), s2 = __dadd_rn(vA2[r], __dmul_rn(pend_x[r], pend_x[r]));
                if (pend_c == zi) {
                  s1 = __dsub_rn(s1, xv[r]);
                  s2 = __dsub_rn(s2, __dmul_rn(xv[r], xv[r]));
                } else {
                  g_sumx[eR] = __dsub_rn(vR1[r], xv[r]);
                  g_sumx2[eR] = __dsub_rn(vR2[r], __dmul_rn(xv[r], xv[r]));
                }
                g_sumx[eA] = s1;
                g_sumx2[eA] = s2;
              } else {
                g_sumx[eR] = __dsub_rn(vR1[r], xv[r]);
                g_sumx2[eR] = __dsub_rn(vR2[r], __dmul_rn(xv[r], xv[r]));
              }
            }
          }
        }
        pend_c = lab - 1;
#pragma unroll
        for (int r = 0; r < XR; r++) pend_x[r] = xv[r];
        if (tid == 0) {
          labp[it] = lab;
          if (!single) s_nj[zi] = nzi - 1;
          s_nj[lab - 1] += 1;
          s_a[lab - 1] = s_tY[lab - 1];
        }
      } else {  // a new cluster takes the auxiliary eta (:829-845), rare
        cp_async_wait_all();
        if (!single) {  // the removal of this step is still in flight: retire it first
#pragma unroll
          for (int r = 0; r < XR; r++) {
            const int p = tid + r * NTH;
            if (p < P) {
              const size_t eR = (size_t)p * kcap + zi;
              if (pend_c == zi) {
                g_sumx[eR] = __dsub_rn(__dadd_rn(vA1[r], pend_x[r]), xv[r]);
                g_sumx2[eR] = __dsub_rn(__dadd_rn(vA2[r], __dmul_rn(pend_x[r], pend_x[r])), __dmul_rn(xv[r], xv[r]));
              } else {
                g_sumx[eR] = __dsub_rn(vR1[r], xv[r]);
                g_sumx2[eR] = __dsub_rn(vR2[r], __dmul_rn(xv[r], xv[r]));
              }
            }
          }
          if (pend_c == zi) pend_c = -1;
        }
        flush_pending();
        const int i = s_pat[it & 3], m = newci - K - 1;
        lab = K + 1;
        TPPMX_ASSERT(nfree >= 1 && lab <= KS, 23);
        const int newcol = s_free[nfree - 1];
        nfree -= 1;
        __syncthreads();
        if (tid == 0) {
          labp[it] = lab;
          if (!single) s_nj[zi] = nzi - 1;
          s_nj[lab - 1] = 1;
          s_col[lab - 1] = s_col[KS + m];
          s_col[KS + m] = newcol;
        }
        for (int k = tid; k < D; k += NTH) g_eta[(size_t)k * kcap + lab - 1] = g_etae[k * CC + m];
#pragma unroll
        for (int r = 0; r < XR; r++) {
          const int p = tid + r * NTH;
          if (p < P) {
            const size_t e = (size_t)p * kcap + lab - 1;
            const double v1 = __dadd_rn(g_sumx[e], xv[r]);
            g_sumx[e] = v1;
            g_sumx2[e] = __dadd_rn(g_sumx2[e], __dmul_rn(xv[r], xv[r]));
            T[(lab - 1) * pstr + p] = fma(v1, iv2, c0);
          }
        }
        __syncthreads();
        if (tid == 0) {
          s_a[lab - 1] = a_of(lab - 1);
          rng_normals(rng, l, TPPMX_SITE_AUX_REFRESH, tt, (uint32_t)i, D, g_etae + m, CC);
        }
        __syncthreads();
        if (!NOLIK)
          for (int k = tid; k < D; k += NTH) E[newcol * D + k] = fast_exp(g_etae[k * CC + m]);
        __syncthreads();
        if (!NOLIK)  // likelihood column of the refreshed auxiliary eta, remaining patients
          for (int it2 = it + 1 + tid; it2 < nt; it2 += NTH) {
            const int i2 = patp[it2];
            lik[(size_t)newcol * ntp + it2] = fast_lik<DR>(E + newcol * D, g_ez + (size_t)i2 * D, g_ljj + (size_t)i2 * D, D);
          }
        K = lab;
        if (K > Kpeak) Kpeak = K;
        dV = dV_of(K);
        resync = true;
      }

      // ---- the operands of the next step have landed
      cp_async_wait_all();
      __syncthreads();  // (E)
      if (resync) {  // the column map changed after the asynchronous fetch was issued: read the next entries again
        if (!NOLIK)
          for (int j = tid; j < K + CC; j += NTH) s_lik[nxt * NC + j] = lik[(size_t)colmap(j, K) * ntp + it + 1];
        resync = false;
        __syncthreads();
      }
    }
    // ---- end of the sweep: the exact statistics are complete in global memory
    cp_async_wait_all();
    flush_pending();
    __syncthreads();
  }

  // ---- write the unit back in the library's global layout
  for (int j = tid; j < KS; j += NTH) g_nj[j] = s_nj[j];
  if (tid == 0) {
    b.nclu[(size_t)chain * b.T + tt] = K;
    fa.resume_l[unit] = bail_l;
    fa.resume_it[unit] = bail_it;
    if (bail_l >= 0) atomicAdd(fa.bail_count, 1);
    atomicMax(fa.kmax_dev, Kpeak);
  }
  __syncthreads();
  // loggamma (:825, :839), as sweep_fast.cuh: eta* is constant during the sweep
  if (n_sweeps > 0) {
    for (int it = tid; it < nt; it += NTH) {
      const int i = patp[it];
      const int lb = labp[it];
      const double *zi = b.z + ((size_t)ds * b.nobs + i) * Q;
      for (int k = 0; k < D; k++) {
        double lg = g_eta[(size_t)k * kcap + lb - 1];
        for (int q = 0; q < Q; q++) lg = __dadd_rn(lg, __dmul_rn(beta[k + q * D], zi[q]));
        b.lg[((size_t)chain * b.nobs + i) * D + k] = lg;
      }
    }
  }
}

// ---- launch -----------------------------------------------------------------------------------------
template <int NTH, int XR, int DR, bool NOLIK>
static cudaError_t big_launch3(const DevBatch &d, const FastArgs &fa, size_t smem, uint64_t seed, int iter0, int n_sweeps,
                               cudaStream_t stream) {
  if (fa.probe) {
    if constexpr (DR == 3 && !NOLIK) {
      cudaError_t e = cudaFuncSetAttribute(sweep_big_kernel<NTH, XR, DR, NOLIK, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
      sweep_big_kernel<NTH, XR, DR, NOLIK, true><<<fa.n_units, NTH, smem, stream>>>(d, fa, seed, iter0, n_sweeps);
      return cudaGetLastError();
    }
    return cudaErrorInvalidConfiguration;
  }
  cudaError_t e = cudaFuncSetAttribute(sweep_big_kernel<NTH, XR, DR, NOLIK, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  sweep_big_kernel<NTH, XR, DR, NOLIK, false><<<fa.n_units, NTH, smem, stream>>>(d, fa, seed, iter0, n_sweeps);
  return cudaGetLastError();
}
template <int NTH, int XR>
static cudaError_t big_launch_impl(const DevBatch &d, const FastArgs &fa, size_t smem, uint64_t seed, int iter0, int n_sweeps,
                                   cudaStream_t stream) {
  const bool nolik = d.model.nolik != 0;
  if (d.D <= 3)
    return nolik ? big_launch3<NTH, XR, 3, true>(d, fa, smem, seed, iter0, n_sweeps, stream)
                 : big_launch3<NTH, XR, 3, false>(d, fa, smem, seed, iter0, n_sweeps, stream);
  return nolik ? big_launch3<NTH, XR, TPPMX_MAXD, true>(d, fa, smem, seed, iter0, n_sweeps, stream)
               : big_launch3<NTH, XR, TPPMX_MAXD, false>(d, fa, smem, seed, iter0, n_sweeps, stream);
}

}  // namespace tppmx
