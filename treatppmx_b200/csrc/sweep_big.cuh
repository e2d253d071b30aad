// sweep_big.cuh — the reallocation-sweep kernel for LONG arms (BASELINE config 4: n_t in the
// thousands, P = 50 covariates, K_t from tens to hundreds of clusters): reference
// src/dm_ppmx_ct.cpp:347-860 with consim == 1, similarity == 1, the same switches and the same
// arithmetic per score as sweep_fast.cuh, laid out for a different regime.
//
// With thousands of steps per sweep and K * P in the thousands, a step is a matrix-vector product
// (clusters x covariates) followed by a softmax draw: throughput of on-chip operand reads and the
// latency of one step both matter, and the cluster statistics of one (chain, arm) do not fit next to
// eleven other units in an SM's shared memory.  So:
//  * unit = one (chain, arm) = one CTA of NTH threads (128, or 256 for more than 128 staged clusters):
//    thread j OWNS cluster j -- its size, the per-size constants, its cached a_j = sum_p t_jp^2 live
//    in registers -- and warps combine max / sum / first-hit through shared memory (four barriers a
//    step).  All (K + CC) candidates are scored in one pass: no 32-cluster cap, no chunk loop.
//  * shared memory holds ONE array: t_jp = sumx_jp / v2 + m0 / s20, [P][KS | 1] (odd stride: column
//    reads by cluster owners and row updates by covariate owners are both conflict free).  The score
//    needs only b_j = sum_p t_jp x_p (one LDS + one DFMA per covariate); a_j is carried
//    incrementally (removal: a -= iv2 (2 b + iv2 q), arrival: a = t_Y of the draw's score) and t by
//    rank-one updates; both are REBUILT from the exact statistics at every sweep start, so the
//    rounding drift is bounded by one sweep (<= 1e-12 on a log-weight at n_t = 6667, tested <= 1e-10
//    on every probability by the probe instantiation).
//  * the exact sufficient statistics sumx / sumx2 stay in global memory (L2) and are updated by the
//    covariate owners with the reference's own operation sequence (bit-identical to the oracle);
//    their read-modify-writes are issued at the top of a step and retired at its end, one step
//    deferred for the arrival, so no L2 round trip sits on the step's dependent chain.
//  * likelihood table, uniforms, patient order, labels: as sweep_fast.cuh's big-arm mode (global,
//    written by a 4x wider parallel pre-pass, prefetched one step ahead).
//  * a unit that reaches the staged capacity KS (<= 256) hands over to the generic kernel at the exact
//    (sweep, patient) position, like the small-arm kernel.
#pragma once
#include "sweep_fast.cuh"

namespace tppmx {

template <int NTH, int DR, bool NOLIK, bool PROBE>
__global__ void __launch_bounds__(NTH, NTH == 128 ? 4 : 2)   // 128 registers per thread: 16 warps per SM
sweep_big_kernel(DevBatch b, FastArgs fa, uint64_t seed, int iter0, int n_sweeps) {
  constexpr int NW = NTH / 32;
  extern __shared__ __align__(16) char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int unit = blockIdx.x;
  const int P = b.P, D = b.D, Q = b.Q, CC = b.CC, KS = fa.KS, ntp = fa.ntp, kcap = b.kcap;
  const int kstr = big_kstr(KS), Pp = fast_pp(P);
  const SimConst s = make_simconst(b);

  // ---- shared memory
  double *T = reinterpret_cast<double *>(smem_raw);        // [P][kstr], column KS = the empty cluster (t = c0)
  double *E = T + (size_t)P * kstr;                        // [KS + CC][D] exp(eta) by likelihood column
  double *XS = E + (size_t)(KS + CC) * D;                  // [2][Pp] covariate row of the current / next patient
  double *s_tot = XS + 2 * Pp;                             // [NW] per-warp sum of e
  float *s_max = reinterpret_cast<float *>(s_tot + NW);    // [NW] per-warp max of w (float image), padded to NW doubles
  int *s_nj = reinterpret_cast<int *>(s_tot + 2 * NW);     // [KS]
  int *s_col = s_nj + KS;                                  // [KS + CC] cluster / auxiliary -> likelihood column
  int *s_free = s_col + KS + CC;                           // [KS + CC] free likelihood columns
  int *s_hit = s_free + KS + CC;                           // [NW] first hit of each warp
  int *s_misc = s_hit + NW;                                // [0] = size of the cluster the next patient sits in

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
  double *lik = fa.lik + (size_t)unit * fa.ncol * ntp;
  const int uqs = ntp + 1;
  double *uq = fa.uq + (size_t)unit * 2 * uqs;
  double *g_ez = b.zb + (size_t)chain * b.nobs * D;    // exp(beta'z) (scratch, rebuilt per sweep)
  double *g_ljj = b.ljj + (size_t)chain * b.nobs * D;
  const double *beta = b.beta + (size_t)chain * Q * D;
  int *patp = fa.patpad + (size_t)unit * (ntp + 1);
  double *lognp = fa.logng + (size_t)unit * (ntp + 1);
  const double2 *tab2 = reinterpret_cast<const double2 *>(fa.tabg);

  int K = b.nclu[(size_t)chain * b.T + tt];
  const double sigma = b.sigma[(size_t)chain * b.T + tt];
  const int idxs = b.idxs[chain];
  const double iv2 = s.iv2, c0 = s.c0, hiv2 = s.hiv2;
  const double simscale = (s.calibration == 2) ? s.calscale : 1.0;

  // ---- stage what lives on chip for the whole launch
  for (int j = tid; j < KS; j += NTH) s_nj[j] = g_nj[j];
  for (int it = tid; it <= ntp; it += NTH) patp[it] = (it < nt) ? g_pat[it] : 0;
  int nfree;
  {
    const int Kc = K < KS ? K : KS;
    for (int j = tid; j < Kc; j += NTH) s_col[j] = j;
    for (int m = tid; m < CC; m += NTH) s_col[KS + m] = Kc + m;
    nfree = fa.ncol - Kc - CC;
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

  int bail_l = -1, bail_it = 0;
  bool live = true;  // CTA-uniform

  // exact statistics in global memory: the read-modify-write of an arrival is retired one step late
  int pend_c = -1;        // cluster whose sumx / sumx2 still miss the patient `pend_x` (covariate owners only)
  double pend_x = 0.0;
  auto flush_pending = [&]() {
    if (tid < P && pend_c >= 0) {
      const size_t e = (size_t)tid * kcap + pend_c;
      g_sumx[e] = __dadd_rn(g_sumx[e], pend_x);
      g_sumx2[e] = __dadd_rn(g_sumx2[e], __dmul_rn(pend_x, pend_x));
    }
    pend_c = -1;
  };

  for (int l = iter0; l < iter0 + n_sweeps && live; l++) {
    if (K > KS) {
      bail_l = l;
      bail_it = 0;
      live = false;
      break;
    }
    // =========================== parallel pre-pass of the sweep ===========================
    for (int mm = tid; mm < CC; mm += NTH)  // :351-360 auxiliary eta ~ N(0, I)
      rng_normals(rng, l, TPPMX_SITE_AUX_SWEEP, tt, (uint32_t)mm, D, g_etae + mm, CC);
    // t = sumx / v2 + m0 / s20 from the exact statistics (drift of the rank-one updates ends here)
    for (int e = tid; e < P * kstr; e += NTH) {
      const int p = e / kstr, j = e - p * kstr;
      T[e] = fma((j < K) ? g_sumx[(size_t)p * kcap + j] : 0.0, iv2, c0);
    }
    __syncthreads();
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
    // per-thread view of "my" cluster (rebuilt after every structural change)
    bool act = false, occ = false;
    int jc = KS, n_j = 0;
    double a = 0.0, dA = 0.0, HN = 0.0, HY = 0.0, lognj = 0.0;
    const double *likp = lik;
    auto load_n = [&](int n) {
      const double2 t01 = tab2[2 * n], t23 = tab2[2 * n + 1];  // (dA, HN), (HY, -)
      dA = t01.x;
      HN = t01.y;
      HY = t23.x;
      lognj = lognp[n];
    };
    auto refresh = [&]() {  // after a barrier that follows the structural change
      act = tid < K + CC;
      occ = tid < K;
      jc = occ ? tid : KS;
      n_j = occ ? s_nj[tid] : 0;
      load_n(n_j);
      double a0 = 0.0, a1 = 0.0;
      int p = 0;
      for (; p + 1 < P; p += 2) {
        const double t0 = T[p * kstr + jc], t1 = T[(p + 1) * kstr + jc];
        a0 = fma(t0, t0, a0);
        a1 = fma(t1, t1, a1);
      }
      if (p < P) a0 = fma(T[p * kstr + jc], T[p * kstr + jc], a0);
      a = a0 + a1;
      likp = lik + (size_t)(act ? colmap(tid, K) : 0) * ntp;
    };
    refresh();
    double dV = dV_of(K);
    int labcur = labp[0];
    double likv = NOLIK ? 0.0 : likp[0], uu = uq[0], qx = uq[uqs];
    if (tid < P) XS[tid] = g_x[(size_t)patp[0] * P + tid];
    if (tid == 0) s_misc[0] = s_nj[labcur - 1];
    bool structural = false;
    __syncthreads();

    for (int it = 0; it < nt; it++) {
      if (K >= KS) {  // a birth could overflow the staged capacity: hand the unit over (state is consistent)
        bail_l = l;
        bail_it = it;
        live = false;
        break;
      }
      const double *xs = XS + (it & 1) * Pp;
      // ---- prefetch the operands of step it+1
      double xr_n = 0.0, lik_n = 0.0;
      if (tid < P) xr_n = g_x[(size_t)patp[it + 1] * P + tid];
      const double uu_n = uq[it + 1], qx_n = uq[uqs + it + 1];
      int lab_n = labp[it + 1];
      if (!NOLIK) lik_n = likp[it + 1];
      const double xv = (tid < P) ? xs[tid] : 0.0;

      // ---- :368-457 take the patient out of its cluster
      const int zi = labcur - 1;
      const int nzi = s_misc[0];
      TPPMX_ASSERT(zi >= 0 && zi < K && nzi >= 1, 21);
      const bool single = nzi <= 1;
      double vR1 = 0.0, vR2 = 0.0, vA1 = 0.0, vA2 = 0.0;  // exact statistics in flight (covariate owners)
      if (single) {  // singleton (:384-457), rare: everything synchronously on the exact statistics
        flush_pending();
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
          if (tid < P) {
            const size_t e1 = (size_t)tid * kcap + iaux - 1, e2 = (size_t)tid * kcap + Kc - 1;
            double t = g_sumx[e1];
            g_sumx[e1] = g_sumx[e2];
            g_sumx[e2] = t;
            t = g_sumx2[e1];
            g_sumx2[e1] = g_sumx2[e2];
            g_sumx2[e2] = t;
            T[tid * kstr + iaux - 1] = fma(g_sumx[e1], iv2, c0);
          }
        }
        __syncthreads();
        if (tid == 0) {  // empty slot Kc-1 (keeps the rounding residue, as the reference does)
          s_nj[Kc - 1] -= 1;
          s_free[nfree] = s_col[Kc - 1];
        }
        nfree += 1;
        if (tid < P) {
          const size_t e2 = (size_t)tid * kcap + Kc - 1;
          g_sumx[e2] = __dsub_rn(g_sumx[e2], xv);
          g_sumx2[e2] = __dsub_rn(g_sumx2[e2], __dmul_rn(xv, xv));
        }
        K = Kc - 1;
        dV = dV_of(K);
        structural = true;
      } else {
        if (tid == 0) s_nj[zi] = nzi - 1;
        if (tid < P) {
          T[tid * kstr + zi] = fma(-iv2, xv, T[tid * kstr + zi]);
          const size_t eR = (size_t)tid * kcap + zi;
          if (pend_c >= 0) {
            const size_t eA = (size_t)tid * kcap + pend_c;
            vA1 = g_sumx[eA];
            vA2 = g_sumx2[eA];
          }
          if (pend_c != zi) {
            vR1 = g_sumx[eR];
            vR2 = g_sumx2[eR];
          }
        }
      }
      __syncthreads();  // (A) removal visible
      if (structural) {
        refresh();
        if (!NOLIK) {
          likv = likp[it];
          lik_n = likp[it + 1];
        }
        lab_n = labp[it + 1];  // a relabelling may have touched it
      } else if (tid == zi) {
        n_j -= 1;
        load_n(n_j);
      }

      // ---- :460-814 score the K occupied + CC auxiliary clusters, normalise, draw
      const int ntot = K + CC;
      double b0 = 0.0, b1 = 0.0;
      if (warp * 32 < ntot) {  // warp-uniform: warps that own no candidate skip the dot product
        const double *tc = T + jc;
        int p = 0;
        for (; p + 1 < P; p += 2) {
          b0 = fma(tc[p * kstr], xs[p], b0);
          b1 = fma(tc[(p + 1) * kstr], xs[p + 1], b1);
        }
        if (p < P) b0 = fma(tc[p * kstr], xs[p], b0);
      }
      const double bb = b0 + b1;
      const double incr = iv2 * (2.0 * bb + iv2 * qx);
      if (!structural && !single && tid == zi) a -= incr;  // a of the cluster the patient just left
      const double tY = a + incr;
      const double d = dA - hiv2 * qx + (HY * tY - HN * a);
      const double w = act ? dV + lognj + simscale * d + likv : -INFINITY;
      {
        int i = __float_as_int((float)w);
        i ^= (i >> 31) & 0x7fffffff;
        i = __reduce_max_sync(TPPMX_FULL, i);
        i ^= (i >> 31) & 0x7fffffff;
        if (lane == 0) s_max[warp] = __int_as_float(i);
      }
      __syncthreads();  // (B)
      float wmaxf = s_max[0];
#pragma unroll
      for (int r = 1; r < NW; r++) wmaxf = fmaxf(wmaxf, s_max[r]);
      const double e0 = act ? fast_exp(w - (double)wmaxf) : 0.0;
      double cum = e0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const double t = __shfl_up_sync(TPPMX_FULL, cum, o);
        if (lane >= o) cum += t;
      }
      if (lane == 31) s_tot[warp] = cum;
      __syncthreads();  // (C)
      double den = 0.0, pre = 0.0;
#pragma unroll
      for (int r = 0; r < NW; r++) {
        const double t = s_tot[r];
        if (r < warp) pre += t;
        den += t;
      }
      cum += pre;
      {
        const unsigned hit = __ballot_sync(TPPMX_FULL, act && (uu * den < cum));
        if (lane == 0) s_hit[warp] = hit ? warp * 32 + __ffs(hit) : 0x7fffffff;
      }
      if constexpr (PROBE) {
        const int pu = unit - fa.probe_unit0;
        if (pu >= 0 && pu < fa.probe_nunits) {
          if (act && tid < fa.probe_stride) fa.probe[((size_t)pu * ntp + it) * fa.probe_stride + tid] = e0 / den;
          if (tid == 0) fa.probe_k[(size_t)pu * ntp + it] = K;
        }
      }
      __syncthreads();  // (D)
      int newci = s_hit[0];
#pragma unroll
      for (int r = 1; r < NW; r++) newci = min(newci, s_hit[r]);
      if (newci == 0x7fffffff) newci = ntot;  // :807 default when rounding leaves uu >= total
      TPPMX_ASSERT(newci >= 1 && newci <= ntot, 22);

      // ---- :820-856 assign, update sufficient statistics
      const bool born = newci > K;
      int lab = newci;
      if (!born) {
        if (tid == 0) {
          labp[it] = lab;
          s_nj[lab - 1] += 1;
        }
        if (tid < P) {
          T[tid * kstr + lab - 1] = fma(iv2, xv, T[tid * kstr + lab - 1]);
          if (!single) {  // retire the reads issued at the top of the step
            const size_t eR = (size_t)tid * kcap + zi;
            if (pend_c >= 0) {
              const size_t eA = (size_t)tid * kcap + pend_c;
              double s1 = __dadd_rn(vA1, pend_x), s2 = __dadd_rn(vA2, __dmul_rn(pend_x, pend_x));
              if (pend_c == zi) {
                s1 = __dsub_rn(s1, xv);
                s2 = __dsub_rn(s2, __dmul_rn(xv, xv));
              } else {
                g_sumx[eR] = __dsub_rn(vR1, xv);
                g_sumx2[eR] = __dsub_rn(vR2, __dmul_rn(xv, xv));
              }
              g_sumx[eA] = s1;
              g_sumx2[eA] = s2;
            } else {
              g_sumx[eR] = __dsub_rn(vR1, xv);
              g_sumx2[eR] = __dsub_rn(vR2, __dmul_rn(xv, xv));
            }
          }
        }
        pend_c = lab - 1;
        pend_x = xv;
        if (tid == lab - 1) {  // (after a singleton removal `a` was rebuilt post-removal: t_Y is still its arrival value)
          a = tY;
          n_j += 1;
          load_n(n_j);
        }
      } else {  // a new cluster takes the auxiliary eta (:829-845), rare
        if (!single && tid < P) {  // the removal of this step is still in flight: retire it first
          const size_t eR = (size_t)tid * kcap + zi;
          if (pend_c >= 0 && pend_c == zi) {
            g_sumx[eR] = __dsub_rn(__dadd_rn(vA1, pend_x), xv);
            g_sumx2[eR] = __dsub_rn(__dadd_rn(vA2, __dmul_rn(pend_x, pend_x)), __dmul_rn(xv, xv));
            pend_c = -1;
          } else {
            g_sumx[eR] = __dsub_rn(vR1, xv);
            g_sumx2[eR] = __dsub_rn(vR2, __dmul_rn(xv, xv));
          }
        }
        if (!single && tid >= P && pend_c == zi) pend_c = -1;  // keep the replicated bookkeeping in step
        flush_pending();
        const int i = patp[it], m = newci - K - 1;
        lab = K + 1;
        TPPMX_ASSERT(nfree >= 1 && lab <= KS, 23);
        const int newcol = s_free[nfree - 1];
        nfree -= 1;
        __syncthreads();
        if (tid == 0) {
          labp[it] = lab;
          s_nj[lab - 1] = 1;
          s_col[lab - 1] = s_col[KS + m];
          s_col[KS + m] = newcol;
        }
        for (int k = tid; k < D; k += NTH) g_eta[(size_t)k * kcap + lab - 1] = g_etae[k * CC + m];
        if (tid < P) {
          const size_t e = (size_t)tid * kcap + lab - 1;
          const double v1 = __dadd_rn(g_sumx[e], xv);
          g_sumx[e] = v1;
          g_sumx2[e] = __dadd_rn(g_sumx2[e], __dmul_rn(xv, xv));
          T[tid * kstr + lab - 1] = fma(v1, iv2, c0);
        }
        __syncthreads();
        if (tid == 0) rng_normals(rng, l, TPPMX_SITE_AUX_REFRESH, tt, (uint32_t)i, D, g_etae + m, CC);
        __syncthreads();
        if (!NOLIK)
          for (int k = tid; k < D; k += NTH) E[newcol * D + k] = fast_exp(g_etae[k * CC + m]);
        __syncthreads();
        if (!NOLIK)  // likelihood column of the refreshed auxiliary eta, remaining patients
          for (int it2 = it + 1 + tid; it2 < nt; it2 += NTH) {
            const int i2 = patp[it2];
            lik[(size_t)newcol * ntp + it2] = fast_lik(E + newcol * D, g_ez + (size_t)i2 * D, g_ljj + (size_t)i2 * D, D);
          }
        K = lab;
        dV = dV_of(K);
        structural = true;
      }

      // ---- stage the operands of the next step
      if (tid < P) XS[((it + 1) & 1) * Pp + tid] = xr_n;
      uu = uu_n;
      qx = qx_n;
      likv = lik_n;
      labcur = lab_n;
      if (tid == 0 && it + 1 < nt) s_misc[0] = s_nj[lab_n - 1];
      __syncthreads();  // (E)
      if (structural) {
        if (born) {  // the view of "my" cluster changed again: rebuild for the next step
          refresh();
          if (!NOLIK) likv = likp[it + 1];
        }
        structural = false;
      }
    }
    // ---- end of the sweep: the exact statistics are complete in global memory
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
template <int NTH, int DR, bool NOLIK>
static cudaError_t big_launch3(const DevBatch &d, const FastArgs &fa, size_t smem, uint64_t seed, int iter0, int n_sweeps,
                               cudaStream_t stream) {
  if (fa.probe) {
    if constexpr (DR == 3 && !NOLIK) {
      cudaError_t e = cudaFuncSetAttribute(sweep_big_kernel<NTH, DR, NOLIK, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
      sweep_big_kernel<NTH, DR, NOLIK, true><<<fa.n_units, NTH, smem, stream>>>(d, fa, seed, iter0, n_sweeps);
      return cudaGetLastError();
    }
    return cudaErrorInvalidConfiguration;
  }
  cudaError_t e = cudaFuncSetAttribute(sweep_big_kernel<NTH, DR, NOLIK, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  sweep_big_kernel<NTH, DR, NOLIK, false><<<fa.n_units, NTH, smem, stream>>>(d, fa, seed, iter0, n_sweeps);
  return cudaGetLastError();
}
template <int NTH>
static cudaError_t big_launch_impl(const DevBatch &d, const FastArgs &fa, size_t smem, uint64_t seed, int iter0, int n_sweeps,
                                   cudaStream_t stream) {
  const bool nolik = d.model.nolik != 0;
  if (d.D <= 3)
    return nolik ? big_launch3<NTH, 3, true>(d, fa, smem, seed, iter0, n_sweeps, stream)
                 : big_launch3<NTH, 3, false>(d, fa, smem, seed, iter0, n_sweeps, stream);
  return nolik ? big_launch3<NTH, TPPMX_MAXD, true>(d, fa, smem, seed, iter0, n_sweeps, stream)
               : big_launch3<NTH, TPPMX_MAXD, false>(d, fa, smem, seed, iter0, n_sweeps, stream);
}

}  // namespace tppmx
