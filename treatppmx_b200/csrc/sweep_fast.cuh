// sweep_fast.cuh — the production reallocation-sweep kernel for the default model switches
// (PPMx=1, Normal-Normal auxiliary similarity, calibration 0|2, no categorical covariates,
// reuse=1): reference src/dm_ppmx_ct.cpp:347-860 with consim==1, similarity==1.
//
// Design (B200-first):
//  * unit = one (chain, arm).  Arms never touch each other's state inside the sweep, so the
//    n_chains*T units are independent sequential chains of n_t reallocation steps.
//  * LPU lanes per unit (16 by default => two units per warp, one warp per CTA): lane j
//    scores cluster j; max / sum / prefix-scan over clusters are width-LPU shuffles.
//  * the sequential step contains NO transcendental except one exp per lane (the softmax):
//      - Normal-Normal similarity difference in closed form from two dot products
//          a = sum_p t_p^2, b = sum_p t_p x_p,   t_p = sumx_jp / v2 + m0 / s20
//        (log terms depend on n_j only -> per-CTA shared-memory table),
//      - cohesion log(n_j - sigma) from a per-unit table over n,
//      - the Dirichlet-multinomial likelihood term L_i(eta_j) (3 exp + 3 lgamma per cluster)
//        is hoisted out of the sequential loop: eta* is constant during a sweep, so
//        Lik[col][it] is filled for all patients x occupied clusters in a fully parallel
//        pre-pass and only a cluster born from a refreshed auxiliary eta needs a (parallel)
//        column fill inside the loop.  Columns follow cluster identities through label swaps
//        via a small indirection table.  gamma = exp(eta + zb) is formed as exp(eta)*exp(zb)
//        and lgamma is a branch-free Stirling evaluation (no divergence across lanes).
//      - uniforms (Philox) and sum_p x_ip^2 are produced in the same pre-pass.
//  * cluster sufficient statistics (sumx, sumx2 [P][33] SoA, n_j, labels) live in shared
//    memory for the whole launch; the patient row x_i, Lik entry and uniform of the next step
//    are prefetched one step ahead, so no global-memory latency sits on the critical path.
//  * the hot loop is kept small (rare paths behind unlikely branches, one lgamma call site):
//    with one warp per CTA every warp is at a different PC, so instruction-cache footprint is
//    a first-order cost (profiles/r1_v1_*).
//  * a unit that would need more than KS (<=32) clusters writes its state back and is
//    finished by the generic kernel from the exact (sweep, patient) position (counter-based
//    RNG makes the hand-over exact).
#pragma once
#include "layout.h"
#include "rng.cuh"
#include "sim.cuh"

namespace tppmx {

#define TPPMX_FAST_KSMAX 32
#define TPPMX_FAST_KSTR 33  // smem stride of the [P][cluster] arrays; column 32 is an all-zero cluster

struct FastArgs {
  double *lik;     // [units][ncol][ntp]   likelihood table (global, L2 resident)
  double *uq;      // [units][2][ntp]      uniforms, sum_p x_ip^2
  int *resume_l;   // [units] sweep at which the unit handed over to the generic kernel, -1 = done
  int *resume_it;  // [units] patient position of the hand-over
  int *bail_count; // [1]
  int KS, ntp, ncol, n_units;
  int unit_bytes, tab_bytes;
};

__host__ __device__ inline int fast_pp(int P) { return (P + 1) & ~1; }
// bytes of shared memory of one unit
__host__ __device__ inline size_t fast_unit_bytes(int P, int KS, int ntp, int CC, int D) {
  size_t dbl = (size_t)2 * P * TPPMX_FAST_KSTR + (ntp + 1) + 2 * fast_pp(P) + (KS + CC) + (size_t)(KS + CC) * D;
  size_t in = (size_t)KS + 2 * (KS + CC) + 2 * ntp;
  return ((dbl * 8 + in * 4) + 15) & ~(size_t)15;
}
__host__ __device__ inline size_t fast_tab_bytes(int ntmax) { return (((size_t)3 * (ntmax + 2) * 8) + 15) & ~(size_t)15; }

struct UnitSmem {
  double *sumx, *sumx2, *logn, *xs, *wsc, *E;
  int *nj, *col, *freec, *lab, *pat;
};

__device__ __forceinline__ UnitSmem fast_carve(char *base, int P, int KS, int ntp, int CC, int D) {
  UnitSmem U;
  double *d = reinterpret_cast<double *>(base);
  U.sumx = d; d += (size_t)P * TPPMX_FAST_KSTR;
  U.sumx2 = d; d += (size_t)P * TPPMX_FAST_KSTR;
  U.logn = d; d += ntp + 1;
  U.xs = d; d += 2 * fast_pp(P);
  U.wsc = d; d += KS + CC;
  U.E = d; d += (size_t)(KS + CC) * D;
  int *q = reinterpret_cast<int *>(d);
  U.nj = q; q += KS;
  U.col = q; q += KS + CC;
  U.freec = q; q += KS + CC;
  U.lab = q; q += ntp;
  U.pat = q;
  return U;
}

template <int LPU>
__device__ __forceinline__ double unit_max(double v, unsigned um) {
#pragma unroll
  for (int o = LPU / 2; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(um, v, o, LPU));
  return v;
}
template <int LPU>
__device__ __forceinline__ double unit_sum(double v, unsigned um) {
#pragma unroll
  for (int o = LPU / 2; o > 0; o >>= 1) v += __shfl_xor_sync(um, v, o, LPU);
  return v;
}
template <int LPU>
__device__ __forceinline__ double unit_incl_scan(double v, unsigned um, int ul) {
#pragma unroll
  for (int o = 1; o < LPU; o <<= 1) {
    const double t = __shfl_up_sync(um, v, o, LPU);
    if (ul >= o) v += t;
  }
  return v;
}

// lgamma(x), x > 0, branch-free: shift x < 8 up by 8 with the recurrence, then the Stirling
// series to y^-13 (truncation < 1e-15 at y >= 8).  |error| <= 7e-15 (1 + |lgamma(x)|) against
// libm on [e^-12, e^12] (tests/test_oracle_kat.py restates and checks the same formula).
__device__ __forceinline__ double lgamma_pos(double x) {
  const bool small = x < 8.0;
  double p = x * (x + 1.0);
  p *= (x + 2.0) * (x + 3.0);
  double q = (x + 4.0) * (x + 5.0);
  q *= (x + 6.0) * (x + 7.0);
  p *= q;
  const double y = small ? x + 8.0 : x;
  const double r = 1.0 / y, r2 = r * r;
  double s = 1.0 / 156.0;
  s = fma(s, r2, -691.0 / 360360.0);
  s = fma(s, r2, 1.0 / 1188.0);
  s = fma(s, r2, -1.0 / 1680.0);
  s = fma(s, r2, 1.0 / 1260.0);
  s = fma(s, r2, -1.0 / 360.0);
  s = fma(s, r2, 1.0 / 12.0);
  const double lg = fma(y - 0.5, log(y), -y) + 0.91893853320467274178 + s * r;
  return lg - log(small ? p : 1.0);
}

// L_i(eta) without its cluster-independent part (:570-576):
//   sum_k [ gamma_k log JJ_ik - lgamma(gamma_k) ],  gamma_k = exp(eta_k) exp(zb_ik)
// The only call site of the transcendental code (keeps the kernel's instruction footprint small).
__device__ __noinline__ double fast_lik(const double *E, const double *Z, const double *ljj, int D) {
  double w = 0.0;
  for (int k = 0; k < D; k++) {
    const double g = E[k] * Z[k];
    w += fma(g, ljj[k], -lgamma_pos(g));
  }
  return w;
}

// cold path: more than LPU candidate clusters.  Scores in chunks through U.wsc; returns newci.
template <int LPU>
__device__ __noinline__ int fast_score_multi(UnitSmem U, const double *t_dA, const double *t_HN,
                                             const double *t_HY, const double *lik, int ntp, int it,
                                             int K, int CC, int KS, int P, const double *xs, double qx,
                                             double uu, double dV, double simscale, double iv2,
                                             double c0, double hiv2, bool nolik, unsigned um, int ul,
                                             int sub) {
  const int ntot = K + CC;
  double wmax = -INFINITY;
  for (int j = ul; j < ntot; j += LPU) {
    const bool occ = j < K;
    const int n = occ ? U.nj[j] : 0;
    const double *sx = U.sumx + (occ ? j : TPPMX_FAST_KSMAX);
    double a = 0.0, bb = 0.0;
    for (int p = 0; p < P; p++) {
      const double t = fma(sx[p * TPPMX_FAST_KSTR], iv2, c0);
      a = fma(t, t, a);
      bb = fma(t, xs[p], bb);
    }
    const double tY = a + iv2 * (2.0 * bb + iv2 * qx);
    const double d = t_dA[n] - hiv2 * qx + (t_HY[n] * tY - t_HN[n] * a);
    const int c = occ ? U.col[j] : U.col[KS + (j - K)];
    const double lv = nolik ? 0.0 : lik[(size_t)c * ntp + it];
    const double w = dV + U.logn[n] + simscale * d + lv;
    U.wsc[j] = w;
    wmax = fmax(wmax, w);
  }
  wmax = unit_max<LPU>(wmax, um);
  __syncwarp(um);
  double den = 0.0;
  for (int j = ul; j < ntot; j += LPU) {
    const double e = exp(U.wsc[j] - wmax);
    U.wsc[j] = e;
    den += e;
  }
  den = unit_sum<LPU>(den, um);
  __syncwarp(um);
  int newci = ntot;
  double carry = 0.0;
  for (int c0j = 0; c0j < ntot; c0j += LPU) {
    const int j = c0j + ul;
    const double pj = (j < ntot) ? U.wsc[j] / den : 0.0;
    const double cum = carry + unit_incl_scan<LPU>(pj, um, ul);
    unsigned hit = __ballot_sync(um, (j < ntot) && (uu < cum));
    hit = (LPU == 32) ? hit : ((hit >> (sub * LPU)) & ((1u << LPU) - 1u));
    if (hit) {
      newci = c0j + __ffs(hit);
      break;
    }
    carry = __shfl_sync(um, cum, LPU - 1, LPU);
  }
  return newci;
}

template <int LPU, int XR>
__global__ void __launch_bounds__(32, 12)
sweep_fast_kernel(DevBatch b, FastArgs fa, uint64_t seed, int iter0, int n_sweeps) {
  constexpr int UPW = 32 / LPU;
  constexpr int KSTR = TPPMX_FAST_KSTR;
  extern __shared__ __align__(16) char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPU, ul = lane % LPU;
  const unsigned um = (LPU == 32) ? 0xffffffffu : (((1u << LPU) - 1u) << (sub * LPU));
  const int unit = blockIdx.x * UPW + sub;
  const int P = b.P, D = b.D, Q = b.Q, CC = b.CC, KS = fa.KS, ntp = fa.ntp, kcap = b.kcap;
  const int Pp = fast_pp(P);
  const SimConst s = make_simconst(b);

  // ---- per-CTA tables over n (Normal-Normal log terms, utils_ct.cpp:389-400 via sim.cuh NNRow)
  double *t_dA = reinterpret_cast<double *>(smem_raw);
  double *t_HN = t_dA + (b.ntmax + 2);
  double *t_HY = t_HN + (b.ntmax + 2);
  for (int n = lane; n <= b.ntmax; n += 32) {
    const NNRow r0 = b.nn_tab[n], r1 = b.nn_tab[n + 1];
    const double dP = (double)P;
    t_dA[n] = (n == 0) ? dP * r1.A0 : dP * (r1.A0 - r0.A0);
    t_HN[n] = (n == 0) ? 0.0 : r0.H;
    t_HY[n] = r1.H;
  }
  __syncwarp();
  if (unit >= fa.n_units) return;

  const int chain = unit / b.T, tt = unit % b.T;
  const int ds = b.chain_ds[chain];
  const int nt = b.arm_n[(size_t)ds * b.T + tt];
  const Rng rng = make_rng(seed, b.chain_id[chain]);
  const UnitSmem U = fast_carve(smem_raw + fa.tab_bytes + (size_t)sub * fa.unit_bytes, P, KS, ntp, CC, D);

  // global state of this unit
  double *g_sumx = b.sumx + ((size_t)chain * b.T + tt) * P * kcap;
  double *g_sumx2 = b.sumx2 + ((size_t)chain * b.T + tt) * P * kcap;
  int *g_nj = b.nj + ((size_t)chain * b.T + tt) * kcap;
  int *g_lab = b.labels + ((size_t)chain * b.T + tt) * b.ntmax;
  double *g_eta = b.eta + ((size_t)chain * b.T + tt) * D * kcap;
  double *g_etae = b.etae + ((size_t)chain * b.T + tt) * D * CC;
  const int *g_pat = b.arm_pat + ((size_t)ds * b.T + tt) * b.ntmax;
  const double *g_x = b.x + (size_t)ds * b.nobs * P;
  double *lik = fa.lik + (size_t)unit * fa.ncol * ntp;
  double *uq = fa.uq + (size_t)unit * 2 * ntp;
  double *g_ez = b.zb + (size_t)chain * b.nobs * D;   // here: exp(beta'z) (scratch, rebuilt per sweep)
  double *g_ljj = b.ljj + (size_t)chain * b.nobs * D;
  const double *beta = b.beta + (size_t)chain * Q * D;

  int K = b.nclu[(size_t)chain * b.T + tt];
  const double sigma = b.sigma[(size_t)chain * b.T + tt];
  const int idxs = b.idxs[chain];
  const double iv2 = s.iv2, c0 = s.c0, hiv2 = s.hiv2;
  const double simscale = (s.calibration == 2) ? s.calscale : 1.0;
  const bool nolik = s.nolik != 0;

  // ---- stage the unit's state in shared memory
  for (int e = ul; e < P * KSTR; e += LPU) {
    const int p = e / KSTR, j = e % KSTR;
    U.sumx[e] = (j < KS) ? g_sumx[(size_t)p * kcap + j] : 0.0;
    U.sumx2[e] = (j < KS) ? g_sumx2[(size_t)p * kcap + j] : 0.0;
  }
  for (int j = ul; j < KS; j += LPU) U.nj[j] = g_nj[j];
  for (int it = ul; it < nt; it += LPU) {
    U.lab[it] = g_lab[it];
    U.pat[it] = g_pat[it];
  }
  // column indirection: clusters 0..K-1 -> cols 0..K-1, aux m -> col K+m, the rest is free
  int nfree = 0;
  {
    const int Kc = K < KS ? K : KS;
    for (int j = ul; j < Kc; j += LPU) U.col[j] = j;
    for (int m = ul; m < CC; m += LPU) U.col[KS + m] = Kc + m;
    nfree = fa.ncol - Kc - CC;
    for (int f = ul; f < nfree; f += LPU) U.freec[f] = Kc + CC + f;
  }
  // cohesion by cluster size: logn[0] is the auxiliary-cluster value (:677 / :680)
  if (ul == 0) U.logn[0] = (s.cohesion == 1) ? log(b.model.alpha) - log((double)CC) : 0.0;
  for (int n = 1 + ul; n <= nt; n += LPU) U.logn[n] = (s.cohesion == 1) ? log((double)n) : log((double)n - sigma);
  __syncwarp(um);

  auto dV_of = [&](int Kcur) -> double {
    if (s.cohesion != 2 || Kcur < 1) return 0.0;
    const size_t base = ((size_t)tt * 2) * b.ntmax;
    return b.logV[(base + b.ntmax + (Kcur - 1)) * b.G + idxs] - b.logV[(base + (Kcur - 1)) * b.G + idxs];
  };
  auto colmap = [&](int j, int Kcur) -> int { return (j < Kcur) ? U.col[j] : U.col[KS + (j - Kcur)]; };

  int bail_l = -1, bail_it = 0;

  for (int l = iter0; l < iter0 + n_sweeps && bail_l < 0; l++) {
    if (K > KS) {
      bail_l = l;
      bail_it = 0;
      break;
    }
    // =========================== parallel pre-pass of the sweep ===========================
    // :351-360 auxiliary eta ~ N(0, I)
    for (int mm = ul; mm < CC; mm += LPU)
      rng_normals(rng, l, TPPMX_SITE_AUX_SWEEP, tt, (uint32_t)mm, D, g_etae + mm, CC);
    __syncwarp(um);
    // per patient: uniform, sum_p x^2, exp(beta'z), log JJ
    for (int it = ul; it < nt; it += LPU) {
      const int i = U.pat[it];
      const double *xi = g_x + (size_t)i * P;
      double qx = 0.0;
      for (int p = 0; p < P; p++) qx = fma(xi[p], xi[p], qx);
      double uu, u1;
      rng_u2(rng, l, TPPMX_SITE_ALLOC_U, tt, (uint32_t)i, 0, uu, u1);
      uq[it] = uu;
      uq[ntp + it] = qx;
      if (!nolik) {
        const double *zi = b.z + ((size_t)ds * b.nobs + i) * Q;
        const double *JJ = b.JJ + ((size_t)chain * b.nobs + i) * D;
        for (int k = 0; k < D; k++) {
          double acc = 0.0;
          for (int q = 0; q < Q; q++) acc += beta[k + q * D] * zi[q];
          g_ez[(size_t)i * D + k] = exp(acc);
          g_ljj[(size_t)i * D + k] = log(JJ[k]);
        }
      }
    }
    if (!nolik) {
      const int ntot = K + CC;
      // exp(eta) of every candidate cluster, stored by column id
      for (int e = ul; e < ntot * D; e += LPU) {
        const int c = e / D, k = e - c * D;
        const double eta = (c < K) ? g_eta[(size_t)k * kcap + c] : g_etae[k * CC + (c - K)];
        U.E[colmap(c, K) * D + k] = exp(eta);
      }
      __syncwarp(um);
      // likelihood table: all (candidate cluster, patient) pairs, flattened over the lanes
      for (int e = ul; e < ntot * nt; e += LPU) {
        const int c = e / nt, it = e - c * nt;
        const int cl = colmap(c, K);
        const int i = U.pat[it];
        lik[(size_t)cl * ntp + it] = fast_lik(U.E + cl * D, g_ez + (size_t)i * D, g_ljj + (size_t)i * D, D);
      }
    }
    __syncwarp(um);

    // =========================== sequential reallocation steps ===========================
    double dV = dV_of(K);
    bool act0 = ul < K + CC;
    const double *likp = lik + (size_t)(act0 ? colmap(ul, K) : 0) * ntp;
    double likv = 0.0, uu = 0.0, qx = 0.0;
    if (nt > 0) {  // operands of step 0
      const int i0 = U.pat[0];
#pragma unroll
      for (int r = 0; r < XR; r++) {
        const int p = ul + r * LPU;
        if (p < P) U.xs[p] = g_x[(size_t)i0 * P + p];
      }
      uu = uq[0];
      qx = uq[ntp];
      if (!nolik && act0) likv = likp[0];
    }
    __syncwarp(um);

    for (int it = 0; it < nt; it++) {
      if (__builtin_expect(K >= KS, 0)) {  // a birth could overflow the staged capacity: hand over
        bail_l = l;
        bail_it = it;
        break;
      }
      const double *xs = U.xs + (it & 1) * Pp;
      // ---- prefetch the operands of step it+1 (addresses do not depend on this step)
      const bool has_next = (it + 1 < nt);
      double xr_n[XR], lik_n = 0.0, uu_n = 0.0, qx_n = 0.0;
      if (has_next) {
        const double *xrow = g_x + (size_t)U.pat[it + 1] * P;
#pragma unroll
        for (int r = 0; r < XR; r++) {
          const int p = ul + r * LPU;
          xr_n[r] = (p < P) ? xrow[p] : 0.0;
        }
        uu_n = uq[it + 1];
        qx_n = uq[ntp + it + 1];
        if (!nolik && act0) lik_n = likp[it + 1];
      }
      bool structural = false;

      // ---- :368-457 take the patient out of its cluster
      const int zi = U.lab[it] - 1;
      const int nzi = U.nj[zi];
      __syncwarp(um);
      if (__builtin_expect(nzi > 1, 1)) {
        if (ul == 0) U.nj[zi] = nzi - 1;
#pragma unroll
        for (int r = 0; r < XR; r++) {
          const int p = ul + r * LPU;
          if (p < P) {
            const double xv = xs[p];
            U.sumx[p * KSTR + zi] = __dsub_rn(U.sumx[p * KSTR + zi], xv);
            U.sumx2[p * KSTR + zi] = __dsub_rn(U.sumx2[p * KSTR + zi], __dmul_rn(xv, xv));
          }
        }
      } else {  // singleton (:384-457)
        structural = true;
        const int iaux = zi + 1, Kc = K;
        if (iaux < Kc) {  // swap with the last cluster
          for (int ii = ul; ii < nt; ii += LPU)
            if (U.lab[ii] == Kc) U.lab[ii] = iaux;
          __syncwarp(um);
          if (ul == 0) {
            U.lab[it] = Kc;
            U.nj[iaux - 1] = U.nj[Kc - 1];
            U.nj[Kc - 1] = 1;
            const int c = U.col[iaux - 1];
            U.col[iaux - 1] = U.col[Kc - 1];
            U.col[Kc - 1] = c;
          }
          for (int k = ul; k < D; k += LPU) {
            const double t = g_eta[(size_t)k * kcap + iaux - 1];
            g_eta[(size_t)k * kcap + iaux - 1] = g_eta[(size_t)k * kcap + Kc - 1];
            g_eta[(size_t)k * kcap + Kc - 1] = t;
          }
          for (int p = ul; p < P; p += LPU) {
            double t = U.sumx[p * KSTR + iaux - 1];
            U.sumx[p * KSTR + iaux - 1] = U.sumx[p * KSTR + Kc - 1];
            U.sumx[p * KSTR + Kc - 1] = t;
            t = U.sumx2[p * KSTR + iaux - 1];
            U.sumx2[p * KSTR + iaux - 1] = U.sumx2[p * KSTR + Kc - 1];
            U.sumx2[p * KSTR + Kc - 1] = t;
          }
          __syncwarp(um);
        }
        // empty slot Kc-1 (keeps the rounding residue, as the reference does)
        if (ul == 0) {
          U.nj[Kc - 1] -= 1;
          U.freec[nfree] = U.col[Kc - 1];
        }
        nfree += 1;
        for (int p = ul; p < P; p += LPU) {
          const double xv = xs[p];
          U.sumx[p * KSTR + Kc - 1] = __dsub_rn(U.sumx[p * KSTR + Kc - 1], xv);
          U.sumx2[p * KSTR + Kc - 1] = __dsub_rn(U.sumx2[p * KSTR + Kc - 1], __dmul_rn(xv, xv));
        }
        K = Kc - 1;
        dV = dV_of(K);
      }
      __syncwarp(um);
      if (__builtin_expect(structural, 0)) {
        act0 = ul < K + CC;
        likp = lik + (size_t)(act0 ? colmap(ul, K) : 0) * ntp;
        likv = (!nolik && act0) ? likp[it] : 0.0;
      }

      // ---- :460-814 score the K occupied + CC auxiliary clusters, normalise, draw
      const int ntot = K + CC;
      int newci;
      if (__builtin_expect(ntot <= LPU, 1)) {
        const bool occ = ul < K;
        const int n = occ ? U.nj[ul] : 0;
        const double *sx = U.sumx + (occ ? ul : TPPMX_FAST_KSMAX);  // column 32: the empty cluster
        double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
        int p = 0;
        for (; p + 1 < P; p += 2) {
          const double t0 = fma(sx[p * KSTR], iv2, c0), t1 = fma(sx[(p + 1) * KSTR], iv2, c0);
          a0 = fma(t0, t0, a0);
          a1 = fma(t1, t1, a1);
          b0 = fma(t0, xs[p], b0);
          b1 = fma(t1, xs[p + 1], b1);
        }
        if (p < P) {
          const double t0 = fma(sx[p * KSTR], iv2, c0);
          a0 = fma(t0, t0, a0);
          b0 = fma(t0, xs[p], b0);
        }
        const double a = a0 + a1, bb = b0 + b1;
        const double tY = a + iv2 * (2.0 * bb + iv2 * qx);
        const double d = t_dA[n] - hiv2 * qx + (t_HY[n] * tY - t_HN[n] * a);
        const double w = act0 ? dV + U.logn[n] + simscale * d + likv : -INFINITY;
        const double wmax = unit_max<LPU>(w, um);
        const double e0 = exp(w - wmax);
        const double den = unit_sum<LPU>(e0, um);
        const double cum = unit_incl_scan<LPU>(e0 / den, um, ul);
        unsigned hit = __ballot_sync(um, act0 && (uu < cum));
        hit = (LPU == 32) ? hit : ((hit >> (sub * LPU)) & ((1u << LPU) - 1u));
        newci = hit ? __ffs(hit) : ntot;  // :807 default when rounding leaves uu >= total
      } else {
        newci = fast_score_multi<LPU>(U, t_dA, t_HN, t_HY, lik, ntp, it, K, CC, KS, P, xs, qx, uu, dV,
                                      simscale, iv2, c0, hiv2, nolik, um, ul, sub);
      }

      // ---- :820-856 assign, update sufficient statistics
      int lab;
      if (__builtin_expect(newci <= K, 1)) {
        lab = newci;
        if (ul == 0) {
          U.lab[it] = lab;
          U.nj[lab - 1] += 1;
        }
      } else {  // a new cluster takes the auxiliary eta (:829-845)
        structural = true;
        const int i = U.pat[it];
        const int m = newci - K - 1;
        lab = K + 1;
        const int newcol = U.freec[nfree - 1];
        nfree -= 1;
        __syncwarp(um);
        if (ul == 0) {
          U.lab[it] = lab;
          U.nj[lab - 1] = 1;
          U.col[lab - 1] = U.col[KS + m];
          U.col[KS + m] = newcol;
        }
        for (int k = ul; k < D; k += LPU) g_eta[(size_t)k * kcap + lab - 1] = g_etae[k * CC + m];
        __syncwarp(um);
        if (ul == 0) rng_normals(rng, l, TPPMX_SITE_AUX_REFRESH, tt, (uint32_t)i, D, g_etae + m, CC);
        __syncwarp(um);
        if (!nolik) {  // likelihood column of the refreshed auxiliary eta, remaining patients
          for (int k = ul; k < D; k += LPU) U.E[newcol * D + k] = exp(g_etae[k * CC + m]);
          __syncwarp(um);
          for (int it2 = it + 1 + ul; it2 < nt; it2 += LPU) {
            const int i2 = U.pat[it2];
            lik[(size_t)newcol * ntp + it2] =
                fast_lik(U.E + newcol * D, g_ez + (size_t)i2 * D, g_ljj + (size_t)i2 * D, D);
          }
        }
        K = lab;
        dV = dV_of(K);
      }
#pragma unroll
      for (int r = 0; r < XR; r++) {
        const int p = ul + r * LPU;
        if (p < P) {
          const double xv = xs[p];
          U.sumx[p * KSTR + lab - 1] = __dadd_rn(U.sumx[p * KSTR + lab - 1], xv);
          U.sumx2[p * KSTR + lab - 1] = __dadd_rn(U.sumx2[p * KSTR + lab - 1], __dmul_rn(xv, xv));
        }
      }
      // ---- stage the operands of the next step
      if (has_next) {
#pragma unroll
        for (int r = 0; r < XR; r++) {
          const int p = ul + r * LPU;
          if (p < P) U.xs[((it + 1) & 1) * Pp + p] = xr_n[r];
        }
        uu = uu_n;
        qx = qx_n;
        likv = lik_n;
      }
      __syncwarp(um);
      if (__builtin_expect(structural, 0)) {
        act0 = ul < K + CC;
        likp = lik + (size_t)(act0 ? colmap(ul, K) : 0) * ntp;
        likv = (has_next && !nolik && act0) ? likp[it + 1] : 0.0;
      }
    }
  }

  // ---- write the unit back in the library's global layout
  __syncwarp(um);
  for (int e = ul; e < P * KSTR; e += LPU) {
    const int p = e / KSTR, j = e % KSTR;
    if (j < KS) {
      g_sumx[(size_t)p * kcap + j] = U.sumx[e];
      g_sumx2[(size_t)p * kcap + j] = U.sumx2[e];
    }
  }
  for (int j = ul; j < KS; j += LPU) g_nj[j] = U.nj[j];
  for (int it = ul; it < nt; it += LPU) g_lab[it] = U.lab[it];
  if (ul == 0) {
    b.nclu[(size_t)chain * b.T + tt] = K;
    fa.resume_l[unit] = bail_l;
    fa.resume_it[unit] = bail_it;
    if (bail_l >= 0) atomicAdd(fa.bail_count, 1);
  }
  __syncwarp(um);
  // loggamma (:825, :839): eta* is constant during the sweep, so the value every patient ends
  // with is eta*[label] + beta'z, evaluated here once in the reference's operation order
  // (utils_ct.cpp:514-527).  A unit that handed over leaves this to the generic kernel for
  // the patients it still has to visit; rewriting the visited ones is value-preserving.
  if (n_sweeps > 0) {
    for (int it = ul; it < nt; it += LPU) {
      const int i = U.pat[it];
      const int lb = U.lab[it];
      const double *zi = b.z + ((size_t)ds * b.nobs + i) * Q;
      for (int k = 0; k < D; k++) {
        double lg = g_eta[(size_t)k * kcap + lb - 1];
        for (int q = 0; q < Q; q++) lg = __dadd_rn(lg, __dmul_rn(beta[k + q * D], zi[q]));
        b.lg[((size_t)chain * b.nobs + i) * D + k] = lg;
      }
    }
  }
}

}  // namespace tppmx
