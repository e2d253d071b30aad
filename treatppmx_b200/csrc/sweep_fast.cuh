// sweep_fast.cuh — the production reallocation-sweep kernel: reference src/dm_ppmx_ct.cpp:347-860.
// The default model switches (PPMx=1, Normal-Normal auxiliary similarity, calibration 0|2, no categorical
// covariates, reuse=1: consim==1, similarity==1) are what the design below describes; template flags add
// calibration 1 (CAL1), the NNIG similarity (NNIG), categorical covariates (CAT), the double dipper (DDIP),
// reuse = 0 (R0) and arms of thousands of patients (BIG) -- each documented at the kernel's template head.
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
#include "fastmath.cuh"
#include "sim.cuh"
#include "sweep_fast_args.h"

namespace tppmx {

struct UnitSmem {
  double *sumx, *sumx2, *logn, *xs, *wsc, *E;
  int *nj, *col, *freec, *lab, *pat;
  int *njc, *xcs;  // categorical counts [ncat * maxC][KSTR], ring of four categorical rows (CAT instantiation)
};

__device__ __forceinline__ UnitSmem fast_carve(char *base, int P, int KS, int ntp, int CC, int D, int big, int ncatC = 0) {
  UnitSmem U;
  double *d = reinterpret_cast<double *>(base);
  U.sumx = d; d += (size_t)P * TPPMX_FAST_KSTR;
  U.sumx2 = d; d += (size_t)P * TPPMX_FAST_KSTR;
  U.logn = d; d += big ? 0 : ntp + 1;
  U.xs = d; d += 4 * fast_pp(P);  // ring of four covariate rows (patients it-1 .. it+2)
  U.wsc = d; d += KS + CC;
  U.E = d; d += (size_t)(KS + CC) * D;
  int *q = reinterpret_cast<int *>(d);
  U.nj = q; q += KS;
  U.col = q; q += KS + CC;
  U.freec = q; q += KS + CC;
  U.lab = q; q += big ? 0 : ntp + 1;
  U.pat = q;  // ntp + 1 entries (the prefetch of step it+1 may read one past the arm)
  q += big ? 0 : ntp + 1;
  U.njc = q;
  U.xcs = q + (size_t)ncatC * TPPMX_FAST_KSTR;
  return U;   // big: logn / lab / pat are re-pointed at global memory by the kernel
}

constexpr unsigned TPPMX_FULL = 0xffffffffu;

// Debug build (-DTPPMX_CHECK, treatppmx_b200/libtppmx_check.so): index invariants of the fast
// kernel are verified on the device and reported through the chain's status word
// (error code 100 + n), since compute-sanitizer is not available on the GPU pool.
#ifdef TPPMX_CHECK
#define TPPMX_ASSERT(cond, n)                                      \
  do {                                                             \
    if (!(cond)) atomicExch(&b.status[chain], 100 + (n));          \
  } while (0)
#else
#define TPPMX_ASSERT(cond, n) \
  do {                        \
  } while (0)
#endif

// The units of a warp run CONVERGED: every collective uses the full mask (width-LPU shuffles
// keep the units apart) and unit-specific work is predicated.  Rare events (singleton removal,
// cluster birth, > LPU candidates) are entered warp-uniformly through __any_sync.
// maximum over the unit's lanes: one redux.sync on an order-preserving integer image of the
// float instead of log2(LPU) shuffle + max levels (the step's dependent chain is what bounds it)
// A shift for the unit's softmax: the largest weight of the unit's lanes with its low mantissa word
// cleared (within 2^-20 relative of the maximum, which is all the shift has to be: common to the lanes
// and close to the maximum).  The high word of an fp64 orders like a sign-magnitude integer, so no
// conversion is needed; one full-mask redux per unit of the warp (a redux over a partial mask is
// serialised per distinct mask behind a WARPSYNC.EXCLUSIVE loop; with the full mask they pipeline).
template <int LPU>
__device__ __forceinline__ double unit_shift(double w, int sub) {
  int i = __double2hiint(w);
  i ^= (i >> 31) & 0x7fffffff;
  int r = i;
#pragma unroll
  for (int s = 0; s < 32 / LPU; s++) {
    const int m = __reduce_max_sync(TPPMX_FULL, (sub == s) ? i : (int)0x80000000);
    if (sub == s) r = m;
  }
  r ^= (r >> 31) & 0x7fffffff;
  return __hiloint2double(r, 0);
}
template <int LPU>
__device__ __forceinline__ double unit_sum(double v) {
#pragma unroll
  for (int o = LPU / 2; o > 0; o >>= 1) v += __shfl_xor_sync(TPPMX_FULL, v, o, LPU);
  return v;
}
template <int LPU>
__device__ __forceinline__ double unit_incl_scan(double v, int ul) {
#pragma unroll
  for (int o = 1; o < LPU; o <<= 1) {
    const double t = __shfl_up_sync(TPPMX_FULL, v, o, LPU);
    if (ul >= o) v += t;
  }
  return v;
}
template <int LPU>
__device__ __forceinline__ int warp_max_over_units(int v) {
#pragma unroll
  for (int o = 16; o >= LPU; o >>= 1) v = max(v, __shfl_xor_sync(TPPMX_FULL, v, o));
  return v;
}

// L_i(eta) without its cluster-independent part (:570-576):
//   sum_k [ gamma_k log JJ_ik - lgamma(gamma_k) ],  gamma_k = exp(eta_k) exp(zb_ik)
// D <= 3 (the package's case): the patient's exp(zb) / log JJ arrive in registers, so the caller can
// fetch them one pair ahead.  The only call site of the lgamma code for D <= 3 (keeps the kernel's
// instruction footprint small).  Missing categories: z = 1, l = 0 (their terms are skipped).
static __device__ __noinline__ double fast_lik3(const double *E, int D, double z0, double z1, double z2, double j0,
                                                double j1, double j2) {
  const bool h1 = D > 1, h2 = D > 2;
  const double g0 = E[0] * z0;
  const double g1 = h1 ? E[1] * z1 : 1.0;
  const double g2 = h2 ? E[2] * z2 : 1.0;
  double l0, l1, l2;
  lgamma_fast3(g0, g1, g2, l0, l1, l2);
  double w = fma(g0, j0, -l0);
  if (h1) w += fma(g1, j1, -l1);
  if (h2) w += fma(g2, j2, -l2);
  return w;
}
static __device__ __noinline__ double fast_lik_loop(const double *E, const double *Z, const double *ljj, int D) {
  double w = 0.0;
  // three outcome categories per trip: the three lgamma evaluations are independent straight-line
  // code, so their dependent fma chains interleave (fp64 latency hiding inside one warp)
  for (int k = 0; k < D; k += 3) {
    const bool h1 = k + 1 < D, h2 = k + 2 < D;
    const double g0 = E[k] * Z[k];
    const double g1 = h1 ? E[k + 1] * Z[k + 1] : 1.0;
    const double g2 = h2 ? E[k + 2] * Z[k + 2] : 1.0;
    double l0, l1, l2;
    lgamma_fast3(g0, g1, g2, l0, l1, l2);
    w += fma(g0, ljj[k], -l0);
    if (h1) w += fma(g1, ljj[k + 1], -l1);
    if (h2) w += fma(g2, ljj[k + 2], -l2);
  }
  return w;
}
// pointer form (cluster birth, D > 3)
template <int DR>
__device__ __forceinline__ double fast_lik(const double *E, const double *Z, const double *ljj, int D) {
  if constexpr (DR == 3)
    return fast_lik3(E, D, Z[0], D > 1 ? Z[1] : 1.0, D > 2 ? Z[2] : 1.0, ljj[0], D > 1 ? ljj[1] : 0.0, D > 2 ? ljj[2] : 0.0);
  else
    return fast_lik_loop(E, Z, ljj, D);
}

// Register version for the pre-pass: the patient's exp(zb) / log JJ stay in registers while the
// lane walks over the candidate clusters (one L2 round trip per patient instead of one per
// table entry).  DR = 3 or TPPMX_MAXD.
template <int DR>
__device__ __forceinline__ double fast_lik_reg(const double *E, const double (&Z)[DR], const double (&ljj)[DR], int D) {
  double w = 0.0;
#pragma unroll
  for (int k = 0; k < DR; k += 3) {
    if (k < D) {
      const bool h1 = (k + 1 < DR) && (k + 1 < D), h2 = (k + 2 < DR) && (k + 2 < D);
      const double g0 = E[k] * Z[k];
      const double g1 = h1 ? E[k + 1] * Z[(k + 1 < DR) ? k + 1 : k] : 1.0;
      const double g2 = h2 ? E[k + 2] * Z[(k + 2 < DR) ? k + 2 : k] : 1.0;
      double l0, l1, l2;
    lgamma_fast3(g0, g1, g2, l0, l1, l2);
      w += fma(g0, ljj[k], -l0);
      if (h1) w += fma(g1, ljj[(k + 1 < DR) ? k + 1 : k], -l1);
      if (h2) w += fma(g2, ljj[(k + 2 < DR) ? k + 2 : k], -l2);
    }
  }
  return w;
}

// cold path: a unit of the warp has more than LPU candidate clusters (`mine`).  Scores in
// chunks through U.wsc; returns newci for the units with `mine`.
template <int LPU>
__device__ __noinline__ int fast_score_multi(UnitSmem U, const double2 *tab2, const double *logn,
                                             const double *lik, int ntp, int it,
                                             int K, int CC, int KS, int P, const double *xs, double qx,
                                             double uu, double dV, double simscale, double iv2,
                                             double c0, double hiv2, bool nolik, bool mine, int ul,
                                             int sub, double *prow, int pstride) {
  const int ntot = mine ? K + CC : 0;
  const int ntot_w = warp_max_over_units<LPU>(ntot);
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
    const double2 t01 = tab2[2 * n], t23 = tab2[2 * n + 1];  // (dA, HN), (HY, -)
    const double d = t01.x - hiv2 * qx + (t23.x * tY - t01.y * a);
    const int c = occ ? U.col[j] : U.col[KS + (j - K)];
    const double lv = nolik ? 0.0 : lik[(size_t)c * ntp + it];
    const double w = dV + logn[n] + simscale * d + lv;
    U.wsc[j] = w;
    wmax = fmax(wmax, w);
  }
  __syncwarp();
#pragma unroll
  for (int o = LPU / 2; o > 0; o >>= 1) wmax = fmax(wmax, __shfl_xor_sync(TPPMX_FULL, wmax, o, LPU));
  double den = 0.0;
  for (int j = ul; j < ntot; j += LPU) {
    const double e = fast_exp(U.wsc[j] - wmax);
    U.wsc[j] = e;
    den += e;
  }
  __syncwarp();
  den = unit_sum<LPU>(den);
  if (prow && mine)  // parity probe: the normalised probabilities the draw is made from
    for (int j = ul; j < ntot && j < pstride; j += LPU) prow[j] = U.wsc[j] / den;
  const double thr = uu * den;
  int newci = ntot;
  bool found = false;
  double carry = 0.0;
  for (int c0j = 0; c0j < ntot_w; c0j += LPU) {
    const int j = c0j + ul;
    const double ej = (j < ntot) ? U.wsc[j] : 0.0;
    const double cum = carry + unit_incl_scan<LPU>(ej, ul);
    unsigned hit = __ballot_sync(TPPMX_FULL, (j < ntot) && (thr < cum));
    hit = (LPU == 32) ? hit : ((hit >> (sub * LPU)) & ((1u << LPU) - 1u));
    if (hit && !found) {
      newci = c0j + __ffs(hit);
      found = true;
    }
    carry = __shfl_sync(TPPMX_FULL, cum, LPU - 1, LPU);
  }
  return newci;
}

#ifndef TPPMX_FAST_MINB
#define TPPMX_FAST_MINB 12
#endif
// PC: compile-time number of covariates (0: run-time b.P)
// CAL1: calibration == 1 (src/dm_ppmx_ct.cpp:717-788): the similarities enter the weight normalised over the
//   clusters -- gtilN over all K + CC candidates ("without" value of an occupied cluster, the singleton value of
//   an auxiliary one), gtilY over the occupied ones -- so the step carries lgN_j and lgY_j themselves
//   (lgN = P A0[n] - sum_p sumx2/(2 v2) + H[n] a, sim.cuh) and two more log-sum-exps per unit.  The host keeps
//   K + CC <= LPU for this instantiation (staged capacity LPU - CC; larger partitions hand over).
// NNIG: consim == 2, the Normal-Normal-Inverse-Gamma similarity (src/utils_ct.cpp:427-459): per covariate the
//   score is nonlinear in (sumx, sumx2), so there is no dot-product form to pipeline; each step evaluates
//   gsimconNNIG for its lane's cluster with and without the patient straight from the staged statistics, in the
//   function's own operation order for the large cancelling terms, everything that depends on n only from an
//   eight-entry table row, and the P logarithms a0s log(b0s) folded into one log of a product per lane.
// CAT: categorical covariates beside the continuous ones (Dirichlet-multinomial similarity, src/utils_ct.cpp:468-495):
//   with the patient's category c* in covariate p the "with" minus "without" term is log(n_jc* + w) - log(n_j + C_p w)
//   (Gamma(x + 1) = x Gamma(x) on both lgamma differences): the counts n_jc live in shared memory beside sumx, the
//   two logarithms come from tables over the integers.
// DDIP: similarity == 2, the double-dipper Normal-Normal form (src/utils_ct.cpp:382-405 with DD = 1):
//   g = P B0[n] - sum_p sumx2/(2 v2) - H[n] sum_p t^2 + H2[n] sum_p u^2,  u = sumx/v2 + t = 2 t - c0, so beside a and b
//   the lane carries st = sum_p t (linear in the statistics: corrected by +- sum_p x / v2 on the touched lanes):
//   sum u^2 = 4 a - 4 c0 st + P c0^2,  sum u x = 2 b - c0 sum_p x.  The fourth table entry of row n is H2[n].
// R0: reuse == 0 (:605-610): the CC auxiliary eta are redrawn for every patient, so an auxiliary column of the
//   likelihood table holds patient-specific entries (drawn and evaluated in the pre-pass: the draws are counter-based,
//   keyed by (patient, slot)), a new cluster takes its patient's draw and gets a column of its own for the patients
//   that follow, and the auxiliary slots keep theirs.
template <int LPU, int XR, int DR, bool NOLIK, bool BIG, bool PROBE, int PC = 0, bool CAL1 = false, bool NNIG = false, bool CAT = false,
          bool DDIP = false, bool R0 = false>
__global__ void __launch_bounds__(32, TPPMX_FAST_MINB)
sweep_fast_kernel(DevBatch b, FastArgs fa, uint64_t seed, int iter0, int n_sweeps) {
  constexpr int UPW = 32 / LPU;
  constexpr int KSTR = TPPMX_FAST_KSTR;
  extern __shared__ __align__(16) char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPU, ul = lane % LPU;
  const int unit_raw = blockIdx.x * UPW + sub;
  const bool valid = unit_raw < fa.n_units;
  const int unit = valid ? unit_raw : 0;  // lanes of a missing unit shadow unit 0 read-only
  const int P = PC > 0 ? PC : b.P, D = b.D, Q = b.Q, CC = b.CC, KS = fa.KS, ntp = fa.ntp, kcap = b.kcap;
  const int Pp = fast_pp(P);
  const SimConst s = make_simconst(b);

  // ---- per-CTA table over n (Normal-Normal log terms, utils_ct.cpp:389-400 via sim.cuh NNRow):
  //      entry n = (dA, HN, HY, -), built on the host (fa.tabg); staged in shared memory unless BIG
  const double2 *tab2;
  if constexpr (BIG || NNIG) {
    tab2 = reinterpret_cast<const double2 *>(fa.tabg);
  } else {
    double2 *ts = reinterpret_cast<double2 *>(smem_raw);
    const double2 *tg = reinterpret_cast<const double2 *>(fa.tabg);
    for (int e = lane; e < 2 * (b.ntmax + 2); e += 32) ts[e] = tg[e];
    tab2 = ts;
  }

  const int chain = unit / b.T, tt = unit % b.T;
  const int ds = b.chain_ds[chain];
  const int nt = valid ? b.arm_n[(size_t)ds * b.T + tt] : 0;
  const int nt_w = warp_max_over_units<LPU>(nt);
  const Rng rng = make_rng(seed, b.chain_id[chain]);
  const int ncat = CAT ? b.ncat : 0, maxC = CAT ? b.maxC : 0, ncC = ncat * maxC;
  UnitSmem U = fast_carve(smem_raw + fa.tab_bytes + (size_t)sub * fa.unit_bytes, P, KS, ntp, CC, D, BIG ? 1 : 0, ncC);

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
  const int uqs = ntp + 1;
  double *uq = fa.uq + (size_t)unit * 4 * uqs;  // rows: uniform, sum_p x^2, x_{it-1} . x_it, sum_p x (double dipper)
  double *g_ez = b.zb + (size_t)chain * b.nobs * D;   // here: exp(beta'z) (scratch, rebuilt per sweep)
  double *g_ljj = b.ljj + (size_t)chain * b.nobs * D;
  const double *beta = b.beta + (size_t)chain * Q * D;

  int K = b.nclu[(size_t)chain * b.T + tt];
  const double sigma = b.sigma[(size_t)chain * b.T + tt];
  const int idxs = b.idxs[chain];
  const double iv2 = s.iv2, c0 = s.c0, hiv2 = s.hiv2;
  // PPMx == 0 (no covariate similarity, :401-409 skipped): the similarity term is multiplied by zero and the
  // statistics, which the reference then never touches, are not written back
  const double simscale = !s.PPMx ? 0.0 : ((s.calibration == 2) ? s.calscale : 1.0);

  // ---- stage the unit's state in shared memory
  for (int e = ul; e < P * KSTR; e += LPU) {
    const int p = e / KSTR, j = e % KSTR;
    U.sumx[e] = (j < KS) ? g_sumx[(size_t)p * kcap + j] : 0.0;
    U.sumx2[e] = (j < KS) ? g_sumx2[(size_t)p * kcap + j] : 0.0;
  }
  for (int j = ul; j < KS; j += LPU) U.nj[j] = g_nj[j];
  int *g_njc = b.njc + ((size_t)chain * b.T + tt) * (ncC > 0 ? ncC : 1) * kcap;
  const int *g_xc = b.xcat + (size_t)ds * b.nobs * (ncat > 0 ? ncat : 1);
  if constexpr (CAT)
    for (int e = ul; e < ncC * KSTR; e += LPU) {
      const int pc = e / KSTR, j = e % KSTR;
      U.njc[e] = (j < KS) ? g_njc[(size_t)pc * kcap + j] : 0;
    }
  // patient-indexed arrays: shared memory, or (BIG: arms of thousands of patients) global memory.
  // Kept in separate variables so the compiler sees one address space per instantiation.
  int *labp, *patp;
  double *lognp;
  if constexpr (BIG) {  // labels are updated in place in the batch's global array (ntmax entries, enough
                        // for every index the loop touches: lab[it] is only read for it < nt)
    labp = g_lab;
    patp = fa.patpad + (size_t)unit * (ntp + 1);
    lognp = fa.logng + (size_t)unit * (ntp + 1);
    if (valid)
      for (int it = ul; it <= ntp; it += LPU) patp[it] = (it < nt) ? g_pat[it] : 0;
  } else {
    labp = U.lab;
    patp = U.pat;
    lognp = U.logn;
    for (int it = ul; it <= ntp; it += LPU) {
      labp[it] = (it < nt) ? g_lab[it] : 1;
      patp[it] = (it < nt) ? g_pat[it] : 0;
    }
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
  if (ul == 0) lognp[0] = (s.cohesion == 1) ? log(b.model.alpha) - log((double)CC) : 0.0;
  for (int n = 1 + ul; n <= nt; n += LPU) lognp[n] = fast_log((s.cohesion == 1) ? (double)n : (double)n - sigma);
  __syncwarp();

  auto dV_of = [&](int Kcur) -> double {
    if (s.cohesion != 2 || Kcur < 1) return 0.0;
    const size_t base = ((size_t)tt * 2) * b.ntmax;
    return b.logV[(base + b.ntmax + (Kcur - 1)) * b.G + idxs] - b.logV[(base + (Kcur - 1)) * b.G + idxs];
  };
  auto colmap = [&](int j, int Kcur) -> int { return (j < Kcur) ? U.col[j] : U.col[KS + (j - Kcur)]; };

  bool live = valid;
  int bail_l = -1, bail_it = 0;

  for (int l = iter0; l < iter0 + n_sweeps; l++) {
    if (live && K > KS) {
      bail_l = l;
      bail_it = 0;
      live = false;
    }
    if (!__any_sync(TPPMX_FULL, live)) break;
    // =========================== parallel pre-pass of the sweep ===========================
    if (live && !R0)  // :351-360 auxiliary eta ~ N(0, I)
      for (int mm = ul; mm < CC; mm += LPU)
        rng_normals(rng, l, TPPMX_SITE_AUX_SWEEP, tt, (uint32_t)mm, D, g_etae + mm, CC);
    __syncwarp();
    if (live && !NOLIK) {  // exp(eta) of every candidate cluster, stored by column id
      const int ntot = R0 ? K : K + CC;  // R0: the auxiliary eta are per patient (phase B)
      for (int e = ul; e < ntot * D; e += LPU) {
        const int c = e / D, k = e - c * D;
        const double eta = (c < K) ? g_eta[(size_t)k * kcap + c] : g_etae[k * CC + (c - K)];
        U.E[colmap(c, K) * D + k] = fast_exp(eta);
      }
    }
    __syncwarp();
    if constexpr (BIG) {
      // long arms: one lane per patient keeps exp(beta'z) / log JJ in registers while it walks over
      // the candidate clusters (one L2 round trip per patient; the arm's rows are not cache resident)
      if (live) {
        const int ntot = K + CC;
        for (int it = ul; it < nt; it += LPU) {
          const int i = patp[it];
          const double *xi = g_x + (size_t)i * P;
          const double *xm = g_x + (size_t)patp[it > 0 ? it - 1 : 0] * P;  // the patient one step earlier
          double qx = 0.0, gx = 0.0;
          for (int p = 0; p < P; p++) {
            qx = fma(xi[p], xi[p], qx);
            gx = fma(xi[p], xm[p], gx);
          }
          double uu, u1;
          rng_u2(rng, l, TPPMX_SITE_ALLOC_U, tt, (uint32_t)i, 0, uu, u1);
          uq[it] = uu;
          uq[uqs + it] = qx;
          uq[2 * uqs + it] = gx;
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
              lik[(size_t)cl * ntp + it] = fast_lik_reg<DR>(U.E + cl * D, Z, ljj, D);
            }
          }
        }
      }
      __syncwarp();
    } else {
      if (live) {
        // phase A, per patient (one lane each): allocation uniform, sum_p x^2, x_{it-1} . x_it, exp(beta'z),
        // log JJ.  Every global load of a patient is issued before the first use (the lane is alone on its
        // dependent chain: a load inside the loop over categories costs a full L2 round trip each time).
        for (int it = ul; it < nt; it += LPU) {
          const int i = patp[it];
          const double *xi = g_x + (size_t)i * P;
          const double *xm = g_x + (size_t)patp[it > 0 ? it - 1 : 0] * P;  // the patient one step earlier
          const double *zi = b.z + ((size_t)ds * b.nobs + i) * Q;
          const double *JJ = b.JJ + ((size_t)chain * b.nobs + i) * D;
          double JJv[DR], acc[DR];
#pragma unroll
          for (int k = 0; k < DR; k++) {
            JJv[k] = (!NOLIK && k < D) ? JJ[k] : 1.0;
            acc[k] = 0.0;
          }
          double qx = 0.0, gx = 0.0;
          if constexpr (PC > 0) {
#pragma unroll
            for (int p = 0; p < PC; p++) {
              const double xv = xi[p];
              qx = fma(xv, xv, qx);
              gx = fma(xv, xm[p], gx);
            }
          } else {
            for (int p = 0; p < P; p++) {
              const double xv = xi[p];
              qx = fma(xv, xv, qx);
              gx = fma(xv, xm[p], gx);
            }
          }
          if constexpr (DDIP) {
            double sxv = 0.0;
            for (int p = 0; p < P; p++) sxv += xi[p];
            uq[3 * uqs + it] = sxv;
          }
          if (!NOLIK)
            for (int q = 0; q < Q; q++) {
              const double zq = zi[q];
#pragma unroll
              for (int k = 0; k < DR; k++)
                if (k < D) acc[k] = fma(beta[k + q * D], zq, acc[k]);
            }
          double uu, u1;
          rng_u2(rng, l, TPPMX_SITE_ALLOC_U, tt, (uint32_t)i, 0, uu, u1);
          uq[it] = uu;
          uq[uqs + it] = qx;
          uq[2 * uqs + it] = gx;
          if (!NOLIK) {
#pragma unroll
            for (int k = 0; k < DR; k++)
              if (k < D) {
                g_ez[(size_t)i * D + k] = fast_exp(acc[k]);
                g_ljj[(size_t)i * D + k] = fast_log(JJv[k]);
              }
          }
        }
      }
      __syncwarp();
      if (live && !NOLIK) {
        // phase B, per (candidate cluster, patient) pair, flattened over the unit's lanes so that
        // every lane carries the same number of likelihood terms whatever n_t and K are (a loop
        // over patients with an inner loop over clusters leaves 20-40 % of the lanes idle):
        // pair e -> (c, it) = (e / nt, e % nt), kept incrementally; consecutive lanes write
        // consecutive rows of one table column.  D <= 3: the operands of the NEXT pair are fetched
        // before the current pair's three lgamma evaluations.
        const int npair = (R0 ? K : K + CC) * nt;
        int c = ul / nt, it = ul - c * nt;
        if constexpr (DR == 3) {
          const bool h1 = D > 1, h2 = D > 2;
          double z0 = 1.0, z1 = 1.0, z2 = 1.0, j0 = 0.0, j1 = 0.0, j2 = 0.0;
          int cl = 0;
          if (ul < npair) {
            const size_t o = (size_t)patp[it] * D;
            cl = colmap(c, K);
            z0 = g_ez[o];
            j0 = g_ljj[o];
            if (h1) { z1 = g_ez[o + 1]; j1 = g_ljj[o + 1]; }
            if (h2) { z2 = g_ez[o + 2]; j2 = g_ljj[o + 2]; }
          }
          for (int e = ul; e < npair; e += LPU) {
            int itn = it + LPU, cn = c;
            while (itn >= nt) {
              itn -= nt;
              cn += 1;
            }
            double y0 = 1.0, y1 = 1.0, y2 = 1.0, m0 = 0.0, m1 = 0.0, m2 = 0.0;
            int cln = 0;
            if (e + LPU < npair) {
              const size_t o = (size_t)patp[itn] * D;
              cln = colmap(cn, K);
              y0 = g_ez[o];
              m0 = g_ljj[o];
              if (h1) { y1 = g_ez[o + 1]; m1 = g_ljj[o + 1]; }
              if (h2) { y2 = g_ez[o + 2]; m2 = g_ljj[o + 2]; }
            }
            {  // inlined (a call would wait for the loads in flight): sum_k gamma_k log JJ_k - lgamma(gamma_k)
              const double *E = U.E + cl * D;
              const double g0 = E[0] * z0;
              const double g1 = h1 ? E[1] * z1 : 1.0;
              const double g2 = h2 ? E[2] * z2 : 1.0;
              double l0, l1, l2;
              lgamma_fast3(g0, g1, g2, l0, l1, l2);
              double w = fma(g0, j0, -l0);
              if (h1) w += fma(g1, j1, -l1);
              if (h2) w += fma(g2, j2, -l2);
              lik[(size_t)cl * ntp + it] = w;
            }
            it = itn; c = cn; cl = cln;
            z0 = y0; z1 = y1; z2 = y2; j0 = m0; j1 = m1; j2 = m2;
          }
        } else {
          for (int e = ul; e < npair; e += LPU) {
            const int i = patp[it];
            const int cl = colmap(c, K);
            lik[(size_t)cl * ntp + it] = fast_lik<DR>(U.E + cl * D, g_ez + (size_t)i * D, g_ljj + (size_t)i * D, D);
            it += LPU;
            while (it >= nt) {
              it -= nt;
              c += 1;
            }
          }
        }
      }
      if constexpr (R0 && DR == 3) {
        if (live && !NOLIK) {  // auxiliary columns: slot m of patient it scores the patient's own draw eta_aux(i, m)
          for (int e = ul; e < CC * nt; e += LPU) {
            const int m = e / nt, it = e - m * nt;
            const int i = patp[it];
            double zn[4], Ea[3];
            rng_normals(rng, l, TPPMX_SITE_AUX_PATIENT, tt, (uint32_t)(i * CC + m), D, zn, 1);
            for (int k = 0; k < 3; k++) Ea[k] = (k < D) ? fast_exp(zn[k]) : 1.0;
            const size_t o = (size_t)i * D;
            lik[(size_t)colmap(K + m, K) * ntp + it] =
                fast_lik3(Ea, D, g_ez[o], D > 1 ? g_ez[o + 1] : 1.0, D > 2 ? g_ez[o + 2] : 1.0, g_ljj[o], D > 1 ? g_ljj[o + 1] : 0.0,
                          D > 2 ? g_ljj[o + 2] : 0.0);
          }
        }
      }
      __syncwarp();
    }

    // =========================== sequential reallocation steps ===========================
    // The step is software-pipelined so that only  draw -> score -> softmax -> draw  is a dependent chain:
    //  * a_j = sum_p t_jp^2 and b_j = sum_p t_jp x_p of patient it+1 are formed during step it, from the
    //    statistics as they stand after the removal of patient it (they do not depend on step it's draw);
    //    the one cluster the draw then changes is corrected in closed form on its own lane,
    //        b += iv2 (x_it . x_it+1),   a += iv2 (2 b_it + iv2 q_it)  [= t_Y of the draw's score],
    //    and so is the next patient's own cluster (b -= iv2 q, a -= iv2 (2 b + iv2 q)); x_it . x_it+1 and
    //    q = sum_p x^2 come from the pre-pass.  Every a, b is re-formed from the exact statistics at
    //    every step, so nothing drifts;
    //  * the exact statistics' read-modify-writes (arrival of patient it-1, removal of patient it) are
    //    merged at the top of step it, off the chain; cluster sizes are carried in the lanes' registers.
    // more candidates than lanes in some unit of the warp?  Changes only with the number of
    // clusters, so the per-step vote is skipped while it cannot be true.
    bool anymulti = __any_sync(TPPMX_FULL, live && K + CC > LPU);
    int labcur = labp[0];
    double dV = dV_of(K);
    bool act0 = ul < K + CC;
    const double *likp = lik + (size_t)(act0 ? colmap(ul, K) : 0) * ntp;
    double likv = 0.0, uu = uq[0], qx = uq[uqs];
    {  // covariate rows of patients 0 and 1 -> ring slots 0 and 1
      const int i0 = patp[0], i1 = patp[1];
#pragma unroll
      for (int r = 0; r < XR; r++) {
        const int p = ul + r * LPU;
        if (p < P) {
          U.xs[p] = g_x[(size_t)i0 * P + p];
          U.xs[Pp + p] = g_x[(size_t)i1 * P + p];
        }
      }
      if (!NOLIK) likv = likp[0];
      if constexpr (CAT)
        for (int p = ul; p < ncat; p += LPU) {
          U.xcs[p] = g_xc[(size_t)i0 * ncat + p];
          U.xcs[ncat + p] = g_xc[(size_t)i1 * ncat + p];
        }
    }
    __syncwarp();
    // a, b of this lane's cluster (lanes >= K: the empty cluster, column 32) against the staged statistics
    auto dots = [&](const double *xv, double &a, double &bb, double &ss) {
      const double *sx = U.sumx + ((ul < K) ? ul : TPPMX_FAST_KSMAX);
      if constexpr (DDIP) {  // sum_p t_jp (returned through ss)
        double q0 = 0.0;
        for (int p = 0; p < P; p++) q0 += fma(sx[p * KSTR], iv2, c0);
        ss = q0;
      }
      if constexpr (CAL1) {  // sum_p sumx2_jp of this lane's cluster (the empty column gives 0)
        const double *sq = U.sumx2 + ((ul < K) ? ul : TPPMX_FAST_KSMAX);
        double q0 = 0.0;
        for (int p = 0; p < P; p++) q0 += sq[p * KSTR];
        ss = q0;
      }
      double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
      if constexpr (PC > 0) {  // compile-time P: straight-line code the scheduler can interleave with the draw's chain
#pragma unroll
        for (int p = 0; p + 1 < PC; p += 2) {
          const double t0 = fma(sx[p * KSTR], iv2, c0), t1 = fma(sx[(p + 1) * KSTR], iv2, c0);
          a0 = fma(t0, t0, a0);
          a1 = fma(t1, t1, a1);
          b0 = fma(t0, xv[p], b0);
          b1 = fma(t1, xv[p + 1], b1);
        }
        if constexpr (PC & 1) {
          const double t0 = fma(sx[(PC - 1) * KSTR], iv2, c0);
          a0 = fma(t0, t0, a0);
          b0 = fma(t0, xv[PC - 1], b0);
        }
      } else {
        int p = 0;
        for (; p + 1 < P; p += 2) {
          const double t0 = fma(sx[p * KSTR], iv2, c0), t1 = fma(sx[(p + 1) * KSTR], iv2, c0);
          a0 = fma(t0, t0, a0);
          a1 = fma(t1, t1, a1);
          b0 = fma(t0, xv[p], b0);
          b1 = fma(t1, xv[p + 1], b1);
        }
        if (p < P) {
          const double t0 = fma(sx[p * KSTR], iv2, c0);
          a0 = fma(t0, t0, a0);
          b0 = fma(t0, xv[p], b0);
        }
      }
      a = a0 + a1;
      bb = b0 + b1;
    };
    int myn = (ul < K) ? U.nj[ul] : 0;  // size of this lane's cluster (0: auxiliary / none)
    int nzi = U.nj[labcur - 1];         // size of the current patient's cluster
    int pend = 0;                       // label whose statistics still miss patient it-1 (0: none)
    double a_c, b_c, ss_c = 0.0;
    dots(U.xs, a_c, b_c, ss_c);
    if (ul == labcur - 1) {  // patient 0 leaves its own cluster
      b_c = fma(-iv2, qx, b_c);
      a_c = fma(-iv2, fma(iv2, qx, 2.0 * b_c), a_c);
      ss_c -= DDIP ? iv2 * uq[3 * uqs] : qx;
    }

    for (int it = 0; it < nt_w; it++) {
      bool on = live && it < nt;
      if (on && K >= KS) {  // a birth could overflow the staged capacity: hand the unit over
        bail_l = l;
        bail_it = it;
        live = false;
        on = false;
      }
      const double *xs = U.xs + (it & 3) * Pp;          // x of patient it
      const double *xs_n = U.xs + ((it + 1) & 3) * Pp;  // x of patient it+1 (staged during step it-1)
      // ---- prefetch: covariate row of patient it+2, scalars of patient it+1 (addresses do not depend
      //      on this step; the arrays are padded so the last steps may read past the arm)
      double xr_n[XR], lik_n = 0.0;
      {
        const double *xrow = g_x + (size_t)patp[min(it + 2, ntp)] * P;
#pragma unroll
        for (int r = 0; r < XR; r++) {
          const int p = ul + r * LPU;
          xr_n[r] = (p < P) ? xrow[p] : 0.0;
        }
      }
      int xc_n = 0;  // lane p < ncat: category of patient it+2 in covariate p
      if constexpr (CAT)
        if (ul < ncat) xc_n = g_xc[(size_t)patp[min(it + 2, ntp)] * ncat + ul];
      const double uu_n = uq[it + 1], qx_n = uq[uqs + it + 1], g_n = uq[2 * uqs + it + 1];
      int lab_n = labp[it + 1];  // label of the next patient (re-read after a relabelling)
      if (!NOLIK) lik_n = likp[it + 1];

      // ---- :368-457 take the patient out of its cluster; :820-856 of the previous step (deferred).
      //      Branch-free: a missing arrival / removal adds 0.0 to the column of the other one, and lanes
      //      beyond P repeat the last covariate's (identical) read-modify-write.
      const int zi = on ? labcur - 1 : 0;
      TPPMX_ASSERT(zi >= 0 && zi < KS && (!on || zi < K), 1);
      TPPMX_ASSERT(!on || nzi >= 1, 2);
      TPPMX_ASSERT(it + 1 <= ntp, 3);
      const bool single = on && nzi <= 1;
      // rows of the tables over n for this lane's cluster after the removal (issued before the statistics'
      // read-modify-writes; re-read below if the partition changes)
      const int n_sc = myn - ((on && !single && ul == zi) ? 1 : 0);
      double2 t01 = tab2[2 * n_sc], t23 = tab2[2 * n_sc + 1];  // (dA, HN), (HY, -); unused by the NNIG instantiation
      double lgn = lognp[n_sc];
      {
        const bool add = pend != 0;        // cluster pend-1 receives patient it-1
        const bool rem = on && !single;    // cluster zi loses patient it
        const int ca = add ? pend - 1 : zi;
        const double *xm = U.xs + ((it + 3) & 3) * Pp;
#pragma unroll
        for (int r = 0; r < XR; r++) {
          if (r * LPU >= P) break;  // warp-uniform
          int p = ul + r * LPU;
          p = p < P ? p : P - 1;
          double *pa1 = U.sumx + p * KSTR + ca, *pa2 = U.sumx2 + p * KSTR + ca;
          double *pr1 = U.sumx + p * KSTR + zi, *pr2 = U.sumx2 + p * KSTR + zi;
          const double sa1 = *pa1, sa2 = *pa2, sr1 = *pr1, sr2 = *pr2;  // all loads before the stores
          const double xa = add ? xm[p] : 0.0, xr = rem ? xs[p] : 0.0;
          const double na1 = __dadd_rn(sa1, xa), na2 = __dadd_rn(sa2, __dmul_rn(xa, xa));
          const bool same = ca == zi;  // remove from what the arrival left
          const double nr1 = __dsub_rn(same ? na1 : sr1, xr), nr2 = __dsub_rn(same ? na2 : sr2, __dmul_rn(xr, xr));
          *pa1 = na1;
          *pa2 = na2;
          *pr1 = nr1;
          *pr2 = nr2;
        }
        if constexpr (CAT)
          if (ul < ncat) {  // the categorical counts follow the same two moves (integers: any order)
            if (add) U.njc[(ul * maxC + U.xcs[((it + 3) & 3) * ncat + ul]) * KSTR + ca] += 1;
            if (rem) U.njc[(ul * maxC + U.xcs[(it & 3) * ncat + ul]) * KSTR + zi] -= 1;
          }
        // cluster sizes: every lane keeps its cluster's size in a register (the arrival was counted when
        // it was drawn) and stores it; clusters beyond the lanes (K > LPU, rare) are updated by lane 0
        myn = n_sc;
        if (ul < KS) U.nj[ul] = myn;
        if (LPU < TPPMX_FAST_KSMAX && anymulti && ul == 0) {
          if (add && ca >= LPU) U.nj[ca] += 1;
          if (rem && zi >= LPU) U.nj[zi] -= 1;
        }
      }
      pend = 0;
      __syncwarp();
      if (__builtin_expect(__any_sync(TPPMX_FULL, single), 0)) {  // singleton (:384-457), rare
        const int iaux = zi + 1, Kc = K;
        const bool swp = single && iaux < Kc;
        if (swp)  // swap with the last cluster
          for (int ii = ul; ii < nt; ii += LPU)
            if (labp[ii] == Kc) labp[ii] = iaux;
        __syncwarp();
        if (swp) {
          if (ul == 0) {
            labp[it] = Kc;
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
          if constexpr (CAT)
            for (int pc = ul; pc < ncC; pc += LPU) {
              const int t = U.njc[pc * KSTR + iaux - 1];
              U.njc[pc * KSTR + iaux - 1] = U.njc[pc * KSTR + Kc - 1];
              U.njc[pc * KSTR + Kc - 1] = t;
            }
        }
        __syncwarp();
        if (single) {  // empty slot Kc-1 (keeps the rounding residue, as the reference does)
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
          if constexpr (CAT)
            if (ul < ncat) U.njc[(ul * maxC + U.xcs[(it & 3) * ncat + ul]) * KSTR + Kc - 1] -= 1;
          K = Kc - 1;
          dV = dV_of(K);
        }
        anymulti = __any_sync(TPPMX_FULL, live && K + CC > LPU);
        __syncwarp();
        // the lanes of a unit whose partition changed re-form their a, b, n from the statistics
        lab_n = labp[it + 1];
        double ar, br, sr = 0.0;
        dots(xs, ar, br, sr);
        if (single) {
          a_c = ar;
          b_c = br;
          ss_c = sr;
          myn = (ul < K) ? U.nj[ul] : 0;
          t01 = tab2[2 * myn];
          t23 = tab2[2 * myn + 1];
          lgn = lognp[myn];
          act0 = ul < K + CC;
          likp = lik + (size_t)(act0 ? colmap(ul, K) : 0) * ntp;
          TPPMX_ASSERT(!act0 || (colmap(ul, K) >= 0 && colmap(ul, K) < fa.ncol), 6);
          TPPMX_ASSERT(nfree >= 0 && nfree <= fa.ncol, 7);
          if (!NOLIK) {
            likv = likp[it];
            lik_n = likp[it + 1];
          }
        }
      }

      // ---- :460-814 score the K occupied + CC auxiliary clusters, normalise, draw
      const int ntot = K + CC;
      int newci;
      const double t1 = fma(iv2, qx, 2.0 * b_c);
      double w, wmax;
      if constexpr (NNIG) {
        // rows myn ("without") and myn + 1 ("with") of the table over n (tppmx.cu:fast_reserve)
        const bool occ = ul < K;
        const double *rN = fa.tabg + (size_t)TPPMX_NNIG_ROW * myn, *rY = rN + TPPMX_NNIG_ROW;
        const double *sx = U.sumx + (occ ? ul : TPPMX_FAST_KSMAX), *sq = U.sumx2 + (occ ? ul : TPPMX_FAST_KSMAX);
        const double k0m0 = __dmul_rn(s.k0, s.m0), b0 = __dmul_rn(__dmul_rn(0.5, s.nu0), s.s20);
        double lgN = 0.0, lgY = 0.0;
        for (int p = 0; p < P; p++) {
          const double S = sx[p * KSTR], SS = sq[p * KSTR], xv = xs[p];
          lgN = __dadd_rn(lgN, gsim_nnig_row(rN, S, SS, s.m0, k0m0, b0));
          lgY = __dadd_rn(lgY, gsim_nnig_row(rY, __dadd_rn(S, xv), __dadd_rn(SS, __dmul_rn(xv, xv)), s.m0, k0m0, b0));
        }
        if (!occ) lgN = 0.0;
        const double wv = dV + lgn + simscale * (lgY - lgN) + likv;
        w = act0 ? wv : -INFINITY;
        wmax = unit_shift<LPU>(w, sub);
      } else if constexpr (DDIP) {
        const double sxc = uq[3 * uqs + it];  // sum_p x of this patient (pre-pass)
        const double tY = fma(iv2, t1, a_c);
        const double au = 4.0 * a_c - 4.0 * c0 * ss_c + (double)P * c0 * c0;   // sum_p u^2 without the patient
        const double bu = 2.0 * b_c - c0 * sxc;                                 // sum_p u x
        const double uY = au + 4.0 * iv2 * (bu + iv2 * qx);                     // ... with it
        const double h2y = tab2[2 * (myn + 1) + 1].y;                           // H2[n + 1]; t23.y = H2[n]
        const double d = t01.x - hiv2 * qx - (t23.x * tY - t01.y * a_c) + (h2y * uY - t23.y * au);
        const double wv = dV + lgn + simscale * d + likv;
        w = act0 ? wv : -INFINITY;
        wmax = unit_shift<LPU>(w, sub);
      } else if constexpr (!CAL1) {
        const double tY = fma(iv2, t1, a_c);
        double d = t01.x - hiv2 * qx + (t23.x * tY - t01.y * a_c);
        if constexpr (CAT) {  // + sum_p [ log(n_jc* + w) - log(n_j + C_p w) ], counts after the removal
          const int colj = (ul < K) ? ul : TPPMX_FAST_KSMAX;
          const int *xc = U.xcs + (it & 3) * ncat;
          for (int p = 0; p < ncat; p++)
            d += fa.cat_lw[U.njc[(p * maxC + xc[p]) * KSTR + colj]] - fa.cat_ln[p * (b.ntmax + 2) + myn];
        }
        const double wv = dV + lgn + simscale * d + likv;
        w = act0 ? wv : -INFINITY;
        wmax = unit_shift<LPU>(w, sub);
      } else {
        const bool occ = ul < K;
        const double tY = fma(iv2, t1, a_c);
        const double lgN = t23.y - hiv2 * ss_c + t01.y * a_c;                     // P A0[n] - ... + H[n] a
        const double lgY = (t23.y + t01.x) - hiv2 * (ss_c + qx) + t23.x * tY;     // P A0[n+1] - ... + H[n+1] t_Y
        const double gN = act0 ? (occ ? lgN : lgY) : -INFINITY;                   // :670-671 auxiliary: the singleton value
        const double gY = occ ? lgY : -INFINITY;
        const double mN = unit_shift<LPU>(gN, sub), mY = unit_shift<LPU>(gY, sub);
        const double sN = unit_sum<LPU>(fast_exp(gN - mN)), sY = unit_sum<LPU>(fast_exp(gY - mY));
        const double ltN = (gN - mN) - fast_log(sN);                              // log gtilN
        const double ltY = (gY - mY) - fast_log(sY);                              // log gtilY (occupied)
        const double wv = dV + lgn + likv + (occ ? ltY - ltN : ltN);
        w = act0 ? wv : -INFINITY;
        wmax = unit_shift<LPU>(w, sub);
      }
      // ---- a, b of patient it+1 against the statistics without patient it: independent of this step's
      //      draw; placed here so that the scheduler interleaves it with the exp / scan chain below.
      //      The next patient's own cluster is corrected for its removal.
      const int zn = (it + 1 < nt) ? lab_n - 1 : 0;
      double a_n, b_n, ss_n = 0.0;
      dots(xs_n, a_n, b_n, ss_n);
      {
        const double bs = fma(-iv2, qx_n, b_n);
        const double as = fma(-iv2, fma(iv2, qx_n, 2.0 * bs), a_n);
        b_n = (ul == zn) ? bs : b_n;
        a_n = (ul == zn) ? as : a_n;
        if constexpr (CAL1) ss_n = (ul == zn) ? ss_n - qx_n : ss_n;
        if constexpr (DDIP) ss_n = (ul == zn) ? ss_n - iv2 * uq[3 * uqs + it + 1] : ss_n;
      }
      const int nj_n = U.nj[zn];  // size of the next patient's cluster before this step's arrival
      {
        const double e0 = fast_exp(w - wmax);  // 0 for w = -inf
        const double cum = unit_incl_scan<LPU>(e0, ul);
        const double den = __shfl_sync(TPPMX_FULL, cum, LPU - 1, LPU);
        // uu < cumsum(e)/den  <=>  uu*den < cumsum(e)   (:805-814)
        unsigned hit = __ballot_sync(TPPMX_FULL, act0 && (uu * den < cum));
        hit = (LPU == 32) ? hit : ((hit >> (sub * LPU)) & ((1u << LPU) - 1u));
        newci = hit ? __ffs(hit) : ntot;  // :807 default when rounding leaves uu >= total
        if constexpr (PROBE) {
          const int pu = unit - fa.probe_unit0;
          if (on && pu >= 0 && pu < fa.probe_nunits) {
            if (act0 && ul < fa.probe_stride) fa.probe[((size_t)pu * ntp + it) * fa.probe_stride + ul] = e0 / den;
            if (ul == 0) fa.probe_k[(size_t)pu * ntp + it] = K;
          }
        }
      }
      const bool multi = on && ntot > LPU;
      if (!CAL1 && !NNIG && !CAT && !DDIP && __builtin_expect(anymulti && __any_sync(TPPMX_FULL, multi), 0)) {
        double *prow = nullptr;
        if constexpr (PROBE) {
          const int pu = unit - fa.probe_unit0;
          if (pu >= 0 && pu < fa.probe_nunits) prow = fa.probe + ((size_t)pu * ntp + it) * fa.probe_stride;
        }
        const int r = fast_score_multi<LPU>(U, tab2, lognp, lik, ntp, it, K, CC, KS, P, xs, qx, uu, dV,
                                            simscale, iv2, c0, hiv2, NOLIK, multi, ul, sub, prow, fa.probe_stride);
        if (multi) newci = r;
      }

      if constexpr (R0) {  // the state the arm leaves behind: eta_empty holds the LAST patient's draws (:605-610)
        if (on && it == nt - 1 && ul < CC)
          rng_normals(rng, l, TPPMX_SITE_AUX_PATIENT, tt, (uint32_t)(patp[it] * CC + ul), D, g_etae + ul, CC);
      }
      // ---- :820-856 assign; the statistics receive the patient at the top of the next step
      TPPMX_ASSERT(!on || (newci >= 1 && newci <= K + CC), 4);
      const bool born = on && newci > K;
      TPPMX_ASSERT(!born || (nfree >= 1 && K + 1 <= KS), 5);
      int lab = newci;
      if (on && !born && ul == 0) labp[it] = lab;
      if (__builtin_expect(__any_sync(TPPMX_FULL, born), 0)) {  // a new cluster takes the auxiliary eta (:829-845), rare
        int m = 0, newcol = 0, i = 0;
        if (born) {
          i = patp[it];
          m = newci - K - 1;
          lab = K + 1;
          newcol = U.freec[nfree - 1];
          nfree -= 1;
        }
        __syncwarp();
        if (born) {
          if (ul == 0) {
            labp[it] = lab;
            U.nj[lab - 1] = 0;  // the arrival is counted with the statistics, at the top of the next step
            if constexpr (R0) {
              U.col[lab - 1] = newcol;  // the auxiliary slot keeps its (patient-specific) column
              rng_normals(rng, l, TPPMX_SITE_AUX_PATIENT, tt, (uint32_t)(i * CC + m), D, g_etae + m, CC);  // this patient's draw
            } else {
              U.col[lab - 1] = U.col[KS + m];
              U.col[KS + m] = newcol;
            }
          }
        }
        if constexpr (R0) __syncwarp();
        if (born) {
          for (int k = ul; k < D; k += LPU) g_eta[(size_t)k * kcap + lab - 1] = g_etae[k * CC + m];
        }
        __syncwarp();
        if (born && ul == 0) rng_normals(rng, l, TPPMX_SITE_AUX_REFRESH, tt, (uint32_t)i, D, g_etae + m, CC);
        __syncwarp();
        if (born && !NOLIK)  // the column filled below: the refreshed auxiliary eta (reuse) / the new cluster's own eta (R0)
          for (int k = ul; k < D; k += LPU) U.E[newcol * D + k] = fast_exp(R0 ? g_eta[(size_t)k * kcap + lab - 1] : g_etae[k * CC + m]);
        __syncwarp();
        if (born) {
          if (!NOLIK)  // likelihood column of the refreshed auxiliary eta, remaining patients
            for (int it2 = it + 1 + ul; it2 < nt; it2 += LPU) {
              const int i2 = patp[it2];
              lik[(size_t)newcol * ntp + it2] =
                  fast_lik<DR>(U.E + newcol * D, g_ez + (size_t)i2 * D, g_ljj + (size_t)i2 * D, D);
            }
          K = lab;
          dV = dV_of(K);
        }
        anymulti = __any_sync(TPPMX_FULL, live && K + CC > LPU);
        __syncwarp();
        if (born) {  // lanes change roles: likelihood column of each lane, next patient's entry
          act0 = ul < K + CC;
          likp = lik + (size_t)(act0 ? colmap(ul, K) : 0) * ntp;
          if (!NOLIK) lik_n = likp[it + 1];
        }
      }
      // ---- the lane of the cluster that receives patient it: its size, and a, b of patient it+1
      //      (a lane >= the old K held the empty cluster's values: a birth needs nothing else)
      {
        const bool mine = on && ul == lab - 1;
        const double bq = fma(iv2, g_n, b_n);
        const double aq = fma(iv2, t1 - ((ul == zn) ? 2.0 * iv2 * g_n : 0.0), a_n);
        myn += mine ? 1 : 0;
        b_n = mine ? bq : b_n;
        a_n = mine ? aq : a_n;
        if constexpr (CAL1) ss_n = mine ? ss_n + qx : ss_n;
        if constexpr (DDIP) ss_n = mine ? ss_n + iv2 * uq[3 * uqs + it] : ss_n;
      }
      pend = on ? lab : 0;
      nzi = nj_n + ((on && lab - 1 == zn) ? 1 : 0);
      // ---- stage the covariate row of patient it+2, rotate the operands
#pragma unroll
      for (int r = 0; r < XR; r++) {
        const int p = ul + r * LPU;
        if (p < P) U.xs[((it + 2) & 3) * Pp + p] = xr_n[r];
      }
      if constexpr (CAT)
        if (ul < ncat) U.xcs[((it + 2) & 3) * ncat + ul] = xc_n;
      uu = uu_n;
      qx = qx_n;
      likv = lik_n;
      labcur = lab_n;
      a_c = a_n;
      b_c = b_n;
      ss_c = ss_n;
      __syncwarp();
    }
    // ---- the statistics still miss the last patient of the arm
    {
      const int ca = pend - 1;
      const double *xm = U.xs + ((nt_w + 3) & 3) * Pp;
      if (ca >= 0) {
#pragma unroll
        for (int r = 0; r < XR; r++) {
          const int p = ul + r * LPU;
          if (p < P) {
            const double xv = xm[p], s1 = U.sumx[p * KSTR + ca], s2 = U.sumx2[p * KSTR + ca];
            U.sumx[p * KSTR + ca] = __dadd_rn(s1, xv);
            U.sumx2[p * KSTR + ca] = __dadd_rn(s2, __dmul_rn(xv, xv));
          }
        }
        if (ul == 0) U.nj[ca] += 1;
        if constexpr (CAT)
          if (ul < ncat) U.njc[(ul * maxC + U.xcs[((nt_w + 3) & 3) * ncat + ul]) * KSTR + ca] += 1;
      }
      __syncwarp();
    }
  }

  // ---- write the unit back in the library's global layout
  __syncwarp();
  if (valid) {
    if (s.PPMx)
      for (int e = ul; e < P * KSTR; e += LPU) {
        const int p = e / KSTR, j = e % KSTR;
        if (j < KS) {
          g_sumx[(size_t)p * kcap + j] = U.sumx[e];
          g_sumx2[(size_t)p * kcap + j] = U.sumx2[e];
        }
      }
    for (int j = ul; j < KS; j += LPU) g_nj[j] = U.nj[j];
    if constexpr (CAT)
      for (int e = ul; e < ncC * KSTR; e += LPU) {
        const int pc = e / KSTR, j = e % KSTR;
        if (j < KS) g_njc[(size_t)pc * kcap + j] = U.njc[e];
      }
    if (!BIG)
      for (int it = ul; it < nt; it += LPU) g_lab[it] = labp[it];
    if (ul == 0) {
      b.nclu[(size_t)chain * b.T + tt] = K;
      fa.resume_l[unit] = bail_l;
      fa.resume_it[unit] = bail_it;
      if (bail_l >= 0) atomicAdd(fa.bail_count, 1);
      // long arms: the next launch sizes its staging from it; special instantiations: the host picks 16 | 32 lanes from it
      if ((BIG || CAL1 || NNIG || CAT || DDIP) && fa.kmax_dev) atomicMax(fa.kmax_dev, K);
    }
  }
  __syncwarp();
  // loggamma (:825, :839): eta* is constant during the sweep, so the value every patient ends
  // with is eta*[label] + beta'z, evaluated here once in the reference's operation order
  // (utils_ct.cpp:514-527).  A unit that handed over leaves this to the generic kernel for
  // the patients it still has to visit; rewriting the visited ones is value-preserving.
  if (valid && n_sweeps > 0) {
    for (int it = ul; it < nt; it += LPU) {
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

// ---- launch: instantiation picked from the batch shape (one translation unit per LPU) ----------
template <int LPU, int XR, int DR, bool NOLIK, bool BIG, bool PROBE, int PC, bool CAL1 = false, bool NNIG = false, bool CAT = false,
          bool DDIP = false, bool R0 = false>
static cudaError_t fast_launch7(const DevBatch &d, const FastArgs &fa, size_t smem, uint64_t seed, int iter0,
                                int n_sweeps, cudaStream_t stream) {
  cudaError_t e = cudaFuncSetAttribute(sweep_fast_kernel<LPU, XR, DR, NOLIK, BIG, PROBE, PC, CAL1, NNIG, CAT, DDIP, R0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const int upw = 32 / LPU;
  const int grid = (fa.n_units + upw - 1) / upw;
  sweep_fast_kernel<LPU, XR, DR, NOLIK, BIG, PROBE, PC, CAL1, NNIG, CAT, DDIP, R0><<<grid, 32, smem, stream>>>(d, fa, seed, iter0, n_sweeps);
  return cudaGetLastError();
}
template <int LPU, int XR, int DR, bool NOLIK, bool BIG, bool PROBE>
static cudaError_t fast_launch6(const DevBatch &d, const FastArgs &fa, size_t smem, uint64_t seed, int iter0,
                                int n_sweeps, cudaStream_t stream) {
  // the package's shape (genmech: 10 predictive covariates, D = 3 outcome categories, two units per warp) is
  // compiled with P as a constant: its dot products unroll into straight-line code
  if (d.model.reuse == 0) {  // auxiliary eta redrawn per patient: the default similarity only (tppmx.cu:fast_eligible)
    if constexpr (DR == 3 && !NOLIK && !BIG) {
      if (d.model.similarity != 1 || d.model.consim != 1 || d.model.calibration == 1 || d.ncat > 0) return cudaErrorInvalidConfiguration;
      return fast_launch7<LPU, XR, DR, NOLIK, BIG, PROBE, 0, false, false, false, false, true>(d, fa, smem, seed, iter0, n_sweeps, stream);
    }
    return cudaErrorInvalidConfiguration;
  }
  if (d.model.similarity == 2) {  // double dipper (Normal-Normal, no categorical covariates, calibration 0 | 2)
    if constexpr (LPU >= 16 && DR == 3 && !NOLIK && !BIG) {
      if (fa.KS + d.CC > LPU || d.model.calibration == 1 || d.model.consim != 1 || d.ncat > 0) return cudaErrorInvalidConfiguration;
      return fast_launch7<LPU, XR, DR, NOLIK, BIG, PROBE, 0, false, false, false, true>(d, fa, smem, seed, iter0, n_sweeps, stream);
    }
    return cudaErrorInvalidConfiguration;
  }
  if (d.ncat > 0) {  // categorical covariates beside the continuous ones (Normal-Normal, calibration 0 | 2)
    if constexpr (LPU >= 16 && DR == 3 && !NOLIK && !BIG) {
      if (fa.KS + d.CC > LPU || d.model.calibration == 1 || d.model.consim != 1 || d.ncat > LPU) return cudaErrorInvalidConfiguration;
      return fast_launch7<LPU, XR, DR, NOLIK, BIG, PROBE, 0, false, false, true>(d, fa, smem, seed, iter0, n_sweeps, stream);
    }
    return cudaErrorInvalidConfiguration;
  }
  if (d.model.consim == 2) {  // NNIG similarity: the same restrictions as calibration 1 (tppmx.cu:fast_eligible)
    if constexpr (LPU >= 16 && DR == 3 && !NOLIK && !BIG) {
      if (fa.KS + d.CC > LPU || d.model.calibration == 1) return cudaErrorInvalidConfiguration;
      return fast_launch7<LPU, XR, DR, NOLIK, BIG, PROBE, 0, false, true>(d, fa, smem, seed, iter0, n_sweeps, stream);
    }
    return cudaErrorInvalidConfiguration;
  }
  if (d.model.calibration == 1) {  // normalised similarities: instantiated for the package's case (D <= 3, short arms, 16 | 32 lanes)
    if constexpr (LPU >= 16 && DR == 3 && !NOLIK && !BIG) {
      if (fa.KS + d.CC > LPU) return cudaErrorInvalidConfiguration;
      return fast_launch7<LPU, XR, DR, NOLIK, BIG, PROBE, 0, true>(d, fa, smem, seed, iter0, n_sweeps, stream);
    }
    return cudaErrorInvalidConfiguration;
  }
  if constexpr ((LPU == 16 || LPU == 32) && XR == 1 && DR == 3 && !NOLIK && !BIG)
    if (d.P == 10) return fast_launch7<LPU, XR, DR, NOLIK, BIG, PROBE, 10>(d, fa, smem, seed, iter0, n_sweeps, stream);
  return fast_launch7<LPU, XR, DR, NOLIK, BIG, PROBE, 0>(d, fa, smem, seed, iter0, n_sweeps, stream);
}
template <int LPU, int XR, int DR, bool NOLIK, bool BIG>
static cudaError_t fast_launch5(const DevBatch &d, const FastArgs &fa, size_t smem, uint64_t seed, int iter0,
                                int n_sweeps, cudaStream_t stream) {
  if (fa.probe) {  // parity probe: instantiated for the package's case (D <= 3, likelihood on) only
    if constexpr (DR == 3 && !NOLIK) return fast_launch6<LPU, XR, DR, NOLIK, BIG, true>(d, fa, smem, seed, iter0, n_sweeps, stream);
    return cudaErrorInvalidConfiguration;
  }
  return fast_launch6<LPU, XR, DR, NOLIK, BIG, false>(d, fa, smem, seed, iter0, n_sweeps, stream);
}
template <int LPU, int XR, int DR, bool NOLIK>
static cudaError_t fast_launch4(const DevBatch &d, const FastArgs &fa, size_t smem, uint64_t seed, int iter0,
                                int n_sweeps, cudaStream_t stream) {
  if (fa.big) {  // long arms, first stage (<= 32 clusters per arm): one unit per warp only; units that outgrow it
                 // are finished by sweep_big_kernel (sweep_big.cuh) from the exact position
    if constexpr (LPU == 32) return fast_launch5<LPU, XR, DR, NOLIK, true>(d, fa, smem, seed, iter0, n_sweeps, stream);
    return cudaErrorInvalidConfiguration;
  }
  return fast_launch5<LPU, XR, DR, NOLIK, false>(d, fa, smem, seed, iter0, n_sweeps, stream);
}
template <int LPU, int XR>
static cudaError_t fast_launch2(const DevBatch &d, const FastArgs &fa, size_t smem, uint64_t seed, int iter0,
                                int n_sweeps, cudaStream_t stream) {
  const bool nolik = d.model.nolik != 0;
  // DR = registers of the pre-pass that hold a patient's exp(zb) / log JJ (D <= 3 is the package's case)
  if (d.D <= 3)
    return nolik ? fast_launch4<LPU, XR, 3, true>(d, fa, smem, seed, iter0, n_sweeps, stream)
                 : fast_launch4<LPU, XR, 3, false>(d, fa, smem, seed, iter0, n_sweeps, stream);
  return nolik ? fast_launch4<LPU, XR, TPPMX_MAXD, true>(d, fa, smem, seed, iter0, n_sweeps, stream)
               : fast_launch4<LPU, XR, TPPMX_MAXD, false>(d, fa, smem, seed, iter0, n_sweeps, stream);
}
template <int LPU>
static cudaError_t fast_launch_impl(const DevBatch &d, const FastArgs &fa, size_t smem, uint64_t seed, int iter0,
                                    int n_sweeps, cudaStream_t stream) {
  // XR = registers per lane that hold the next patient's covariate row (P <= XR * LPU)
  if (d.P <= LPU) return fast_launch2<LPU, 1>(d, fa, smem, seed, iter0, n_sweeps, stream);
  return fast_launch2<LPU, 4>(d, fa, smem, seed, iter0, n_sweeps, stream);
}

}  // namespace tppmx
