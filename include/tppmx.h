/*
 * tppmx.h — C ABI of the B200-native PPMx reallocation-sweep / predictive library.
 *
 * This is the drop-in boundary for the hot path of mattpedone/treatppmx:
 *   - collapsed-Gibbs cluster reallocation sweep   reference src/dm_ppmx_ct.cpp:347-860
 *   - posterior-predictive step for new patients   reference src/dm_ppmx_ct.cpp:1080-1385
 *   - similarity / linear-predictor helpers        reference src/utils_ct.cpp:161-200, 382-535
 * The reference reaches that code through ONE .Call entry,
 *   _treatppmx_dm_ppmx_ct (42 SEXPs)              reference src/RcppExports.cpp:16-65, :187
 * An Rcpp shim (treatppmx_b200/csrc/rcpp_shim/) keeps that symbol and forwards to
 * tppmx_dm_ppmx_ct() below; see INTEGRATION.md.
 *
 * Conventions
 *   - plain C: ints, doubles, caller-owned pointers.  The library never frees or keeps
 *     caller memory; all device memory is owned by tppmx_ctx / tppmx_batch.
 *   - every function returns 0 on success, non-zero on error; tppmx_last_error() gives text.
 *   - "col-major" = Armadillo / R layout, so the shim passes .memptr() without copies.
 *   - host-facing chain state uses the REFERENCE's layouts (labels 1-based, cluster-major
 *     sufficient statistics); the device layout (SoA) is private to the library.
 */
#ifndef TPPMX_H
#define TPPMX_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TPPMX_VERSION 100          /* 0.1.0 */
#define TPPMX_MAXD 8               /* outcome categories  (reference dim = y.n_cols, :46)   */
#define TPPMX_MAXQ 16              /* prognostic covariates (reference Q = z.n_cols, :47)   */
#define TPPMX_MAXCC 8              /* auxiliary clusters  (reference CC, R/dm_ppmx_ct.R:93) */
#define TPPMX_MAXT 16              /* treatment arms      (reference A / nT)                */

/* ---- status codes ------------------------------------------------------------------- */
enum {
  TPPMX_OK = 0,
  TPPMX_ERR_ARG = 1,        /* bad argument / unsupported switch combination            */
  TPPMX_ERR_CUDA = 2,       /* CUDA runtime error (text in tppmx_last_error)             */
  TPPMX_ERR_CAPACITY = 3,   /* a chain needed more than kcap clusters in one arm         */
  TPPMX_ERR_NOMEM = 4,
  TPPMX_ERR_INTERRUPTED = 5 /* the caller's interrupt flag was raised (tppmx_set_interrupt_flag) */
};

/* ---- RNG sites: counter-based Philox4x32-10 ------------------------------------------
 * The reference draws from R's global RNG (RNGScope, src/RcppExports.cpp:19).  Here every
 * random number is a pure function of (seed, chain, iteration, site, arm, index, sub):
 *   key     = (seed_lo, seed_hi)
 *   counter = (index, sub | site<<16 | arm<<24, iteration, chain)
 * One Philox block gives two uniforms in (0,1) (52-bit, never 0 or 1) or, by Box-Muller,
 * two standard normals.  The CPU oracle (oracle/) implements the same mapping, which is
 * what makes lock-step parity tests possible.  The shim derives `seed` from two
 * unif_rand() draws so set.seed() still governs a run. */
enum {
  TPPMX_SITE_INIT_SS = 1,     /* ss_i ~ Gamma(1, 1/TT_i)               :88-91   index=i          */
  TPPMX_SITE_INIT_ETA = 2,    /* eta_star ~ U(0,1)                     :140     index=j*D+k, arm */
  TPPMX_SITE_INIT_EMPTY = 3,  /* eta_empty ~ U(0,1)                    :210     index=mm*D+k,arm */
  TPPMX_SITE_AUX_SWEEP = 4,   /* reuse==1: CC aux eta ~ N(0,I)         :351-360 index=mm, arm    */
  TPPMX_SITE_AUX_PATIENT = 5, /* reuse==0: redraw per patient          :605-610 index=i*CC+mm    */
  TPPMX_SITE_ALLOC_U = 6,     /* uu for the reallocation draw          :805     index=i, arm     */
  TPPMX_SITE_AUX_REFRESH = 7, /* refresh the aux that was consumed     :844     index=i, arm     */
  TPPMX_SITE_ETA_PROP = 8,    /* eta_update proposal + accept  utils_ct.cpp:642-676 index=j, arm */
  TPPMX_SITE_HIER_W = 9,      /* Wishart draw of Lambda                :895     arm              */
  TPPMX_SITE_HIER_THETA = 10, /* Theta ~ N(fptp, sptp)                 :899     arm              */
  TPPMX_SITE_SIGKAP = 11,     /* (sigma,kappa) grid draw               :1000    arm              */
  TPPMX_SITE_BETA = 12,       /* beta_update proposal + accept utils_ct.cpp:761-781 index=k*Q+q  */
  TPPMX_SITE_HS_LAMBDA = 13,  /* horseshoe lambda slice        utils_ct.cpp:876-894 index=h      */
  TPPMX_SITE_HS_TAU = 14,     /* horseshoe tau slice           utils_ct.cpp:896-915              */
  TPPMX_SITE_JJ = 15,         /* JJ_ik ~ Gamma(y+exp(lg), 1/(ss+1))    :1045    index=i*D+k      */
  TPPMX_SITE_SS = 16,         /* ss_i ~ Gamma(1, 1/TT_i)               :1054    index=i          */
  TPPMX_SITE_ISPRED = 17,     /* in-sample multinomial draw            :1065    index=i          */
  TPPMX_SITE_PRED_U = 18,     /* predictive cluster draw               :1351    index=pp, arm    */
  TPPMX_SITE_PRED_ETA = 19,   /* eta_pred ~ N(Theta_t, Sigma_t)        :1369    index=pp, arm    */
  TPPMX_SITE_PRED_Y = 20      /* ypred ~ Multinomial(1, pi)            :1381    index=pp, arm    */
};

/* ---- model switches and hyper-parameters: scalar arguments of dm_ppmx_ct (:10-23) ---- */
typedef struct tppmx_model {
  int32_t PPMx;          /* 1: covariate-dependent partition model, 0: plain PPM           */
  int32_t cohesion;      /* 1: DP (log n_j / log alpha - log CC), 2: NGG (V table)  :559-565 */
  int32_t similarity;    /* 1: auxiliary, 2: double dipper                         :472-516 */
  int32_t consim;        /* 1: Normal-Normal, 2: Normal-NormalInverseGamma                  */
  int32_t calibration;   /* 0: none, 1: standardised similarities :717-788, 2: coarsened    */
  int32_t coardegree;    /* 1: 1/(P+ncat), 2: 1/sqrt(P+ncat)                       :588-593 */
  int32_t reuse;         /* 1: Favaro-Teh reuse of auxiliary draws, 0: redraw per patient   */
  int32_t CC;            /* number of auxiliary clusters (<= TPPMX_MAXCC)                   */
  int32_t upd_hier;      /* update (Theta,Sigma) hierarchy                         :880     */
  int32_t hsp;           /* horseshoe prior on beta                                :1029    */
  int32_t noprog;        /* 1: no prognostic covariates, beta frozen               :1009    */
  int32_t nolik;         /* 1: drop the response-likelihood term (:570-576) = the prior-only  */
                         /* sampler prior_ppmx_core (src/prior_ppmx.cpp:101-460); default 0    */
  double alpha;          /* DP mass (reference `alpha`, R `kappa`)                          */
  double similparam[7];  /* m0, s20, v, k0, nu0, (unused), dirichlet weight        :220-227 */
  double hP0_mu0[TPPMX_MAXD];
  double hP0_nu0, hP0_s0, hP0_Lambda0;
  double mhtune_eta, mhtune_beta;   /* mhtunepar(0), mhtunepar(1)                            */
} tppmx_model;

typedef struct tppmx_dims {
  int32_t nobs;    /* n: historical patients                                                 */
  int32_t T;       /* arms                                                                   */
  int32_t D;       /* outcome categories                                                     */
  int32_t Q;       /* prognostic covariates                                                  */
  int32_t P;       /* continuous predictive covariates (ncon)                                */
  int32_t ncat;    /* categorical predictive covariates                                      */
  int32_t maxC;    /* max_C = max(catvec) (:107); 0 if ncat==0                               */
  int32_t npred;   /* new patients                                                           */
  int32_t G;       /* (sigma,kappa) grid points = grid.n_cols (:56); >=1                     */
  int32_t ntmax;   /* max_t n_t = leading cluster dimension of the reference's state arrays  */
  int32_t kcap;    /* device cluster capacity per arm, 1..ntmax (0 => ntmax).  If a chain    */
                   /* ever needs more, the call returns TPPMX_ERR_CAPACITY.                  */
  int32_t reserved0;
} tppmx_dims;

/* Inputs shared by every dataset of a batch. */
typedef struct tppmx_shared {
  const int32_t *catvec;  /* [ncat] categories per categorical covariate                     */
  const double *grid;     /* 2 x G col-major: grid[2g]=sigma_g, grid[2g+1]=kappa_g  (:52-55) */
  const double *logV;     /* compact NGG table [T][2][ntmax][G]:                              */
                          /*   logV[((t*2+r)*ntmax + (K-1))*G + g] = log V(n_t-1+r, K; g)     */
                          /* i.e. r=1 is log Vwm(n_t-1,K-1,g), r=0 is log Vwm(n_t-2,K-1,g)    */
                          /* of the reference's dense cube (:563).  NULL if cohesion==1.      */
} tppmx_shared;

/* One dataset (replicate).  Layouts are the ones dm_ppmx_ct receives (:10-23). */
typedef struct tppmx_dataset {
  const int32_t *treat;   /* [nobs] arm of each patient, 0-based (`treatments`)              */
  const double *y;        /* nobs x D col-major, one-hot outcome                              */
  const double *z;        /* nobs x Q col-major, prognostic covariates                        */
  const double *xcon;     /* nobs x P ROW-major (xcon[i*P+p], R: as.vector(t(xcon)))          */
  const int32_t *xcat;    /* nobs x ncat row-major, categories 0-based                        */
  const double *zpred;    /* npred x Q col-major                                              */
  const double *xconp;    /* npred x P row-major                                              */
  const int32_t *xcatp;   /* npred x ncat row-major                                           */
} tppmx_dataset;

/* Chain state in the REFERENCE's layouts (all caller-owned host memory).
 * Used to upload / download one chain; any pointer may be NULL to skip that field. */
typedef struct tppmx_state {
  int32_t *labels;     /* T x ntmax col-major: curr_clu(t,it), 1-based (:113)                */
  int32_t *nj;         /* T x ntmax col-major: nj_curr(t,j) (:114)                           */
  int32_t *nclu;       /* [T] nclu_curr (:115)                                               */
  double *sumx;        /* T x (ntmax*P) col-major: sumx(t, j*P+p) (:179)                     */
  double *sumx2;       /* same (:180)                                                        */
  int32_t *njc;        /* T x (ntmax*ncat*maxC) col-major: njc(t,(j*ncat+p)*maxC+c) (:181)   */
  double *eta_star;    /* D x ntmax x T col-major cube: eta_star_curr(k,j,t) (:140)          */
  double *eta_empty;   /* D x CC x T col-major cube (:210)                                   */
  double *JJ;          /* nobs x D col-major (:70)                                           */
  double *ss;          /* [nobs] (:88)                                                       */
  double *loggamma;    /* nobs x D col-major (:161)                                          */
  double *beta;        /* [Q*D], index h = k + q*D (utils_ct.cpp:524)                        */
  double *sigma;       /* [T] (:50)                                                          */
  double *kappa;       /* [T] (:51)                                                          */
  int32_t *idxs;       /* [1] shared grid index (:45,:49) -- quirk: one scalar for all arms  */
  double *Theta;       /* D x T col-major (:271)                                             */
  double *Sigma;       /* D x D x T col-major cube (:272)                                    */
  double *lambda;      /* [Q*D] horseshoe local scales (:154)                                */
  double *tau;         /* [1] horseshoe global scale (:155)                                  */
  double *sigma_beta;  /* [Q*D] (:153)                                                       */
} tppmx_state;

typedef struct tppmx_ctx tppmx_ctx;
typedef struct tppmx_batch tppmx_batch;

/* ---- context ------------------------------------------------------------------------ */
int tppmx_version(void);
int tppmx_create(tppmx_ctx **ctx, int device);
int tppmx_destroy(tppmx_ctx *ctx);
/* Interruptibility of the long calls (the reference is interruptible only through R's own checks;
 * a blocking C call is not).  tppmx_batch_run and tppmx_dm_ppmx_ct drain the device every
 * `check_every` iterations (>= 1; 64 keeps the cost below 1 %) and read *flag: non-zero stops the
 * run with TPPMX_ERR_INTERRUPTED, the batch left in the state after the last completed iteration
 * (tppmx_dm_ppmx_ct returns without filling its outputs).  The flag is only ever read, from the
 * calling thread: an Rcpp shim passes &R_interrupts_pending, another host a flag its signal handler
 * or another thread sets.  flag == NULL switches the checks off (default). */
int tppmx_set_interrupt_flag(tppmx_ctx *ctx, const volatile int32_t *flag, int32_t check_every);
const char *tppmx_last_error(const tppmx_ctx *ctx);

/* Test hook: evaluates the library's device fp64 math on n inputs (which: 0 = log, 1 = exp,
 * 2 = lgamma for x > 0 in Stirling form, 3 = table-driven log, 4 = table-driven lgamma; the
 * branch-free versions the kernels use). */
int tppmx_probe_math(tppmx_ctx *ctx, int32_t which, const double *x, int32_t n, double *out);

/* ---- batch of chains ---------------------------------------------------------------- */
/* chain c runs on dataset chain_dataset[c] and uses Philox stream id chain_id[c]
 * (NULL => c).  All datasets share dims; copies everything it needs to the device. */
int tppmx_batch_create(tppmx_ctx *ctx, const tppmx_model *model, const tppmx_dims *dims,
                       const tppmx_shared *shared, int32_t n_datasets,
                       const tppmx_dataset *datasets, int32_t n_chains,
                       const int32_t *chain_dataset, const int32_t *chain_id,
                       tppmx_batch **out);
int tppmx_batch_destroy(tppmx_batch *b);

/* State initialisation exactly as the reference's prologue (:49-236): JJ=max(y,1e-100),
 * ss~Gamma, eta_star/eta_empty~U(0,1), beta=initbeta, Sigma=I, sufficient statistics
 * from the initial partition.  init_labels: n_datasets x (T x ntmax col-major) 1-based
 * labels (`curr_cluster`), init_nclu: n_datasets x T.  initbeta: [Q*D]. */
int tppmx_batch_init(tppmx_batch *b, uint64_t seed, const int32_t *init_labels,
                     const int32_t *init_nclu, const double *initbeta);

int tppmx_batch_set_state(tppmx_batch *b, int32_t chain, const tppmx_state *s);
int tppmx_batch_get_state(tppmx_batch *b, int32_t chain, tppmx_state *s);

/* Bulk download of every chain's partition (what the reference saves per iteration as
 * cl_lab / nclus, src/dm_ppmx_ct.cpp:1401-1404).  Either pointer may be NULL.
 *   labels [n_chains][T x ntmax col-major] 1-based, nclu [n_chains][T] */
int tppmx_batch_get_partitions(tppmx_batch *b, int32_t *labels, int32_t *nclu);

/* ---- hot path 1: reallocation sweep (:347-860) -------------------------------------- */
/* Parity hook.  Takes patient `it` (position inside arm `arm`) of chain `chain` out of its
 * cluster (:368-457) on a scratch copy of the arm and returns the K+CC log-weights
 * (:460-788, before the max-shift) and normalised probabilities (:791-802).  The chain
 * state is NOT modified.  K_out = occupied clusters after the removal. */
int tppmx_score_patient(tppmx_batch *b, int32_t chain, int32_t arm, int32_t it,
                        uint64_t seed, int32_t iter, double *logw_out, double *prob_out,
                        int32_t *K_out);

/* n_sweeps Gibbs sweeps over all arms and patients of every chain; everything that the
 * sweep does not update (eta_star of occupied clusters, beta, JJ, ss, sigma, idxs) stays
 * frozen.  Iteration numbers iter0..iter0+n_sweeps-1 key the RNG.  */
int tppmx_batch_sweep(tppmx_batch *b, uint64_t seed, int32_t iter0, int32_t n_sweeps);

/* Parity hook of the PRODUCTION kernel.  Runs ONE sweep (iteration `iter`) of every chain exactly
 * as tppmx_batch_sweep does, through an instantiation of the fast sweep kernel that also records,
 * for every step of chain `chain`, the normalised allocation probabilities the draw is made from
 * (reference pweight, src/dm_ppmx_ct.cpp:791-802) and the number of occupied clusters after the
 * removal.  The state advances like a normal sweep, so lock-step with the oracle is kept.
 *   prob_out [T][ntmax][stride] (row = position in the arm; entries 0..K+CC-1 are filled)
 *   K_out    [T][ntmax]         (-1: the step was not run by the fast kernel -- hand-over -- or
 *                                lies beyond the arm)
 * stride >= staged cluster capacity + CC (min(kcap, 32), or min(kcap, 256) for long arms; min(kcap, 256)
 * + CC always suffices).  Fails when the model is not eligible for the fast kernel. */
int tppmx_batch_sweep_probe(tppmx_batch *b, uint64_t seed, int32_t iter, int32_t chain,
                            int32_t stride, double *prob_out, int32_t *K_out);

/* Kernel selection.  The library has three sweep kernels with identical results:
 *   fast    — auxiliary similarity: cluster statistics staged in shared memory,
 *             likelihood terms hoisted into a parallel pre-pass, TPPMX_OPT_LANES_PER_UNIT (8|16|32)
 *             lanes per (chain, arm); 0 = automatic (16; 32 up to 10 units per SM and for long arms).
 *             Default switches (consim=1, similarity=1, calibration 0|2, ncat=0) for any shape, and
 *             -- for D <= 3 outcome categories, arms short enough for shared-memory label tables,
 *             16|32 lanes and at most lanes - CC staged clusters per arm -- one of: calibration 1;
 *             consim=2 (NNIG, P <= 32); categorical covariates (ncat <= 16); similarity=2 (double
 *             dipper, Normal-Normal); reuse=0; PPMx=0
 *   long-arm — arms of thousands of patients (default switches): the fast kernel in big-arm mode up
 *             to 32 clusters per arm, then one CTA per (chain, arm) with up to 256 staged clusters
 *   generic — every switch combination of the reference, one warp per (chain, arm)
 * TPPMX_OPT_SWEEP_IMPL: 0 = automatic (fast when eligible), 1 = generic, 2 = fast (error if
 * the model is not eligible).  A (chain, arm) that outgrows the staged cluster capacity is finished
 * from its exact position by the next kernel in the same call (TPPMX_INFO_LAST_HANDOVERS). */
enum {
  TPPMX_OPT_SWEEP_IMPL = 1,
  TPPMX_OPT_LANES_PER_UNIT = 2,
  TPPMX_OPT_FAST_STAGED_CLUSTERS = 3, /* clusters per arm the fast kernel stages on chip: 1..32, default
                                        min(kcap,32); long arms (thousands of patients) 1..256, default
                                        min(kcap,256); set before the first sweep */
  TPPMX_OPT_BIG_THREADS = 4,         /* long arms: threads per (chain, arm) of the long-arm kernel, 32 | 128 | 256;
                                        0 = automatic (32 when the units outnumber what the SMs hold) */
  TPPMX_OPT_CHAIN_GROUPS = 5         /* tppmx_batch_run: the chains are split into this many contiguous groups,
                                        each iterating on its own streams, so that the sweep of one group
                                        overlaps the updates of another (chains are independent; results do not
                                        depend on the grouping).  1..TPPMX_MAX_GROUPS, 0 = automatic (1: measured on
                                        B200, two groups of 512 chains run no faster than one of 1024) */
};
#define TPPMX_MAX_GROUPS 8
enum { TPPMX_INFO_LAST_SWEEP_IMPL = 1, TPPMX_INFO_LAST_HANDOVERS = 2, TPPMX_INFO_LAST_CHAIN_GROUPS = 3 };
int tppmx_batch_set_option(tppmx_batch *b, int32_t opt, int32_t value);
int tppmx_batch_get_info(const tppmx_batch *b, int32_t what, int32_t *value);

/* Device time (ms, CUDA events on the launching stream) and kernel launches of the last
 * tppmx_batch_sweep / tppmx_batch_predict / tppmx_batch_run call. */
int tppmx_batch_last_timing(const tppmx_batch *b, double *ms, int32_t *launches);

/* ---- hot path 2: posterior predictive (:1080-1385) ---------------------------------- */
/* For every chain, arm and new patient: score K+CC clusters on covariates only, draw the
 * cluster, take its eta (or N(Theta,Sigma) for an auxiliary pick), pi = softmax(eta+beta'z),
 * one multinomial draw.  Outputs (any may be NULL), chain-major:
 *   pipred    [n_chains][npred x D x T col-major]   (reference pii_pred,  :296)
 *   ypred     [n_chains][npred x D x T col-major]   (reference ppred,     :300)
 *   predclass [n_chains][npred x T col-major]       (reference predclass, :301) */
int tppmx_batch_predict(tppmx_batch *b, uint64_t seed, int32_t iter, double *pipred,
                        double *ypred, int32_t *predclass);

/* ---- the rest of an MCMC iteration and whole runs on the device ------------
 * tppmx_batch_update: everything of iteration `iter` that follows the sweep, for every chain:
 *   eta* Metropolis step per cluster (src/utils_ct.cpp:579-702), (Theta,Sigma) hierarchy
 *   (src/dm_ppmx_ct.cpp:880-902), (sigma,kappa) grid draw (:907-1007; needs ncat == 0 like the
 *   reference), beta Metropolis + horseshoe (src/utils_ct.cpp:710-800, 876-915),
 *   JJ / ss gamma draws (:1042-1055).
 * tppmx_batch_run: iterations iter0 .. iter0+n_iter-1 = sweep + update, and on the reference's
 *   saved iterations ((l > burn-1) && (l+1) % thin == 0, :1060) tppmx_batch_accumulate with the
 *   given flags.  Nothing returns to the host in between.
 * tppmx_batch_get_acceptance: acceptance counters of one chain in the reference's layouts
 *   (eta_flag T x ntmax, beta_flag Q x D, both col-major; either may be NULL). */
int tppmx_batch_update(tppmx_batch *b, uint64_t seed, int32_t iter);
int tppmx_batch_run(tppmx_batch *b, uint64_t seed, int32_t iter0, int32_t n_iter, int32_t burn,
                    int32_t thin, int32_t do_psm, int32_t do_pred);
int tppmx_batch_get_acceptance(tppmx_batch *b, int32_t chain, int32_t *eta_acc, int32_t *beta_acc);

/* ---- drop-in for the reference's entry point -------------------------------------------------
 * tppmx_dm_ppmx_ct runs what dm_ppmx_ct() runs (src/dm_ppmx_ct.cpp:10-1472) for ONE chain,
 * entirely on the device, from the same 42 arguments (in the reference's order and layouts:
 * doubles, col-major matrices, row-major flattened covariates, exactly what
 * _treatppmx_dm_ppmx_ct receives, src/RcppExports.cpp:16-65) into the same 18 outputs
 * (src/dm_ppmx_ct.cpp:1453-1471), caller-allocated and col-major so an Rcpp shim wraps them
 * without reshuffling.  nout = (iter - burn) / thin.  Any output pointer may be NULL.
 * Extra inputs: the leading dimensions Armadillo carries implicitly, `seed` (the shim draws it
 * from unif_rand() so set.seed() governs the run), and optionally the compact log V table
 * (logV_compact, layout of tppmx_shared.logV) instead of the dense Vwm cube. */
typedef struct tppmx_ppmxct_args {
  int32_t iter, burn, thin, nobs;
  const double *treatments;    /* [nobs], 0-based arm of each patient                    */
  int32_t PPMx, ncon, ncat;
  const double *catvec;        /* [ncat]                                                 */
  double alpha;
  const double *grid;          /* 2 x G col-major                                        */
  int32_t G;
  const double *Vwm;           /* Vwm_rows x Vwm_cols x G col-major cube (R/dm_ppmx_ct.R:243) or NULL */
  int32_t Vwm_rows, Vwm_cols;
  const double *logV_compact;  /* [A][2][ntmax][G] or NULL (used when Vwm == NULL)       */
  int32_t cohesion, CC, reuse, consim, similarity, calibration, coardegree;
  const double *y;             /* nobs x D col-major                                     */
  int32_t D;
  const double *z;             /* nobs x Q col-major                                     */
  int32_t Q;
  const double *zpred;         /* npred x Q col-major                                    */
  int32_t noprog;
  const double *xcon, *xcat, *xconp, *xcatp; /* row-major flattened (as.vector(t(.)))     */
  int32_t npred;
  const double *similparam;    /* [7]                                                    */
  const double *hP0_mu0;       /* [D]                                                    */
  double hP0_nu0, hP0_s0, hP0_Lambda0;
  int32_t upd_hier;
  const double *initbeta;      /* [Q*D]                                                  */
  int32_t hsp;
  const double *mhtunepar;     /* [2]                                                    */
  int32_t A;
  const double *n_a;           /* [A]                                                    */
  const double *curr_cluster;  /* A x ntmax col-major, 1-based labels                    */
  const double *card_cluster;  /* A x ntmax col-major (recomputed from curr_cluster)     */
  int32_t ntmax;               /* = ncol(curr_cluster) = max n_a                         */
  const double *ncluster_curr; /* [A]                                                    */
  uint64_t seed;
} tppmx_ppmxct_args;

typedef struct tppmx_ppmxct_out { /* names of the returned list, :1453-1471 */
  double *eta;         /* field nout x A of D x ntmax matrices: block (ll + nout*t)        */
  double *beta;        /* (Q*D) x nout                                                     */
  double *sigma_ngg;   /* A x nout                                                         */
  double *kappa_ngg;   /* A x nout                                                         */
  double *eta_acc;     /* A x ntmax : eta_flag / sumtotclu (:1432-1434)                    */
  double *beta_acc;    /* Q x D     : beta_flag (raw counts, as the reference returns)     */
  double *cl_lab;      /* A x ntmax x nout                                                 */
  double *num_treat;   /* [A]                                                              */
  double *pi;          /* nobs x D x nout                                                  */
  double *pipred;      /* nout blocks of npred x D x A                                     */
  double *yispred;     /* nobs x D x nout                                                  */
  double *ypred;       /* nout blocks of npred x D x A                                     */
  double *clupred;     /* (nout*npred) x A                                                 */
  double *nclus;       /* A x nout                                                         */
  double *WAIC;        /* [1]                                                              */
  double *lpml;        /* [nout]                                                           */
  double *theta_prior; /* D x A                                                            */
  double *sigma_prior; /* D x D x A                                                        */
} tppmx_ppmxct_out;

int tppmx_dm_ppmx_ct(tppmx_ctx *ctx, const tppmx_ppmxct_args *in, tppmx_ppmxct_out *out);

/* Point-estimate partition per replicate dataset and arm from the co-clustering matrix of the last
 * tppmx_batch_summaries[_device] call: what users compute next with mcclust.ext::minVI /
 * minbinder.ext (imported by the package, R/treatppmx.R:30-31; un-vendored, so the published
 * algorithm is restated -- Wade & Ghahramani 2018, Bayesian Analysis 13(2) section 4, method "avg" / "comp"):
 * hierarchical clustering of 1 - psm (linkage 0 = average, 1 = complete), every cut with
 * k = 1..max_k clusters (max_k = 0 => ceiling(n_t / 8), the package's default) scored with the loss
 * (0 = Binder sum_{i<j} |1{c_i = c_j} - psm_ij|, 1 = lower bound of the variation of information),
 * the best cut returned.  labels_out [n_datasets][T x ntmax col-major] 1-based in order of first
 * appearance (0 beyond the arm), k_out / loss_out [n_datasets][T] (may be NULL).  n_t <= 144. */
int tppmx_batch_point_estimate(tppmx_batch *b, int32_t loss, int32_t linkage, int32_t max_k,
                               int32_t *labels_out, int32_t *k_out, double *loss_out);

/* ---- drop-in for the package's second sampler ------------------------------------------------
 * tppmx_prior_ppmx_core runs what prior_ppmx_core() runs (reference src/prior_ppmx.cpp:9-466;
 * .Call entry _treatppmx_prior_ppmx_core, src/RcppExports.cpp:189; R wrapper R/prior_ppmx.R:32-132):
 * the reallocation sweep with ONE arm, continuous covariates only and NO likelihood term, from the
 * same 21 arguments in the reference's order and layouts, into the two elements of its returned
 * list.  Entirely on the device: one sweep launch per iteration and one save kernel per saved
 * iteration ((l > burn-1) && (l+1) % thin == 0, :455), nothing returns to the host in between.
 * Extra inputs: the leading dimensions of Vwm, `seed`, and n_chains (>= 1 independent replicas of
 * the run, Philox stream ids 0..n_chains-1; the reference runs one). */
typedef struct tppmx_prior_args {
  int32_t iter, burn, thin, nobs, PPMx, ncon, ncat;
  double alpha, sigma;
  const double *Vwm;           /* Vwm_rows x Vwm_cols col-major, V_{n,k} itself (not its log); rows    */
  int32_t Vwm_rows, Vwm_cols;  /* nobs-2 and nobs-1 are read (:239); may be NULL when cohesion == 1     */
  int32_t cohesion, CC, consim, similarity, calibration, coardegree;
  const double *xcon;          /* nobs x ncon row-major (as.vector(t(xcon)), R/prior_ppmx.R:121)        */
  const double *similparam;    /* [7]                                                                  */
  const double *curr_cluster;  /* [nobs] 1-based labels                                                */
  const double *card_cluster;  /* [nobs] (recomputed from curr_cluster)                                */
  int32_t ncluster_curr;
  uint64_t seed;
  int32_t n_chains;
} tppmx_prior_args;

typedef struct tppmx_prior_out { /* names of the returned list, src/prior_ppmx.cpp:463-465 */
  double *nclu;   /* [n_chains][nout]                        number of clusters per saved iteration */
  double *njout;  /* [n_chains][nobs x nout col-major]       cluster sizes per saved iteration      */
} tppmx_prior_out;

int tppmx_prior_ppmx_core(tppmx_ctx *ctx, const tppmx_prior_args *in, tppmx_prior_out *out);

/* ---- posterior summaries on the device (what a multi-GPU run gathers) ----------------------
 * The package returns per-iteration cl_lab / pipred (src/dm_ppmx_ct.cpp:1402-1404, :1396) and
 * leaves the co-clustering matrix (mcclust::comp.psm, R/treatppmx.R:29) and the treatment
 * utility sum_k w_k E[pi_k] to user scripts.  For batched chains these are accumulated on the
 * device per replicate dataset, over all of its chains and all saved iterations:
 *   tppmx_batch_accumulate   adds the current partition of every chain to its dataset's
 *                            co-clustering counts (do_psm) and runs the predictive step
 *                            (RNG keyed by `iter`) adding pipred to the dataset's sums (do_pred)
 *   tppmx_batch_summaries    means / utilities / chosen arm, copied to host (any pointer NULL):
 *        psm      [n_datasets][T][ntmax][ntmax]   P(label_a == label_b)
 *        pi_mean  [n_datasets][npred x D x T col-major]
 *        utility  [n_datasets][T][npred]           sum_k util_w[k] * pi_mean
 *        best_arm [n_datasets][npred]              arg max_t utility (0-based)
 *   tppmx_batch_summaries_device  the same four arrays left on the device (pointers returned),
 *        for a collective (NCCL) gather without a host round trip
 *   tppmx_batch_reset_summaries   zero the accumulators */
int tppmx_batch_accumulate(tppmx_batch *b, uint64_t seed, int32_t iter, int32_t do_psm, int32_t do_pred);
int tppmx_batch_summaries(tppmx_batch *b, const double *util_w, double *psm, double *pi_mean,
                          double *utility, int32_t *best_arm);
int tppmx_batch_summaries_device(tppmx_batch *b, const double *util_w, double **psm, double **pi_mean,
                                 double **utility, int32_t **best_arm);
int tppmx_batch_reset_summaries(tppmx_batch *b);
/* npc of R/countUT.R:26-53 from the predictive means of the last tppmx_batch_summaries[_device]
 * call: per replicate dataset, how many new patients have arg max_k E[pipred(pp, k, trtsgn_pp)]
 * equal to the observed outcome.  trtsgn (assigned arm) and outcome (observed category) are
 * [n_datasets][npred], 0-based; npc_out is [n_datasets]. */
int tppmx_batch_npc(tppmx_batch *b, const int32_t *trtsgn, const int32_t *outcome, int32_t *npc_out);

/* ---- one process, several GPUs (tppmx_create(ctx**, device_ids[], ndev)) ----------
 * The reference has no parallelism at all; its workload (a simulation study over replicate data
 * sets, R/gendata.R:198-216) is embarrassingly parallel.  tppmx_multi_run_study shards the replicate
 * datasets, each with all of its chains, over the devices in contiguous blocks (one host thread and
 * one context per device, no data-path collective), runs n_iter full iterations per chain with the
 * summaries accumulated on the saved iterations, and gathers the per-dataset posterior summaries with
 * NCCL over NVLink (ncclCommInitAll; grouped ncclSend / ncclRecv into device 0) before one copy to
 * the caller.  Chain c of dataset d uses Philox stream d * chains_per_dataset + c, so the results do
 * not depend on the number of devices.  NCCL is loaded at run time and only when ndev > 1. */
typedef struct tppmx_multi tppmx_multi;
typedef struct tppmx_study {
  const tppmx_model *model;
  const tppmx_dims *dims;
  const tppmx_shared *shared;
  int32_t n_datasets;
  const tppmx_dataset *datasets;   /* [n_datasets]                                              */
  int32_t chains_per_dataset;
  const int32_t *init_labels;      /* [n_datasets][T x ntmax col-major], as tppmx_batch_init      */
  const int32_t *init_nclu;        /* [n_datasets][T]                                            */
  const double *initbeta;          /* [Q*D]                                                      */
  uint64_t seed;
  int32_t n_iter, burn, thin;      /* saved iterations: (l > burn-1) && (l+1) % thin == 0         */
  int32_t do_psm, do_pred;
  const double *util_w;            /* [D] utility weights (NULL = zeros)                         */
} tppmx_study;
typedef struct tppmx_study_out {   /* host buffers (any may be NULL), layouts of tppmx_batch_summaries */
  double *psm;                     /* [n_datasets][T][ntmax][ntmax]                              */
  double *pi_mean;                 /* [n_datasets][npred x D x T col-major]                      */
  double *utility;                 /* [n_datasets][T][npred]                                     */
  int32_t *best_arm;               /* [n_datasets][npred]                                        */
  double ms_run;                   /* device time of the chains, max over the devices            */
  double ms_gather;                /* device time of the NCCL gather (0 on one device)           */
  int32_t n_devices, used_nccl;
} tppmx_study_out;
int tppmx_multi_create(tppmx_multi **m, const int32_t *device_ids, int32_t ndev); /* device_ids NULL = 0..ndev-1 */
int tppmx_multi_destroy(tppmx_multi *m);
int tppmx_multi_device_count(const tppmx_multi *m);
const char *tppmx_multi_last_error(const tppmx_multi *m);
int tppmx_multi_run_study(tppmx_multi *m, const tppmx_study *in, tppmx_study_out *out);

#ifdef __cplusplus
}
#endif
#endif /* TPPMX_H */
