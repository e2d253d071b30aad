/*
 * ppmx_oracle.h — CPU oracle for the treatppmx hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the reference's algorithm, used as the parity checker by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.  Nothing in
 * treatppmx_b200/ (the product) may include, link or call this.
 *
 * Parity pin status: PINNED against the reference's own sources executed here.  The reference
 * ships no tests, golden vectors or fixtures for this path and R / Rcpp /
 * Armadillo are absent, so src/utils_ct.cpp, src/dm_ppmx_ct.cpp and src/prior_ppmx.cpp are
 * compiled UNMODIFIED against the stand-in headers of oracle/refshim/ into oracle/_ref/libref.so
 * (make -C oracle ref) and compared with this oracle (tests/test_oracle_vs_reference.py):
 * helper by helper on random inputs (<= 1e-13), and as whole runs of dm_ppmx_ct /
 * prior_ppmx_core driven by the tape of variates this oracle consumes (oracle_tape_*): all 18
 * outputs, every switch combination the reference can run, bit-identical labels / counts /
 * draws, <= 1e-12 on reals.  The reference's outputs are committed as tests/golden/ref_*.npz.
 * What stays a stand-in: the library layer under the sources (R's RNG stream, Armadillo's
 * mvnrnd / wishrnd construction, libRmath pgamma / qgamma) -- see oracle/refshim/RcppArmadillo.h.
 * Closed-form known-answer tests remain as an independent check (tests/test_oracle_kat.py).
 *
 * State and inputs use the structs of include/tppmx.h (reference layouts).
 */
#ifndef PPMX_ORACLE_H
#define PPMX_ORACLE_H
#include "../include/tppmx.h"

#ifdef __cplusplus
extern "C" {
#endif

/* bundle of the read-only inputs of one chain */
typedef struct oracle_problem {
  tppmx_model model;
  tppmx_dims dims;
  tppmx_shared shared;
  tppmx_dataset data;
} oracle_problem;

/* helpers, exported for the known-answer tests (reference src/utils_ct.cpp) */
double oracle_dnorm_log(double x, double mu, double sigma);                       /* R::dnorm(.,1) */
double oracle_dinvgamma(double y, double alpha, double beta, int logout);         /* :161-176 */
double oracle_dN_IG(double mu, double sig2, double mu0, double k0, double a0, double b0,
                    int logout);                                                  /* :181-200 */
double oracle_gsimconNN(double m0, double v2, double s20, double sumx, double sumx2, int n,
                        int DD, int logout);                                      /* :382-405 */
double oracle_gsimconNNIG(double m0, double k0, double nu0, double s20, double sumx,
                          double sumx2, int n, int DD, int logout);               /* :427-459 */
double oracle_gsimcatDM(const double *nobsj, const double *dirweights, int C, int DD,
                        int logout);                                              /* :468-495 */

/* RNG contract probes (for the device/host RNG agreement test) */
void oracle_rng_u2(uint64_t seed, uint32_t chain, uint32_t iter, uint32_t site, uint32_t arm,
                   uint32_t index, uint32_t sub, double *out2);
void oracle_rng_n2(uint64_t seed, uint32_t chain, uint32_t iter, uint32_t site, uint32_t arm,
                   uint32_t index, uint32_t sub, double *out2);
double oracle_rng_gamma(uint64_t seed, uint32_t chain, uint32_t iter, uint32_t site,
                        uint32_t arm, uint32_t index, double shape);

/* variate tape: record every variate the oracle consumes, in call order, for replay into the
 * reference's own sources (oracle/refshim, tests/test_oracle_vs_reference.py) */
void oracle_tape_start(void);
long oracle_tape_stop(void);
void oracle_tape_get(double *kind, double *val, double *aux, long n);

/* prologue of dm_ppmx_ct (src/dm_ppmx_ct.cpp:49-236) */
int oracle_init_state(const oracle_problem *pb, tppmx_state *s, uint64_t seed, uint32_t chain_id,
                      const int32_t *init_labels, const int32_t *init_nclu,
                      const double *initbeta);

/* remove patient (arm,it) on a private copy, return log-weights / probabilities */
int oracle_score_patient(const oracle_problem *pb, const tppmx_state *s, int32_t arm,
                         int32_t it, uint64_t seed, uint32_t chain_id, int32_t iter,
                         double *logw_out, double *prob_out, int32_t *K_out);

/* n_sweeps passes of src/dm_ppmx_ct.cpp:349-860 with everything else frozen */
int oracle_sweep(const oracle_problem *pb, tppmx_state *s, uint64_t seed, uint32_t chain_id,
                 int32_t iter0, int32_t n_sweeps);

/* one sweep (iteration l) recording each step's pweight (:791-802) and K after removal:
 * prob_out [T][ntmax][stride], K_out [T][ntmax] */
int oracle_sweep_probe(const oracle_problem *pb, tppmx_state *s, uint64_t seed, uint32_t chain_id,
                       int32_t l, int32_t stride, double *prob_out, int32_t *K_out);

/* src/dm_ppmx_ct.cpp:1081-1384 */
int oracle_predict(const oracle_problem *pb, const tppmx_state *s, uint64_t seed,
                   uint32_t chain_id, int32_t iter, double *pipred, double *ypred,
                   int32_t *predclass);

/* ppmx_chain.c: everything of iteration l after the sweep (src/dm_ppmx_ct.cpp:866-1055), and
 * n_iter full iterations (sweep + updates).  eta_acc: T x ntmax col-major, beta_acc: Q x D
 * col-major acceptance counters (may be NULL). */
int oracle_update_iter(const oracle_problem *pb, tppmx_state *s, uint64_t seed, uint32_t chain_id,
                       int32_t l, int32_t *eta_acc, int32_t *beta_acc);
int oracle_iterate(const oracle_problem *pb, tppmx_state *s, uint64_t seed, uint32_t chain_id,
                   int32_t iter0, int32_t n_iter, int32_t *eta_acc, int32_t *beta_acc);
int oracle_insample(const oracle_problem *pb, const tppmx_state *s, uint64_t seed, uint32_t chain_id,
                    int32_t l, double *pi, double *ispred, double *like);
double oracle_gammainc(double a, double x);
double oracle_gammaincinv(double a, double p);
double oracle_dmvnrm_log(const double *x, const double *mean, const double *sigma, int n);
void oracle_wishart(uint64_t seed, uint32_t chain, int32_t l, const double *S, double df, int n, double *W);

#ifdef __cplusplus
}
#endif
#endif
