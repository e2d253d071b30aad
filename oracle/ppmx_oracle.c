/*
 * ppmx_oracle.c — CPU oracle of the treatppmx hot path.  TEST INFRASTRUCTURE ONLY
 * (see ppmx_oracle.h for the rules and the parity-pin status: pinned against the reference's own
 * sources compiled into oracle/_ref and run on this oracle's variate tape, plus closed-form
 * known-answer tests).
 *
 * Restates, statement by statement and with the same loop structure, evaluation order and
 * quirks:
 *   reallocation sweep        /root/reference/src/dm_ppmx_ct.cpp:347-860
 *   posterior predictive      /root/reference/src/dm_ppmx_ct.cpp:1080-1385
 *   state prologue            /root/reference/src/dm_ppmx_ct.cpp:49-236
 *   helpers                   /root/reference/src/utils_ct.cpp:161-200, 382-535
 * Armadillo containers become flat arrays in the same (column-major) layouts; R's
 * dnorm(log=TRUE) is written out with R's own operation order; lgamma / R::lgammafn are
 * libm lgamma; the RNG is the Philox contract of include/tppmx.h (philox_ref.h).
 */
#include "ppmx_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "philox_ref.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
#define O_LN_SQRT_2PI 0.918938533204672741780329736406 /* R's M_LN_SQRT_2PI */

/* ---------------------------------------------------------------------------------------
 * helpers (reference src/utils_ct.cpp)
 * ------------------------------------------------------------------------------------- */

/* R::dnorm(x, mu, sigma, give_log = 1): R's nmath/dnorm.c computes x=(x-mu)/sigma and
 * returns -(M_LN_SQRT_2PI + 0.5*x*x + log(sigma)). */
double oracle_dnorm_log(double x, double mu, double sigma) {
  double t = (x - mu) / sigma;
  return -(O_LN_SQRT_2PI + 0.5 * t * t + log(sigma));
}

/* utils_ct.cpp:161-176 */
double oracle_dinvgamma(double y, double alpha, double beta, int logout) {
  double ldens = alpha * log(beta) - lgamma(alpha) - (alpha + 1) * log(y) - (beta / y);
  return logout ? ldens : exp(ldens);
}

/* utils_ct.cpp:181-200 */
double oracle_dN_IG(double mu, double sig2, double mu0, double k0, double a0, double b0,
                    int logout) {
  /* note: the reference passes `logout` as give_log to both pieces and then exponentiates
   * again if !logout; only logout==1 is ever used on the path. */
  double ldens = oracle_dnorm_log(mu, mu0, sqrt(sig2 / k0)) + oracle_dinvgamma(sig2, a0, b0, 1);
  return logout ? ldens : exp(ldens);
}

/* utils_ct.cpp:382-405 */
double oracle_gsimconNN(double m0, double v2, double s20, double sumx, double sumx2, int n,
                        int DD, int logout) {
  double mus, muss, s2s, s2ss;
  double ld1, ld2, ld3, ld4, out;

  s2s = 1 / ((n / v2) + (1 / s20));
  mus = s2s * ((1 / v2) * sumx + (1 / s20) * m0);

  s2ss = 1 / ((n / v2) + (1 / s2s));
  muss = s2ss * ((1 / v2) * sumx + (1 / s2s) * mus);

  ld1 = -0.5 * n * log(2 * M_PI * v2) - 0.5 * (1 / v2) * sumx2;
  ld2 = oracle_dnorm_log(m0, 0, sqrt(s20));
  ld3 = oracle_dnorm_log(mus, 0, sqrt(s2s));
  ld4 = oracle_dnorm_log(muss, 0, sqrt(s2ss));

  out = ld1 + ld2 - ld3;
  if (DD == 1) out = ld1 + ld3 - ld4;
  if (!logout) out = exp(out);
  return out;
}

/* utils_ct.cpp:427-459 */
double oracle_gsimconNNIG(double m0, double k0, double nu0, double s20, double sumx,
                          double sumx2, int n, int DD, int logout) {
  double a0, b0, m0s, m0ss, k0s, k0ss, a0s, a0ss, b0s, b0ss;
  double ld1, ld2, ld3, ld4, out;
  double mu = 10, v2 = 0.1;
  double xbar = sumx * (1 / (double)n);

  a0 = 0.5 * nu0;
  b0 = 0.5 * nu0 * s20;

  m0s = (k0 * m0 + n * xbar) / (k0 + n);
  k0s = k0 + (double)n;
  a0s = a0 + 0.5 * n;
  b0s = b0 + 0.5 * (sumx2 - n * xbar * xbar) + 0.5 * n * k0 * (xbar - m0) * (xbar - m0) / (k0 + n);

  m0ss = (k0s * m0s + n * xbar) / (k0s + n);
  k0ss = k0s + (double)n;
  a0ss = a0s + 0.5 * n;
  b0ss = b0s + 0.5 * (sumx2 - n * xbar * xbar) +
         0.5 * n * k0s * (xbar - m0s) * (xbar - m0s) / (k0s + n);

  ld1 = -0.5 * n * log(2 * M_PI * v2) - 0.5 * (1 / v2) * (sumx2 - 2 * sumx * mu + n * mu * mu);
  ld2 = oracle_dN_IG(mu, v2, m0, k0, a0, b0, 1);
  ld3 = oracle_dN_IG(mu, v2, m0s, k0s, a0s, b0s, 1);
  ld4 = oracle_dN_IG(mu, v2, m0ss, k0ss, a0ss, b0ss, 1);

  out = ld1 + ld2 - ld3;
  if (DD == 1) out = ld1 + ld3 - ld4;
  if (!logout) out = exp(out);
  return out;
}

/* utils_ct.cpp:468-495 */
double oracle_gsimcatDM(const double *nobsj, const double *dirweights, int C, int DD, int logout) {
  int ii, sumc;
  double tmp1 = 0.0, tmp2 = 0.0, tmp3 = 0.0, tmp4 = 0.0, tmp5 = 0.0, tmp6 = 0.0;
  double out;

  sumc = 0;
  for (ii = 0; ii < C; ii++) {
    sumc = sumc + (int)nobsj[ii];
    tmp1 = tmp1 + dirweights[ii];
    tmp2 = tmp2 + lgamma(dirweights[ii]);
    tmp3 = tmp3 + (double)nobsj[ii] + dirweights[ii];
    tmp4 = tmp4 + lgamma((double)nobsj[ii] + dirweights[ii]);
    tmp5 = tmp5 + 2 * ((double)nobsj[ii]) + dirweights[ii];
    tmp6 = tmp6 + lgamma(2 * ((double)nobsj[ii]) + dirweights[ii]);
  }
  out = (lgamma(tmp1) - tmp2) + (tmp4 - lgamma(tmp3));
  if (DD == 1) out = (lgamma(tmp3) - tmp4) + (tmp6 - lgamma(tmp5));
  if (sumc == 0) out = log(1);
  if (!logout) out = exp(out);
  return out;
}

/* RNG probes */
/* ---- variate tape (philox_ref.h) ------------------------------------------------------------- */
int o_tape_on = 0, o_tape_depth = 0;
static double *o_tape_buf = NULL; /* [cap][3] = kind, value, aux */
static long o_tape_n = 0, o_tape_cap = 0;
void o_tape_rec(int kind, double val, double aux) {
  if (o_tape_n == o_tape_cap) {
    o_tape_cap = o_tape_cap ? 2 * o_tape_cap : (1 << 16);
    o_tape_buf = (double *)realloc(o_tape_buf, sizeof(double) * 3 * (size_t)o_tape_cap);
  }
  o_tape_buf[3 * o_tape_n + 0] = (double)kind;
  o_tape_buf[3 * o_tape_n + 1] = val;
  o_tape_buf[3 * o_tape_n + 2] = aux;
  o_tape_n++;
}
void oracle_tape_start(void) {
  o_tape_n = 0;
  o_tape_depth = 0;
  o_tape_on = 1;
}
long oracle_tape_stop(void) {
  o_tape_on = 0;
  return o_tape_n;
}
/* copies records [0, n) as three parallel arrays */
void oracle_tape_get(double *kind, double *val, double *aux, long n) {
  for (long i = 0; i < n && i < o_tape_n; i++) {
    kind[i] = o_tape_buf[3 * i];
    val[i] = o_tape_buf[3 * i + 1];
    aux[i] = o_tape_buf[3 * i + 2];
  }
}

void oracle_rng_u2(uint64_t seed, uint32_t chain, uint32_t iter, uint32_t site, uint32_t arm,
                   uint32_t index, uint32_t sub, double *out2) {
  o_rng g = {seed, chain};
  o_rng_u2(&g, iter, site, arm, index, sub, &out2[0], &out2[1]);
}
void oracle_rng_n2(uint64_t seed, uint32_t chain, uint32_t iter, uint32_t site, uint32_t arm,
                   uint32_t index, uint32_t sub, double *out2) {
  o_rng g = {seed, chain};
  o_rng_n2(&g, iter, site, arm, index, sub, &out2[0], &out2[1]);
}
double oracle_rng_gamma(uint64_t seed, uint32_t chain, uint32_t iter, uint32_t site,
                        uint32_t arm, uint32_t index, double shape) {
  o_rng g = {seed, chain};
  return o_rng_gamma(&g, iter, site, arm, index, 0, shape);
}

/* ---------------------------------------------------------------------------------------
 * layout accessors (Armadillo column-major, as declared in include/tppmx.h)
 * ------------------------------------------------------------------------------------- */
#include "oracle_internal.h"

int o_ctx_open(o_ctx *c, const oracle_problem *pb, tppmx_state *s, uint64_t seed,
                      uint32_t chain_id) {
  memset(c, 0, sizeof(*c));
  c->pb = pb;
  c->s = s;
  const tppmx_dims *d = &pb->dims;
  c->nobs = d->nobs; c->T = d->T; c->D = d->D; c->Q = d->Q; c->P = d->P;
  c->ncat = d->ncat; c->maxC = d->maxC; c->npred = d->npred; c->G = d->G;
  c->ntmax = d->ntmax; c->CC = pb->model.CC;
  if (c->T > TPPMX_MAXT || c->D > TPPMX_MAXD || c->Q > TPPMX_MAXQ || c->CC > TPPMX_MAXCC) return 1;
  for (int i = 0; i < c->nobs; i++) c->num_treat[pb->data.treat[i]]++;
  int mc = c->maxC > 0 ? c->maxC : 1;
  c->dirweights = (double *)malloc(sizeof(double) * mc);
  c->njctmp = (double *)calloc(mc, sizeof(double));
  for (int k = 0; k < mc; k++) c->dirweights[k] = pb->model.similparam[6]; /* :220-221 */
  int W = c->ntmax + c->CC;
  c->weight = (double *)calloc(6 * (size_t)W, sizeof(double));
  c->pweight = c->weight + W; c->gtilN = c->pweight + W; c->gtilY = c->gtilN + W;
  c->lgtilN = c->gtilY + W; c->lgtilY = c->lgtilN + W;
  c->rng.seed = seed; c->rng.chain = chain_id;
  return 0;
}
void o_ctx_close(o_ctx *c) {
  free(c->dirweights); free(c->njctmp); free(c->weight);
}

/* utils_ct.cpp:497-535 — eta given as a column pointer (the n_cols==1 branch and the
 * (k, clu_lg) branch read the same element). */
static double o_calculate_gamma(const o_ctx *c, const double *eta_col, const double *zz,
                                int zrows, int k, int i, int Log) {
  double lg = eta_col[k];
  for (int q = 0; q < c->Q; q++) {
    int h = k + q * c->D;
    lg += (c->s->beta[h] * zz[i + zrows * q]);
  }
  return Log == 1 ? lg : exp(lg);
}

/* one continuous-covariate similarity term, switch structure of :472-516 */
double o_gcon(const o_ctx *c, double sx, double sx2, int n) {
  const tppmx_model *m = &c->pb->model;
  double m0 = m->similparam[0], s20 = m->similparam[1], v = m->similparam[2];
  double k0 = m->similparam[3], nu0 = m->similparam[4];
  double out = 0.0;
  if (m->similarity == 1) {
    if (m->consim == 1) out = oracle_gsimconNN(m0, v, s20, sx, sx2, n, 0, 1);
    if (m->consim == 2) out = oracle_gsimconNNIG(m0, k0, nu0, s20, sx, sx2, n, 0, 1);
  }
  if (m->similarity == 2) {
    if (m->consim == 1) out = oracle_gsimconNN(m0, v, s20, sx, sx2, n, 1, 1);
    if (m->consim == 2) out = oracle_gsimconNNIG(m0, k0, nu0, s20, sx, sx2, n, 1, 1);
  }
  return out;
}
static double o_gcat(const o_ctx *c, int p) {
  const tppmx_model *m = &c->pb->model;
  double out = 0.0;
  if (m->similarity == 1) out = oracle_gsimcatDM(c->njctmp, c->dirweights, c->pb->shared.catvec[p], 0, 1);
  if (m->similarity == 2) out = oracle_gsimcatDM(c->njctmp, c->dirweights, c->pb->shared.catvec[p], 1, 1);
  return out;
}

/* log Vwm(n_t-1, K-1, idxs) - log Vwm(n_t-2, K-1, idxs)   (:563, :680) from the compact table */
static double o_logV(const o_ctx *c, int t, int r, int K, int g) {
  return c->pb->shared.logV[(((size_t)t * 2 + r) * c->ntmax + (K - 1)) * c->G + g];
}

/* log V(n_t, K) - log V(n_t - 1, K) at grid index g (:563, :680).  K = 0 can only happen while
 * the single patient of a one-patient arm is out of its cluster; the reference would index
 * Vwm(., -1, .) there (an Armadillo bounds error), so the case is undefined upstream.  The term is
 * common to all candidate clusters and cancels in the normalisation: it is taken as 0, as on the
 * device. */
static double o_dV(const o_ctx *c, int t, int K, int g) {
  if (K < 1) return 0.0;
  return o_logV(c, t, 1, K, g) - o_logV(c, t, 0, K, g);
}

/* the DM-augmented likelihood piece, :570-576 */
static double o_lik(const o_ctx *c, double w, const double *eta_col, int i) {
  if (c->pb->model.nolik) return w; /* prior_ppmx_core: no likelihood term */
  for (int k = 0; k < c->D; k++) {
    double wo = o_calculate_gamma(c, eta_col, c->pb->data.z, c->nobs, k, i, 0);
    w += wo * log(JJm(i, k)) - lgamma(wo) - JJm(i, k) * (c->s->ss[i] + 1.0);
    if (Ym(i, k) != 1) w -= log(JJm(i, k));
  }
  return w;
}

/* ---------------------------------------------------------------------------------------
 * one patient of the sweep: src/dm_ppmx_ct.cpp:364-858
 *   score_only: stop after the weights (used by oracle_score_patient)
 * ------------------------------------------------------------------------------------- */
static void o_realloc_patient(o_ctx *c, int tt, int it, int i, int l, int score_only,
                              double *logw_out, double *prob_out, int *K_out) {
  const tppmx_model *m = &c->pb->model;
  const int ncon = c->P, ncat = c->ncat, max_C = c->maxC, dim = c->D, CC = c->CC;
  const int PPMx = m->PPMx, cohesion = m->cohesion, calibration = m->calibration;
  const int coardegree = m->coardegree;
  double *weight = c->weight, *pweight = c->pweight;
  double *gtilN = c->gtilN, *gtilY = c->gtilY, *lgtilN = c->lgtilN, *lgtilY = c->lgtilY;
  int *nclu_curr = c->s->nclu;
  const int idxs = c->s->idxs[0];
  int zi, iaux, ii, p, cidx, j, jj, mm, k;
  double auxreal, sumxtmp, sumx2tmp, lgconN, lgconY, lgcatN, lgcatY, lgcondraw, lgcatdraw, tmp;
  double auxv[TPPMX_MAXD];

  zi = LAB(tt, it) - 1; /* :368 */

  if (NJ(tt, zi) > 1) { /* :370-383 */
    NJ(tt, zi) -= 1;
    if (PPMx == 1) {
      for (p = 0; p < ncon; p++) {
        SUMX(tt, zi * ncon + p) -= XCON(i * ncon + p);
        SUMX2(tt, zi * ncon + p) -= XCON(i * ncon + p) * XCON(i * ncon + p);
      }
      for (p = 0; p < ncat; p++) NJC(tt, (zi * ncat + p) * max_C + XCAT(i * ncat + p)) -= 1;
    }
  } else { /* singleton, :384-457 */
    iaux = LAB(tt, it);
    if (iaux < nclu_curr[tt]) {
      for (ii = 0; ii < c->num_treat[tt]; ii++) {
        if (LAB(tt, ii) == nclu_curr[tt]) LAB(tt, ii) = iaux;
      }
      LAB(tt, it) = nclu_curr[tt];

      for (k = 0; k < dim; k++) auxv[k] = ETAS(k, iaux - 1, tt);
      for (k = 0; k < dim; k++) ETAS(k, iaux - 1, tt) = ETAS(k, nclu_curr[tt] - 1, tt);
      for (k = 0; k < dim; k++) ETAS(k, nclu_curr[tt] - 1, tt) = auxv[k];

      NJ(tt, iaux - 1) = NJ(tt, nclu_curr[tt] - 1);
      NJ(tt, nclu_curr[tt] - 1) = 1;

      if (PPMx == 1) {
        for (p = 0; p < ncon; p++) {
          auxreal = SUMX(tt, (iaux - 1) * ncon + p);
          SUMX(tt, (iaux - 1) * ncon + p) = SUMX(tt, (nclu_curr[tt] - 1) * ncon + p);
          SUMX(tt, (nclu_curr[tt] - 1) * ncon + p) = auxreal;

          auxreal = SUMX2(tt, (iaux - 1) * ncon + p);
          SUMX2(tt, (iaux - 1) * ncon + p) = SUMX2(tt, (nclu_curr[tt] - 1) * ncon + p);
          SUMX2(tt, (nclu_curr[tt] - 1) * ncon + p) = auxreal;
        }
        for (p = 0; p < ncat; p++) {
          for (cidx = 0; cidx < max_C; cidx++) {
            int auxint = NJC(tt, ((iaux - 1) * ncat + p) * max_C + cidx);
            NJC(tt, ((iaux - 1) * ncat + p) * max_C + cidx) =
                NJC(tt, ((nclu_curr[tt] - 1) * ncat + p) * max_C + cidx);
            NJC(tt, ((nclu_curr[tt] - 1) * ncat + p) * max_C + cidx) = auxint;
          }
        }
      }
    }

    NJ(tt, nclu_curr[tt] - 1) -= 1;

    if (PPMx == 1) {
      for (p = 0; p < ncon; p++) {
        SUMX(tt, (nclu_curr[tt] - 1) * ncon + p) -= XCON(i * ncon + p);
        SUMX2(tt, (nclu_curr[tt] - 1) * ncon + p) -= XCON(i * ncon + p) * XCON(i * ncon + p);
      }
      for (p = 0; p < ncat; p++)
        NJC(tt, ((nclu_curr[tt] - 1) * ncat + p) * max_C + XCAT(i * ncat + p)) -= 1;
    }

    nclu_curr[tt] -= 1;

    /* :442-454 recompute loggamma of every other patient of the arm from eta/beta */
    int iii = 0;
    for (ii = 0; ii < c->nobs; ii++) {
      if (ii != i) {
        if (c->pb->data.treat[ii] == tt) {
          int clu_lg = LAB(tt, iii) - 1;
          for (k = 0; k < dim; k++)
            LG(ii, k) = o_calculate_gamma(c, &ETAS(0, clu_lg, tt), c->pb->data.z, c->nobs, k, ii, 1);
          iii += 1;
        }
      }
    }
    /* NOTE (faithful): `iii` is not advanced when ii==i, so patients after i in the arm read
     * the label one position early.  Only loggamma (not on the sweep's weights) is affected,
     * and every patient's loggamma is rewritten at its own reallocation (:825, :839). */
  }

  /* existing clusters, :460-602 */
  for (j = 0; j < nclu_curr[tt]; j++) {
    lgconY = 0.0; lgconN = 0.0; lgcatY = 0.0; lgcatN = 0.0;
    if (PPMx == 1) {
      for (p = 0; p < ncon; p++) {
        sumxtmp = SUMX(tt, j * ncon + p);
        sumx2tmp = SUMX2(tt, j * ncon + p);
        lgconN += o_gcon(c, sumxtmp, sumx2tmp, (int)NJ(tt, j));
        sumxtmp += XCON(i * ncon + p);
        sumx2tmp += XCON(i * ncon + p) * XCON(i * ncon + p);
        lgconY += o_gcon(c, sumxtmp, sumx2tmp, (int)NJ(tt, j) + 1);
      }
      for (p = 0; p < ncat; p++) {
        for (cidx = 0; cidx < max_C; cidx++) c->njctmp[cidx] = NJC(tt, (j * ncat + p) * max_C + cidx);
        lgcatN += o_gcat(c, p);
        c->njctmp[XCAT(i * ncat + p)] += 1;
        lgcatY += o_gcat(c, p);
      }
      gtilY[j] = lgconY + lgcatY;
      gtilN[j] = lgconN + lgcatN;
    }

    /* cohesion (:559-565 == :581-587) */
    if (cohesion == 1) weight[j] = log((double)NJ(tt, j));
    if (cohesion == 2) {
      weight[j] = o_dV(c, tt, nclu_curr[tt], idxs);
      weight[j] += lgamma(NJ(tt, j) + 1 - c->s->sigma[tt]) - lgamma(NJ(tt, j) - c->s->sigma[tt]);
    }
    if ((PPMx == 0) | ((calibration != 2) & (PPMx == 1))) {
      weight[j] += lgcatY - lgcatN + lgconY - lgconN;
    }
    if ((calibration == 2) & (PPMx == 1)) {
      if (coardegree == 1)
        weight[j] += (1 / ((double)ncon + (double)ncat)) * (lgcatY + lgconY - lgcatN - lgconN);
      if (coardegree == 2)
        weight[j] += (1 / (pow(((double)ncon + (double)ncat), 1.0 / 2.0))) *
                     (lgcatY + lgconY - lgcatN - lgconN);
    }
    weight[j] = o_lik(c, weight[j], &ETAS(0, j, tt), i);
  }

  /* auxiliary clusters, :605-715 */
  if (m->reuse == 0) {
    for (mm = 0; mm < CC; mm++)
      o_rng_normals(&c->rng, (uint32_t)l, TPPMX_SITE_AUX_PATIENT, (uint32_t)tt,
                    (uint32_t)(i * CC + mm), 0, dim, &ETAE(0, mm, tt));
  }
  for (j = nclu_curr[tt]; j < (nclu_curr[tt] + CC); j++) {
    jj = j - nclu_curr[tt];
    lgcondraw = 0.0; lgcatdraw = 0.0;
    if (PPMx == 1) {
      for (p = 0; p < ncon; p++) {
        tmp = XCON(i * ncon + p);
        lgcondraw += o_gcon(c, tmp, tmp * tmp, 1);
      }
      for (p = 0; p < ncat; p++) {
        for (cidx = 0; cidx < c->pb->shared.catvec[p]; cidx++) c->njctmp[cidx] = 0;
        c->njctmp[XCAT(i * ncat + p)] = 1;
        lgcatdraw += o_gcat(c, p);
      }
      gtilY[j] = lgcondraw + lgcatdraw;
      gtilN[j] = lgcondraw + lgcatdraw;
    }
    if (cohesion == 1) weight[j] = log(m->alpha) - log((double)CC);
    if (cohesion == 2)
      weight[j] = o_dV(c, tt, nclu_curr[tt], idxs);
    if ((PPMx == 0) | ((calibration != 2) & (PPMx == 1))) weight[j] += lgcondraw + lgcatdraw;
    if ((calibration == 2) & (PPMx == 1)) {
      if (coardegree == 1) weight[j] += (1 / ((double)ncon + (double)ncat)) * (lgcondraw + lgcatdraw);
      if (coardegree == 2)
        weight[j] += (1 / (pow(((double)ncon + (double)ncat), 1.0 / 2.0))) * (lgcondraw + lgcatdraw);
    }
    weight[j] = o_lik(c, weight[j], &ETAE(0, jj, tt), i);
  }

  /* calibration == 1, :717-788 */
  if ((calibration == 1) & (PPMx == 1)) {
    double maxgtilN = gtilN[0], maxgtilY = gtilY[0], sgY = 0.0, sgN = 0.0, lgtilNk, lgtilYk;
    for (j = 1; j < nclu_curr[tt] + CC; j++) {
      if (maxgtilN < gtilN[j]) maxgtilN = gtilN[j];
      if (j < nclu_curr[tt]) {
        if (maxgtilY < gtilY[j]) maxgtilY = gtilY[j];
      }
    }
    for (j = 0; j < nclu_curr[tt] + CC; j++) {
      lgtilN[j] = gtilN[j] - maxgtilN;
      sgN = sgN + exp(lgtilN[j]);
      if (j < nclu_curr[tt]) {
        lgtilY[j] = gtilY[j] - maxgtilY;
        sgY += exp(lgtilY[j]);
      }
    }
    for (j = 0; j < nclu_curr[tt]; j++) {
      lgtilNk = lgtilN[j] - log(sgN);
      lgtilYk = lgtilY[j] - log(sgY);
      if (cohesion == 1) weight[j] = log((double)NJ(tt, j));
      if (cohesion == 2) {
        weight[j] = o_dV(c, tt, nclu_curr[tt], idxs);
        weight[j] += lgamma(NJ(tt, j) + 1 - c->s->sigma[tt]) - lgamma(NJ(tt, j) - c->s->sigma[tt]);
      }
      weight[j] += lgtilYk - lgtilNk;
      weight[j] = o_lik(c, weight[j], &ETAS(0, j, tt), i);
    }
    for (j = nclu_curr[tt]; j < nclu_curr[tt] + CC; j++) {
      jj = j - nclu_curr[tt];
      lgtilNk = lgtilN[j] - log(sgN);
      if (cohesion == 1) weight[j] = log(m->alpha) - log((double)CC);
      if (cohesion == 2)
        weight[j] = o_dV(c, tt, nclu_curr[tt], idxs);
      weight[j] += lgtilNk;
      weight[j] = o_lik(c, weight[j], &ETAE(0, jj, tt), i);
    }
  }

  const int ntot = nclu_curr[tt] + CC;
  if (logw_out) memcpy(logw_out, weight, sizeof(double) * ntot);
  if (K_out) *K_out = nclu_curr[tt];

  /* :791-802 */
  double maxwei = weight[0], denwei = 0.0;
  for (j = 1; j < ntot; j++)
    if (maxwei < weight[j]) maxwei = weight[j];
  for (j = 0; j < ntot; j++) {
    weight[j] = exp(weight[j] - maxwei);
    denwei += weight[j];
  }
  for (j = 0; j < ntot; j++) pweight[j] = weight[j] / denwei;
  if (prob_out) memcpy(prob_out, pweight, sizeof(double) * ntot);
  if (score_only) return;

  /* :805-814 */
  double uu, u1, cweight = 0.0;
  o_rng_u2(&c->rng, (uint32_t)l, TPPMX_SITE_ALLOC_U, (uint32_t)tt, (uint32_t)i, 0, &uu, &u1);
  int newci = ntot;
  for (j = 0; j < ntot; j++) {
    cweight += pweight[j];
    if (uu < cweight) {
      newci = j + 1;
      break;
    }
  }

  /* :820-845 */
  if (newci <= nclu_curr[tt]) {
    LAB(tt, it) = newci;
    NJ(tt, LAB(tt, it) - 1) += 1;
    for (k = 0; k < dim; k++)
      LG(i, k) = o_calculate_gamma(c, &ETAS(0, LAB(tt, it) - 1, tt), c->pb->data.z, c->nobs, k, i, 1);
  } else {
    int id_empty = newci - nclu_curr[tt] - 1;
    nclu_curr[tt] += 1;
    LAB(tt, it) = nclu_curr[tt];
    NJ(tt, nclu_curr[tt] - 1) = 1;
    for (k = 0; k < dim; k++) ETAS(k, nclu_curr[tt] - 1, tt) = ETAE(k, id_empty, tt);
    for (k = 0; k < dim; k++)
      LG(i, k) = o_calculate_gamma(c, &ETAS(0, LAB(tt, it) - 1, tt), c->pb->data.z, c->nobs, k, i, 1);
    o_rng_normals(&c->rng, (uint32_t)l, TPPMX_SITE_AUX_REFRESH, (uint32_t)tt, (uint32_t)i, 0, dim,
                  &ETAE(0, id_empty, tt));
  }

  /* :847-856 */
  if (PPMx == 1) {
    for (p = 0; p < ncon; p++) {
      SUMX(tt, (LAB(tt, it) - 1) * ncon + p) += XCON(i * ncon + p);
      SUMX2(tt, (LAB(tt, it) - 1) * ncon + p) += XCON(i * ncon + p) * XCON(i * ncon + p);
    }
    for (p = 0; p < ncat; p++) NJC(tt, ((LAB(tt, it) - 1) * ncat + p) * max_C + XCAT(i * ncat + p)) += 1;
  }
}

/* the per-arm part of one iteration: :349-860 */
void o_sweep_iter(o_ctx *c, int l) {
  const tppmx_model *m = &c->pb->model;
  for (int tt = 0; tt < c->T; tt++) {
    if (m->reuse == 1) { /* :351-360 */
      for (int mm = 0; mm < c->CC; mm++)
        o_rng_normals(&c->rng, (uint32_t)l, TPPMX_SITE_AUX_SWEEP, (uint32_t)tt, (uint32_t)mm, 0,
                      c->D, &ETAE(0, mm, tt));
    }
    int it = 0;
    for (int i = 0; i < c->nobs; i++) {
      if (c->pb->data.treat[i] == tt) {
        o_realloc_patient(c, tt, it, i, l, 0, NULL, NULL, NULL);
        it += 1;
      }
    }
  }
}

/* ---------------------------------------------------------------------------------------
 * state helpers
 * ------------------------------------------------------------------------------------- */
static size_t o_cnt(const tppmx_dims *d, int CC, int which) {
  size_t T = d->T, nt = d->ntmax, P = d->P, D = d->D, Q = d->Q, n = d->nobs;
  switch (which) {
    case 0: return T * nt;                                   /* labels */
    case 1: return T * nt;                                   /* nj */
    case 2: return T;                                        /* nclu */
    case 3: return T * nt * (P > 0 ? P : 1);                 /* sumx */
    case 4: return T * nt * (size_t)(d->ncat > 0 ? d->ncat * d->maxC : 1); /* njc */
    case 5: return D * nt * T;                               /* eta_star */
    case 6: return D * (size_t)CC * T;                       /* eta_empty */
    case 7: return n * D;                                    /* JJ / loggamma */
    case 8: return n;                                        /* ss */
    case 9: return Q * D;                                    /* beta, lambda, sigma_beta */
    case 10: return D * T;                                   /* Theta */
    case 11: return D * D * T;                               /* Sigma */
  }
  return 0;
}

/* deep copy of the fields the sweep touches, so score_patient leaves `s` untouched */
static int o_state_clone(const oracle_problem *pb, const tppmx_state *s, tppmx_state *o) {
  const tppmx_dims *d = &pb->dims;
  int CC = pb->model.CC;
  memset(o, 0, sizeof(*o));
#define CL(field, type, which)                                           \
  do {                                                                   \
    size_t nb = sizeof(type) * o_cnt(d, CC, which);                      \
    o->field = (type *)malloc(nb);                                       \
    if (!o->field) return 1;                                             \
    if (s->field) memcpy(o->field, s->field, nb); else memset(o->field, 0, nb); \
  } while (0)
  CL(labels, int32_t, 0); CL(nj, int32_t, 1); CL(nclu, int32_t, 2);
  CL(sumx, double, 3); CL(sumx2, double, 3); CL(njc, int32_t, 4);
  CL(eta_star, double, 5); CL(eta_empty, double, 6);
  CL(JJ, double, 7); CL(ss, double, 8); CL(loggamma, double, 7); CL(beta, double, 9);
#undef CL
  o->sigma = (double *)malloc(sizeof(double) * d->T);
  o->kappa = (double *)malloc(sizeof(double) * d->T);
  o->idxs = (int32_t *)malloc(sizeof(int32_t));
  memcpy(o->sigma, s->sigma, sizeof(double) * d->T);
  if (s->kappa) memcpy(o->kappa, s->kappa, sizeof(double) * d->T);
  o->idxs[0] = s->idxs[0];
  return 0;
}
static void o_state_free(tppmx_state *o) {
  free(o->labels); free(o->nj); free(o->nclu); free(o->sumx); free(o->sumx2); free(o->njc);
  free(o->eta_star); free(o->eta_empty); free(o->JJ); free(o->ss); free(o->loggamma);
  free(o->beta); free(o->sigma); free(o->kappa); free(o->idxs);
}

/* ---------------------------------------------------------------------------------------
 * public entry points
 * ------------------------------------------------------------------------------------- */

/* src/dm_ppmx_ct.cpp:49-236.  The reference fills eta_star_curr / eta_empty with
 * arma::fill::randu and ss with R::rgamma from R's RNG; here the same distributions come
 * from the Philox contract. */
int oracle_init_state(const oracle_problem *pb, tppmx_state *s, uint64_t seed, uint32_t chain_id,
                      const int32_t *init_labels, const int32_t *init_nclu,
                      const double *initbeta) {
  o_ctx cc, *c = &cc;
  if (o_ctx_open(c, pb, s, seed, chain_id)) return 1;
  const int nobs = c->nobs, dim = c->D, T = c->T, ntmax = c->ntmax, ncon = c->P, ncat = c->ncat;
  const int max_C = c->maxC, CC = c->CC;
  int i, k, tt, ii, p, j;

  s->idxs[0] = 0; /* :49 */
  for (tt = 0; tt < T; tt++) { /* :52-55 */
    s->sigma[tt] = pb->shared.grid[0 + 2 * 0];
    s->kappa[tt] = pb->shared.grid[1 + 2 * 0];
  }

  /* :75-84 JJ, TT; :88-91 ss ~ Gamma(ypiu=1, scale 1/TT) */
  for (i = 0; i < nobs; i++) {
    double TTi = 0.0;
    for (k = 0; k < dim; k++) {
      JJm(i, k) = Ym(i, k);
      if (JJm(i, k) < pow(10.0, -100.0)) JJm(i, k) = pow(10.0, -100.0);
      TTi += JJm(i, k);
    }
    s->ss[i] = o_rng_gamma(&c->rng, 0, TPPMX_SITE_INIT_SS, 0, (uint32_t)i, 0, 1.0) * (1.0 / TTi);
  }

  /* :113-115 */
  memcpy(s->labels, init_labels, sizeof(int32_t) * T * ntmax);
  memcpy(s->nclu, init_nclu, sizeof(int32_t) * T);
  memset(s->nj, 0, sizeof(int32_t) * T * ntmax);
  for (tt = 0; tt < T; tt++)
    for (ii = 0; ii < c->num_treat[tt]; ii++) NJ(tt, LAB(tt, ii) - 1) += 1; /* card_cluster */

  /* :140 eta_star_curr ~ U(0,1), whole cube */
  for (tt = 0; tt < T; tt++)
    for (j = 0; j < ntmax; j++)
      for (k = 0; k < dim; k++) {
        double u0, u1;
        o_rng_u2(&c->rng, 0, TPPMX_SITE_INIT_ETA, (uint32_t)tt, (uint32_t)(j * dim + k), 0, &u0, &u1);
        ETAS(k, j, tt) = u0;
      }

  /* :149-155 */
  for (k = 0; k < c->Q * dim; k++) {
    s->beta[k] = initbeta[k];
    if (s->sigma_beta) s->sigma_beta[k] = 1.0;
    if (s->lambda) s->lambda[k] = 1.0;
  }
  if (s->tau) s->tau[0] = 1.0;

  /* :161-173 loggamma */
  for (tt = 0; tt < T; tt++) {
    ii = 0;
    for (i = 0; i < nobs; i++) {
      if (pb->data.treat[i] == tt) {
        int clu_lg = LAB(tt, ii) - 1;
        for (k = 0; k < dim; k++)
          LG(i, k) = o_calculate_gamma(c, &ETAS(0, clu_lg, tt), pb->data.z, nobs, k, i, 1);
        ii += 1;
      }
    }
  }

  /* :179-204 sufficient statistics */
  memset(s->sumx, 0, sizeof(double) * o_cnt(&pb->dims, CC, 3));
  memset(s->sumx2, 0, sizeof(double) * o_cnt(&pb->dims, CC, 3));
  memset(s->njc, 0, sizeof(int32_t) * o_cnt(&pb->dims, CC, 4));
  if (pb->model.PPMx == 1) {
    for (tt = 0; tt < T; tt++) {
      ii = 0;
      for (i = 0; i < nobs; i++) {
        if (pb->data.treat[i] == tt) {
          for (p = 0; p < ncon; p++) {
            SUMX(tt, (LAB(tt, ii) - 1) * ncon + p) += XCON(i * ncon + p);
            SUMX2(tt, (LAB(tt, ii) - 1) * ncon + p) += XCON(i * ncon + p) * XCON(i * ncon + p);
          }
          for (p = 0; p < ncat; p++) NJC(tt, ((LAB(tt, ii) - 1) * ncat + p) * max_C + XCAT(i * ncat + p)) += 1;
          ii += 1;
        }
      }
    }
  }

  /* :210 eta_empty ~ U(0,1) */
  for (tt = 0; tt < T; tt++)
    for (j = 0; j < CC; j++)
      for (k = 0; k < dim; k++) {
        double u0, u1;
        o_rng_u2(&c->rng, 0, TPPMX_SITE_INIT_EMPTY, (uint32_t)tt, (uint32_t)(j * dim + k), 0, &u0, &u1);
        ETAE(k, j, tt) = u0;
      }

  /* :271-278 Theta = 0, Sigma = I */
  if (s->Theta) memset(s->Theta, 0, sizeof(double) * dim * T);
  if (s->Sigma) {
    memset(s->Sigma, 0, sizeof(double) * dim * dim * T);
    for (tt = 0; tt < T; tt++)
      for (k = 0; k < dim; k++) s->Sigma[k + dim * (k + dim * tt)] = 1.0;
  }
  o_ctx_close(c);
  return 0;
}

int oracle_score_patient(const oracle_problem *pb, const tppmx_state *s, int32_t arm, int32_t it,
                         uint64_t seed, uint32_t chain_id, int32_t iter, double *logw_out,
                         double *prob_out, int32_t *K_out) {
  tppmx_state tmp;
  if (o_state_clone(pb, s, &tmp)) return 1;
  o_ctx cc, *c = &cc;
  if (o_ctx_open(c, pb, &tmp, seed, chain_id)) return 1;
  /* patient index of position `it` in arm `arm` */
  int i = -1, cnt = 0;
  for (int r = 0; r < c->nobs; r++)
    if (pb->data.treat[r] == arm) {
      if (cnt == it) { i = r; break; }
      cnt++;
    }
  if (i < 0) { o_ctx_close(c); o_state_free(&tmp); return 2; }
  int K = 0;
  o_realloc_patient(c, arm, it, i, iter, 1, logw_out, prob_out, &K);
  if (K_out) *K_out = K;
  o_ctx_close(c);
  o_state_free(&tmp);
  return 0;
}

int oracle_sweep(const oracle_problem *pb, tppmx_state *s, uint64_t seed, uint32_t chain_id,
                 int32_t iter0, int32_t n_sweeps) {
  o_ctx cc, *c = &cc;
  if (o_ctx_open(c, pb, s, seed, chain_id)) return 1;
  for (int l = iter0; l < iter0 + n_sweeps; l++) o_sweep_iter(c, l);
  o_ctx_close(c);
  return 0;
}

/* One sweep (iteration l) that also records, per step, the normalised allocation probabilities
 * pweight (:791-802) and the number of occupied clusters after the removal: what the device's
 * probe instantiation of the production kernel is compared with (tppmx_batch_sweep_probe).
 * prob_out [T][ntmax][stride], K_out [T][ntmax]; entries beyond min(K+CC, stride) are left. */
int oracle_sweep_probe(const oracle_problem *pb, tppmx_state *s, uint64_t seed, uint32_t chain_id,
                       int32_t l, int32_t stride, double *prob_out, int32_t *K_out) {
  o_ctx cc, *c = &cc;
  if (o_ctx_open(c, pb, s, seed, chain_id)) return 1;
  const tppmx_model *m = &pb->model;
  double *row = (double *)malloc(sizeof(double) * (size_t)(c->ntmax + c->CC));
  for (int tt = 0; tt < c->T; tt++) {
    if (m->reuse == 1)
      for (int mm = 0; mm < c->CC; mm++)
        o_rng_normals(&c->rng, (uint32_t)l, TPPMX_SITE_AUX_SWEEP, (uint32_t)tt, (uint32_t)mm, 0, c->D, &ETAE(0, mm, tt));
    int it = 0;
    for (int i = 0; i < c->nobs; i++) {
      if (pb->data.treat[i] != tt) continue;
      int K = 0;
      o_realloc_patient(c, tt, it, i, l, 0, NULL, row, &K);
      const int ntot = K + c->CC, ncp = ntot < stride ? ntot : stride;
      memcpy(prob_out + ((size_t)tt * c->ntmax + it) * stride, row, sizeof(double) * ncp);
      K_out[(size_t)tt * c->ntmax + it] = K;
      it += 1;
    }
  }
  free(row);
  o_ctx_close(c);
  return 0;
}

/* ---------------------------------------------------------------------------------------
 * posterior predictive, src/dm_ppmx_ct.cpp:1081-1384
 * ------------------------------------------------------------------------------------- */
void o_chol_lower(const double *A, int n, double *L) { /* A col-major sym pos-def */
  memset(L, 0, sizeof(double) * n * n);
  for (int i = 0; i < n; i++)
    for (int j = 0; j <= i; j++) {
      double sum = A[i + n * j];
      for (int k = 0; k < j; k++) sum -= L[i + n * k] * L[j + n * k];
      if (i == j) L[i + n * i] = sqrt(sum);
      else L[i + n * j] = sum / L[j + n * j];
    }
}

int oracle_predict(const oracle_problem *pb, const tppmx_state *s_in, uint64_t seed,
                   uint32_t chain_id, int32_t iter, double *pipred, double *ypred,
                   int32_t *predclass) {
  o_ctx cc, *c = &cc;
  if (o_ctx_open(c, pb, (tppmx_state *)s_in, seed, chain_id)) return 1; /* read-only use */
  const tppmx_model *m = &pb->model;
  const int ncon = c->P, ncat = c->ncat, max_C = c->maxC, dim = c->D, CC = c->CC, npred = c->npred;
  const int PPMx = m->PPMx, cohesion = m->cohesion, calibration = m->calibration;
  const int coardegree = m->coardegree;
  const int *nclu_curr = s_in->nclu;
  const int idxs = s_in->idxs[0];
  const double *xconp = pb->data.xconp;
  const int32_t *xcatp = pb->data.xcatp;
  double *weight = c->weight, *pweight = c->pweight;
  double *gtilN = c->gtilN, *gtilY = c->gtilY, *lgtilN = c->lgtilN, *lgtilY = c->lgtilY;
  int tt, pp, j, p, cidx, k;
  int newci = 0; /* quirk 7: NOT reset per patient (:1351-1360); the reference leaves it
                    uninitialised before the first draw, 0 here */
  double eta_pred[TPPMX_MAXD], loggamma_pred[TPPMX_MAXD], L[TPPMX_MAXD * TPPMX_MAXD];

  for (tt = 0; tt < c->T; tt++) {
    for (pp = 0; pp < npred; pp++) {
      for (j = 0; j < nclu_curr[tt]; j++) {
        double lgconN = 0.0, lgconY = 0.0, lgcatN = 0.0, lgcatY = 0.0;
        if (PPMx == 1) {
          for (p = 0; p < ncon; p++) {
            double sumxtmp = SUMX(tt, j * ncon + p);
            double sumx2tmp = SUMX2(tt, j * ncon + p);
            lgconN += o_gcon(c, sumxtmp, sumx2tmp, (int)NJ(tt, j));
            sumxtmp += xconp[pp * ncon + p];
            sumx2tmp += xconp[pp * ncon + p] * xconp[pp * ncon + p];
            lgconY += o_gcon(c, sumxtmp, sumx2tmp, (int)NJ(tt, j) + 1);
          }
          for (p = 0; p < ncat; p++) {
            for (cidx = 0; cidx < max_C; cidx++) c->njctmp[cidx] = NJC(tt, (j * ncat + p) * max_C + cidx);
            lgcatN += o_gcat(c, p);
            /* quirk 6 (:1157): indexes the TRAINING xcat with the prediction index */
            c->njctmp[XCAT(pp * ncat + p)] += 1;
            lgcatY += o_gcat(c, p);
          }
          gtilY[j] = lgconY + lgcatY;
          gtilN[j] = lgconN + lgcatN;
        }
        if (cohesion == 1) weight[j] = log((double)NJ(tt, j));
        if (cohesion == 2) {
          weight[j] = o_dV(c, tt, nclu_curr[tt], idxs);
          weight[j] += lgamma(NJ(tt, j) + 1 - s_in->sigma[tt]) - lgamma(NJ(tt, j) - s_in->sigma[tt]);
        }
        if ((PPMx == 0) | ((calibration != 2) & (PPMx == 1))) weight[j] += lgcatY - lgcatN + lgconY - lgconN;
        if ((calibration == 2) & (PPMx == 1)) {
          if (coardegree == 1)
            weight[j] += (1 / ((double)ncon + (double)ncat)) * (lgcatY + lgconY - lgcatN - lgconN);
          if (coardegree == 2)
            weight[j] += (1 / (pow(((double)ncon + (double)ncat), 1.0 / 2.0))) *
                         (lgcatY + lgconY - lgcatN - lgconN);
        }
      }
      for (j = nclu_curr[tt]; j < nclu_curr[tt] + CC; j++) {
        double lgcondraw = 0.0, lgcatdraw = 0.0;
        if (PPMx == 1) {
          for (p = 0; p < ncon; p++) {
            double tmp = xconp[pp * ncon + p];
            lgcondraw += o_gcon(c, tmp, tmp * tmp, 1);
          }
          for (p = 0; p < ncat; p++) {
            for (cidx = 0; cidx < pb->shared.catvec[p]; cidx++) c->njctmp[cidx] = 0;
            c->njctmp[xcatp[pp * ncat + p]] = 1;
            lgcatdraw += o_gcat(c, p);
          }
          gtilY[j] = lgcondraw + lgcatdraw;
          gtilN[j] = lgcondraw + lgcatdraw;
        }
        if (cohesion == 1) weight[j] = log(m->alpha) - log((double)CC);
        if (cohesion == 2)
          weight[j] = o_dV(c, tt, nclu_curr[tt], idxs);
        if ((PPMx == 0) | ((calibration != 2) & (PPMx == 1))) weight[j] += lgcondraw + lgcatdraw;
        if ((calibration == 2) & (PPMx == 1)) {
          if (coardegree == 1) weight[j] += (1 / ((double)ncon + (double)ncat)) * (lgcondraw + lgcatdraw);
          if (coardegree == 2)
            weight[j] += (1 / (pow(((double)ncon + (double)ncat), 1.0 / 2.0))) * (lgcondraw + lgcatdraw);
        }
      }
      if ((calibration == 1) & (PPMx == 1)) { /* :1282-1334 */
        double maxgtilN = gtilN[0], maxgtilY = gtilY[0], sgY = 0.0, sgN = 0.0;
        for (j = 1; j < nclu_curr[tt] + CC; j++) {
          if (maxgtilN < gtilN[j]) maxgtilN = gtilN[j];
          if (j < nclu_curr[tt])
            if (maxgtilY < gtilY[j]) maxgtilY = gtilY[j];
        }
        for (j = 0; j < nclu_curr[tt] + CC; j++) {
          lgtilN[j] = gtilN[j] - maxgtilN;
          sgN = sgN + exp(lgtilN[j]);
          if (j < nclu_curr[tt]) {
            lgtilY[j] = gtilY[j] - maxgtilY;
            sgY = sgY + exp(lgtilY[j]);
          }
        }
        for (j = 0; j < nclu_curr[tt]; j++) {
          double lgtilNk = lgtilN[j] - log(sgN), lgtilYk = lgtilY[j] - log(sgY);
          if (cohesion == 1) weight[j] = log((double)NJ(tt, j));
          if (cohesion == 2) {
            weight[j] = o_dV(c, tt, nclu_curr[tt], idxs);
            weight[j] += lgamma(NJ(tt, j) + 1 - s_in->sigma[tt]) - lgamma(NJ(tt, j) - s_in->sigma[tt]);
          }
          weight[j] += lgtilYk - lgtilNk;
        }
        for (j = nclu_curr[tt]; j < nclu_curr[tt] + CC; j++) {
          double lgtilNk = lgtilN[j] - log(sgN);
          if (cohesion == 1) weight[j] = log(m->alpha) - log((double)CC);
          if (cohesion == 2)
            weight[j] = o_dV(c, tt, nclu_curr[tt], idxs);
          weight[j] += lgtilNk;
        }
      }

      /* :1337-1348 */
      const int ntot = nclu_curr[tt] + CC;
      double maxwei = weight[0], denwei = 0.0;
      for (j = 1; j < ntot; j++)
        if (maxwei < weight[j]) maxwei = weight[j];
      for (j = 0; j < ntot; j++) {
        weight[j] = exp(weight[j] - maxwei);
        denwei += weight[j];
      }
      for (j = 0; j < ntot; j++) pweight[j] = weight[j] / denwei;

      /* :1351-1360 */
      double uu, u1, cweight = 0.0;
      o_rng_u2(&c->rng, (uint32_t)iter, TPPMX_SITE_PRED_U, (uint32_t)tt, (uint32_t)pp, 0, &uu, &u1);
      for (j = 0; j < ntot; j++) {
        cweight += pweight[j];
        if (uu < cweight) {
          newci = j + 1;
          break;
        }
      }

      /* :1366-1370 */
      if (newci <= nclu_curr[tt]) {
        for (k = 0; k < dim; k++) eta_pred[k] = ETAS(k, newci - 1, tt);
      } else { /* arma::mvnrnd(Theta.col(tt), Sigma.slice(tt)) = Theta + chol_lower(Sigma) * N(0,I) */
        double zn[TPPMX_MAXD];
        o_rng_normals(&c->rng, (uint32_t)iter, TPPMX_SITE_PRED_ETA, (uint32_t)tt, (uint32_t)pp, 0, dim, zn);
        o_chol_lower(&s_in->Sigma[dim * dim * tt], dim, L);
        for (k = 0; k < dim; k++) {
          double acc = s_in->Theta[k + dim * tt];
          for (int r = 0; r <= k; r++) acc += L[k + dim * r] * zn[r];
          eta_pred[k] = acc;
        }
      }

      /* :1373-1379 — no max-shift, as in the reference */
      double sumrow = 0.0;
      for (k = 0; k < dim; k++) {
        loggamma_pred[k] = o_calculate_gamma(c, eta_pred, pb->data.zpred, npred, k, pp, 1);
        sumrow += exp(loggamma_pred[k]);
      }
      double pi_row[TPPMX_MAXD];
      for (k = 0; k < dim; k++) {
        pi_row[k] = exp(loggamma_pred[k]) / sumrow;
        if (pipred) pipred[pp + npred * (k + dim * tt)] = pi_row[k];
      }

      /* :1381 rmultinom(1, pi): one categorical draw (R draws successive binomials from the
       * same law; a single inverse-CDF uniform is used here and on device) */
      double uy, uy1, cw = 0.0;
      o_rng_u2(&c->rng, (uint32_t)iter, TPPMX_SITE_PRED_Y, (uint32_t)tt, (uint32_t)pp, 0, &uy, &uy1);
      int pick = dim - 1;
      for (k = 0; k < dim; k++) {
        cw += pi_row[k];
        if (uy < cw) { pick = k; break; }
      }
      if (ypred)
        for (k = 0; k < dim; k++) ypred[pp + npred * (k + dim * tt)] = (k == pick) ? 1.0 : 0.0;
      if (predclass) predclass[pp + npred * tt] = newci; /* :1382 */
    }
  }
  o_ctx_close(c);
  return 0;
}
