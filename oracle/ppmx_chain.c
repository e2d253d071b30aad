/*
 * ppmx_chain.c — CPU oracle of the per-iteration updates that follow the reallocation sweep.
 * TEST INFRASTRUCTURE ONLY (see ppmx_oracle.h).
 *
 * Restates, with the reference's loop structure, evaluation order and quirks:
 *   eta* Metropolis step per cluster     /root/reference/src/utils_ct.cpp:579-702
 *                                        (called at src/dm_ppmx_ct.cpp:866-878)
 *   (Theta, Sigma) hierarchy             /root/reference/src/dm_ppmx_ct.cpp:880-902
 *   (sigma, kappa) grid draw             /root/reference/src/dm_ppmx_ct.cpp:907-1007
 *   beta Metropolis step + horseshoe     /root/reference/src/utils_ct.cpp:710-800, 876-915
 *                                        (called at src/dm_ppmx_ct.cpp:1009-1034)
 *   JJ / ss gamma draws                  /root/reference/src/dm_ppmx_ct.cpp:1042-1055
 * Random numbers follow the Philox contract of include/tppmx.h; where the reference calls a
 * library sampler (arma::wishrnd, arma::mvnrnd, R::rgamma, Rcpp sample) the same law is drawn
 * with a published algorithm (Bartlett, Cholesky, Marsaglia-Tsang, inverse CDF) that the device
 * mirrors draw for draw.  R::pgamma / R::qgamma -> csrc/incgamma.h.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "../treatppmx_b200/csrc/incgamma.h"
#include "oracle_internal.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ---- small dense helpers (D <= TPPMX_MAXD), column-major ------------------------------------ */
static void m_inv_spd(const double *A, int n, double *Ainv) { /* inverse through the Cholesky factor */
  double L[TPPMX_MAXD * TPPMX_MAXD], Li[TPPMX_MAXD * TPPMX_MAXD];
  o_chol_lower(A, n, L);
  memset(Li, 0, sizeof(Li));
  for (int c = 0; c < n; c++) { /* Li = L^-1 by forward substitution, column by column */
    for (int r = 0; r < n; r++) {
      double s = (r == c) ? 1.0 : 0.0;
      for (int k = 0; k < r; k++) s -= L[r + n * k] * Li[k + n * c];
      Li[r + n * c] = s / L[r + n * r];
    }
  }
  for (int r = 0; r < n; r++)
    for (int c = 0; c < n; c++) {
      double s = 0.0;
      for (int k = 0; k < n; k++) s += Li[k + n * r] * Li[k + n * c];
      Ainv[r + n * c] = s;
    }
}

/* dmvnrm(x, mean, sigma, logd = true), utils_ct.cpp:242-266:
 *   rooti = inv(chol_upper(sigma))' = inv(chol_lower(sigma));  z = rooti (x - mean)
 *   out = -(d/2) log(2 pi) - 0.5 z'z + sum(log(diag(rooti))) */
static double o_dmvnrm_log(const double *x, const double *mean, const double *sigma, int n) {
  double L[TPPMX_MAXD * TPPMX_MAXD], z[TPPMX_MAXD];
  o_chol_lower(sigma, n, L);
  double rootisum = 0.0, zz = 0.0;
  for (int r = 0; r < n; r++) {
    double s = x[r] - mean[r];
    for (int k = 0; k < r; k++) s -= L[r + n * k] * z[k];
    z[r] = s / L[r + n * r];
    zz += z[r] * z[r];
    rootisum += log(1.0 / L[r + n * r]);
  }
  const double log2pi = log(2.0 * M_PI);
  return -(double)n / 2.0 * log2pi - 0.5 * zz + rootisum;
}

/* ---- eta_update, utils_ct.cpp:579-702 ---------------------------------------------------------- */
static void o_eta_update(o_ctx *c, int t, int jj, int l, int32_t *eta_acc) {
  const int dim = c->D, nobs = c->nobs;
  const double mhtune = c->pb->model.mhtune_eta;
  const double *mu_star = &c->s->Theta[dim * t];
  const double *sigma_star = &c->s->Sigma[dim * dim * t];
  double eta[TPPMX_MAXD], eta_p[TPPMX_MAXD], nrm[TPPMX_MAXD];
  int i, k, it;
  double log_num, log_den, ln_acp, lnu;

  for (k = 0; k < dim; k++) eta[k] = ETAS(k, jj, t);

  log_den = 0.0;
  it = 0;
  for (i = 0; i < nobs; i++) {
    if (c->pb->data.treat[i] == t) {
      if (LAB(t, it) == (jj + 1)) {
        for (k = 0; k < dim; k++) log_den -= lgamma(exp(LG(i, k))) + exp(LG(i, k)) * log(JJm(i, k));
      }
      it += 1;
    }
  }
  log_den += o_dmvnrm_log(eta, mu_star, sigma_star, dim);

  /* :642-644 proposal eta + rnorm(0, mhtune) */
  o_rng_normals(&c->rng, (uint32_t)l, TPPMX_SITE_ETA_PROP, (uint32_t)t, (uint32_t)jj, 0, dim, nrm);
  for (k = 0; k < dim; k++) eta_p[k] = eta[k] + mhtune * nrm[k];

  log_num = 0.0;
  it = 0;
  for (i = 0; i < nobs; i++) {
    if (c->pb->data.treat[i] == t) {
      if ((LAB(t, it) - 1) == jj) {
        for (k = 0; k < dim; k++) {
          const double lgp = LG(i, k) - eta[k] + eta_p[k]; /* :651 */
          log_num -= lgamma(exp(lgp)) + exp(lgp) * log(JJm(i, k));
        }
      }
      it += 1;
    }
  }
  /* quirk 9 (:671): the proposal's prior density is added to log_den, not log_num */
  log_den += o_dmvnrm_log(eta_p, mu_star, sigma_star, dim);

  ln_acp = log_num - log_den;
  double u0, u1;
  o_rng_u2(&c->rng, (uint32_t)l, TPPMX_SITE_ETA_PROP, (uint32_t)t, (uint32_t)jj, 0x100u, &u0, &u1);
  lnu = log(u0);

  if (lnu < ln_acp) {
    if (eta_acc) eta_acc[t + c->T * jj] += 1;
    it = 0;
    for (i = 0; i < nobs; i++) {
      if (c->pb->data.treat[i] == t) {
        if ((LAB(t, it) - 1) == jj)
          for (k = 0; k < dim; k++) LG(i, k) = LG(i, k) - eta[k] + eta_p[k];
        it += 1;
      }
    }
    for (k = 0; k < dim; k++) ETAS(k, jj, t) = eta_p[k];
  }
}

/* Lambda ~ Wishart(S, df) (arma::wishrnd, :895), Bartlett: W = L A A' L', L = chol_lower(S),
 * A lower triangular, A_ii = sqrt(chi2(df - i)), A_ij ~ N(0,1) (i > j) */
static void o_wishart(const o_rng *rng, int l, int tt, const double *S, double df, int dim, double *W) {
  double L[TPPMX_MAXD * TPPMX_MAXD], A[TPPMX_MAXD * TPPMX_MAXD], LA[TPPMX_MAXD * TPPMX_MAXD];
  int i, k, r;
  o_chol_lower(S, dim, L);
  memset(A, 0, sizeof(A));
  for (k = 0; k < dim; k++) {
    const double g = o_rng_gamma(rng, (uint32_t)l, TPPMX_SITE_HIER_W, (uint32_t)tt, (uint32_t)k, 0, 0.5 * (df - (double)k));
    A[k + dim * k] = sqrt(2.0 * g);
    for (r = 0; r < k; r++) {
      double n0, n1;
      o_rng_n2(rng, (uint32_t)l, TPPMX_SITE_HIER_W, (uint32_t)tt, (uint32_t)(1000 + k * dim + r), 0, &n0, &n1);
      A[k + dim * r] = n0;
    }
  }
  for (k = 0; k < dim; k++)
    for (r = 0; r < dim; r++) {
      double acc = 0.0;
      for (i = 0; i < dim; i++) acc += L[k + dim * i] * A[i + dim * r];
      LA[k + dim * r] = acc;
    }
  for (k = 0; k < dim; k++)
    for (r = 0; r < dim; r++) {
      double acc = 0.0;
      for (i = 0; i < dim; i++) acc += LA[k + dim * i] * LA[r + dim * i];
      W[k + dim * r] = acc;
    }
}

/* ---- hierarchy, src/dm_ppmx_ct.cpp:880-902 ------------------------------------------------------
 * Lambda0 is declared without initialisation and only its diagonal is set (:266-269); the
 * off-diagonal is taken as 0 here. */
static void o_hierarchy(o_ctx *c, int tt, int l) {
  const tppmx_model *m = &c->pb->model;
  const int dim = c->D, nt = c->num_treat[tt];
  const double nu0_theta = m->hP0_nu0, s0_lambda = m->hP0_s0;
  double etamean[TPPMX_MAXD], etacov[TPPMX_MAXD * TPPMX_MAXD], splp[TPPMX_MAXD * TPPMX_MAXD] = {0};
  double Lambda[TPPMX_MAXD * TPPMX_MAXD], tmp[TPPMX_MAXD * TPPMX_MAXD], sptp[TPPMX_MAXD * TPPMX_MAXD];
  double fptp[TPPMX_MAXD];
  int i, k, r;
  const double dn = (double)nt;

  for (k = 0; k < dim; k++) { /* arma::mean(..., 1) */
    double acc = 0.0;
    for (i = 0; i < nt; i++) acc += ETAS(k, LAB(tt, i) - 1, tt);
    etamean[k] = acc / dn;
  }
  for (k = 0; k < dim; k++) /* arma::cov(X.t(), 1): normalised by N */
    for (r = 0; r < dim; r++) {
      double acc = 0.0;
      for (i = 0; i < nt; i++)
        acc += (ETAS(k, LAB(tt, i) - 1, tt) - etamean[k]) * (ETAS(r, LAB(tt, i) - 1, tt) - etamean[r]);
      etacov[k + dim * r] = acc / dn;
    }
  const double fplp = s0_lambda + (dn * .5);
  for (k = 0; k < dim; k++)
    for (r = 0; r < dim; r++)
      splp[k + dim * r] = ((k == r) ? m->hP0_Lambda0 : 0.0) +
                          (dn * .5) * (etacov[k + dim * r] + (nu0_theta / (nu0_theta + dn)) *
                                                                 ((etamean[k] - m->hP0_mu0[k]) * (etamean[r] - m->hP0_mu0[r])));

  o_wishart(&c->rng, l, tt, splp, fplp, dim, Lambda);
  /* Theta ~ N(fptp, inv(Lambda (n + nu0))) (arma::mvnrnd = mean + chol_lower(cov) z) */
  for (k = 0; k < dim; k++) fptp[k] = (m->hP0_mu0[k] * nu0_theta + dn * etamean[k]) / (nu0_theta + dn);
  for (k = 0; k < dim * dim; k++) tmp[k] = Lambda[k] * (dn + nu0_theta);
  m_inv_spd(tmp, dim, sptp);
  double Ls[TPPMX_MAXD * TPPMX_MAXD], zn[TPPMX_MAXD];
  o_chol_lower(sptp, dim, Ls);
  o_rng_normals(&c->rng, (uint32_t)l, TPPMX_SITE_HIER_THETA, (uint32_t)tt, 0, 0, dim, zn);
  for (k = 0; k < dim; k++) {
    double acc = fptp[k];
    for (r = 0; r <= k; r++) acc += Ls[k + dim * r] * zn[r];
    c->s->Theta[k + dim * tt] = acc;
  }
  m_inv_spd(Lambda, dim, &c->s->Sigma[dim * dim * tt]);
}

/* ---- (sigma, kappa), src/dm_ppmx_ct.cpp:907-1007 ------------------------------------------------
 * omegasigma (the clusters' similarity) does not depend on the grid point and cancels in the
 * normalisation; it is kept (with the correct indices: the reference's double-dipper branch
 * reads sumx(tt,j) / nj_curr(tt,j*ncon+p), quirk 8).  The categorical branch of the reference
 * cannot run (passes the whole njc matrix, :942) -> ncat must be 0 here.  Rcpp's sample() walks
 * the probabilities sorted in decreasing order; one inverse-CDF uniform in grid order draws
 * from the same law. */
static int o_sigma_kappa(o_ctx *c, int tt, int l, double *os) {
  const tppmx_model *m = &c->pb->model;
  const int ncon = c->P, K = c->G;
  int j, p, g;
  if (c->ncat > 0) return 1;
  const int nclu = c->s->nclu[tt];
  double omega = 0.0;
  for (j = 0; j < nclu; j++) {
    double gt = 0.0;
    if (m->PPMx == 1)
      for (p = 0; p < ncon; p++) gt += o_gcon(c, SUMX(tt, j * ncon + p), SUMX2(tt, j * ncon + p), (int)NJ(tt, j));
    if ((m->calibration == 2) & (m->PPMx == 1)) {
      if (m->coardegree == 1) gt = (1 / ((double)ncon + (double)c->ncat)) * gt;
      if (m->coardegree == 2) gt = (1 / (pow(((double)ncon + (double)c->ncat), 1.0 / 2.0))) * gt;
    }
    omega += gt;
  }
  for (g = 0; g < K; g++) {
    const double sigma_temp = c->pb->shared.grid[0 + 2 * g];
    double v = omega;
    for (j = 0; j < nclu; j++) v += lgamma(NJ(tt, j) - sigma_temp) - lgamma(1 - sigma_temp);
    /* log(Vwm(num_treat-1, nclu-1, g)) = compact row r=1 */
    v += c->pb->shared.logV[(((size_t)tt * 2 + 1) * c->ntmax + (nclu - 1)) * K + g];
    os[g] = v;
  }
  double maxwei = os[0], denwei = 0.0;
  for (g = 1; g < K; g++)
    if (maxwei < os[g]) maxwei = os[g];
  for (g = 0; g < K; g++) {
    os[g] = exp(os[g] - maxwei);
    denwei += os[g];
  }
  double u0, u1, cw = 0.0;
  o_rng_u2(&c->rng, (uint32_t)l, TPPMX_SITE_SIGKAP, (uint32_t)tt, 0, 0, &u0, &u1);
  int pick = K - 1;
  for (g = 0; g < K; g++) {
    cw += os[g] / denwei;
    if (u0 < cw) {
      pick = g;
      break;
    }
  }
  c->s->idxs[0] = pick; /* quirk 2: one shared index, the last arm's draw survives */
  c->s->sigma[tt] = c->pb->shared.grid[0 + 2 * pick];
  c->s->kappa[tt] = c->pb->shared.grid[1 + 2 * pick];
  return 0;
}

/* ---- beta_update, utils_ct.cpp:710-800, for category kk ------------------------------------------ */
static void o_beta_update(o_ctx *c, int kk, int l, int32_t *beta_acc) {
  const int Q = c->Q, nobs = c->nobs, dim = c->D;
  const double mhtune = c->pb->model.mhtune_beta, mu_beta = 0.0;
  int i, q, h;
  for (q = 0; q < Q; q++) {
    h = kk + q * dim;
    const double sb = c->s->sigma_beta[h];
    double log_den = 0.0, log_num = 0.0;
    for (i = 0; i < nobs; i++) log_den = log_den - lgamma(exp(LG(i, kk))) + exp(LG(i, kk)) * log(JJm(i, kk));
    log_den += oracle_dnorm_log(c->s->beta[h], mu_beta, sb);
    double n0, n1;
    o_rng_n2(&c->rng, (uint32_t)l, TPPMX_SITE_BETA, 0, (uint32_t)(kk * Q + q), 0, &n0, &n1);
    const double beta_p = c->s->beta[h] + mhtune * n0;
    for (i = 0; i < nobs; i++) {
      const double lgp = LG(i, kk) - c->s->beta[h] * Zm(i, q) + beta_p * Zm(i, q);
      log_num = log_num - lgamma(exp(lgp)) + exp(lgp) * log(JJm(i, kk));
    }
    log_num += oracle_dnorm_log(beta_p, mu_beta, sb);
    const double ln_acp = log_num - log_den;
    double u0, u1;
    o_rng_u2(&c->rng, (uint32_t)l, TPPMX_SITE_BETA, 0, (uint32_t)(kk * Q + q), 1, &u0, &u1);
    if (log(u0) < ln_acp) {
      if (beta_acc) beta_acc[q + Q * kk] += 1;
      for (i = 0; i < nobs; i++) LG(i, kk) = LG(i, kk) - c->s->beta[h] * Zm(i, q) + beta_p * Zm(i, q);
      c->s->beta[h] = beta_p;
    }
  }
}

/* ---- horseshoe, utils_ct.cpp:876-915 --------------------------------------------------------------- */
static void o_horseshoe(o_ctx *c, int l) {
  const int len = c->Q * c->D;
  double *lambda = c->s->lambda, *beta = c->s->beta;
  double tau = c->s->tau[0];
  int k;
  for (k = 0; k < len; k++) { /* up_lambda_hs */
    double u0, u1;
    o_rng_u2(&c->rng, (uint32_t)l, TPPMX_SITE_HS_LAMBDA, 0, (uint32_t)k, 0, &u0, &u1);
    if (O_TAPE_TOP()) o_tape_rec(1, u1, 0.0); /* this site consumes both uniforms of the block */
    double eta = pow(lambda[k], -2.0);
    const double upsi = u0 * pow((1 + eta), -1.0); /* runif(0, 1/(1+eta)) */
    const double tempps = pow(beta[k], 2.0) / (2 * pow(tau, 2.0));
    const double ub = (1 - upsi) / upsi;
    double Fub = 1.0 - exp(-tempps * ub);
    if (Fub < pow(10.0, -4)) Fub = pow(10.0, -4);
    const double up = u1 * Fub; /* runif(0, Fub) */
    eta = -log(1.0 - up) / tempps;
    lambda[k] = pow(eta, -0.5);
  }
  { /* up_tau_hs */
    double tempt = 0.0, u0, u1;
    for (k = 0; k < len; k++) tempt += pow((beta[k] / lambda[k]), 2.0);
    tempt = tempt / 2;
    double et = pow(tau, -2.0);
    o_rng_u2(&c->rng, (uint32_t)l, TPPMX_SITE_HS_TAU, 0, 0, 0, &u0, &u1);
    if (O_TAPE_TOP()) o_tape_rec(1, u1, 0.0); /* both uniforms consumed */
    const double utau = u0 * pow((1.0 + et), -1.0);
    const double ubt = (1.0 - utau) / utau;
    const double shape = (((double)len) + 1.0) / 2;
    double Fubt = tppmx_gammainc(shape, ubt * tempt); /* pgamma(ubt, shape, scale = 1/tempt) */
    if (Fubt < pow(10.0, -4)) Fubt = pow(10.0, -4);
    const double ut = u1 * Fubt;
    et = tppmx_gammaincinv(shape, ut) / tempt; /* qgamma(ut, shape, scale = 1/tempt) */
    if (et < .0001) et = .0001;
    tau = pow(et, -0.5);
    if (tau < .0001) tau = .0001;
    c->s->tau[0] = tau;
  }
  for (k = 0; k < len; k++) c->s->sigma_beta[k] = lambda[k] * tau;
}

/* ---- JJ, ss, src/dm_ppmx_ct.cpp:1042-1055 ----------------------------------------------------------- */
static void o_jj_ss(o_ctx *c, int l) {
  const int nobs = c->nobs, dim = c->D;
  int i, k;
  for (i = 0; i < nobs; i++) {
    double TT = 0.0;
    for (k = 0; k < dim; k++) {
      const double shape = Ym(i, k) + exp(LG(i, k));
      double v = o_rng_gamma(&c->rng, (uint32_t)l, TPPMX_SITE_JJ, 0, (uint32_t)(i * dim + k), 0, shape) *
                 pow(c->s->ss[i] + 1.0, -1.0);
      if (v < pow(10.0, -100.0)) v = pow(10.0, -100.0);
      JJm(i, k) = v;
      TT += v;
    }
    (void)TT;
  }
  for (i = 0; i < nobs; i++) { /* :1053-1055, ypiu(i) = 1 (:66) */
    double TT = 0.0;
    for (k = 0; k < dim; k++) TT += JJm(i, k);
    c->s->ss[i] = o_rng_gamma(&c->rng, (uint32_t)l, TPPMX_SITE_SS, 0, (uint32_t)i, 0, 1.0) * pow(TT, -1.0);
  }
}

/* Everything of iteration l after the sweep: src/dm_ppmx_ct.cpp:866-1055.
 * eta_acc (T x ntmax col-major) / beta_acc (Q x D col-major) accumulate acceptances, may be NULL. */
int oracle_update_iter(const oracle_problem *pb, tppmx_state *s, uint64_t seed, uint32_t chain_id,
                       int32_t l, int32_t *eta_acc, int32_t *beta_acc) {
  o_ctx cc, *c = &cc;
  if (o_ctx_open(c, pb, s, seed, chain_id)) return 1;
  const tppmx_model *m = &pb->model;
  int tt, j, k, rc = 0;
  for (tt = 0; tt < c->T; tt++)
    for (j = 0; j < s->nclu[tt]; j++) o_eta_update(c, tt, j, l, eta_acc);
  if (m->upd_hier == 1)
    for (tt = 0; tt < c->T; tt++) o_hierarchy(c, tt, l);
  if (m->cohesion == 2) {
    double *os = (double *)malloc(sizeof(double) * c->G);
    for (tt = 0; tt < c->T && !rc; tt++) rc = o_sigma_kappa(c, tt, l, os);
    free(os);
  }
  if (m->noprog != 1 && !rc) {
    for (k = 0; k < c->D; k++) o_beta_update(c, k, l, beta_acc);
    if (m->hsp == 1) o_horseshoe(c, l);
  }
  if (!rc) o_jj_ss(c, l);
  o_ctx_close(c);
  return rc;
}

/* One full iteration: sweep (src/dm_ppmx_ct.cpp:349-860) then the updates above. */
int oracle_iterate(const oracle_problem *pb, tppmx_state *s, uint64_t seed, uint32_t chain_id,
                   int32_t iter0, int32_t n_iter, int32_t *eta_acc, int32_t *beta_acc) {
  for (int l = iter0; l < iter0 + n_iter; l++) {
    int rc = oracle_sweep(pb, s, seed, chain_id, l, 1);
    if (!rc) rc = oracle_update_iter(pb, s, seed, chain_id, l, eta_acc, beta_acc);
    if (rc) return rc;
  }
  return 0;
}

/* probes for the known-answer tests */
double oracle_gammainc(double a, double x) { return tppmx_gammainc(a, x); }
double oracle_gammaincinv(double a, double p) { return tppmx_gammaincinv(a, p); }
double oracle_dmvnrm_log(const double *x, const double *mean, const double *sigma, int n) {
  return o_dmvnrm_log(x, mean, sigma, n);
}
void oracle_wishart(uint64_t seed, uint32_t chain, int32_t l, const double *S, double df, int n, double *W) {
  o_rng g = {seed, chain};
  o_wishart(&g, l, 0, S, df, n, W);
}

/* utils_ct.cpp:842-874 */
static double o_dmultinom(const double *x_in, int size, const double *prob_in, int K, int Log) {
  double prob[TPPMX_MAXD], x[TPPMX_MAXD];
  double s = 0.0;
  for (int i = 0; i < K; i++) s += prob_in[i];
  for (int i = 0; i < K; i++) prob[i] = prob_in[i] / s;
  for (int i = 0; i < K; i++) x[i] = floor(x_in[i] + .5);
  double r = lgamma(size + 1);
  for (int i = 0; i < K; i++) {
    /* the reference evaluates x*log(prob) for every i; with x == 0 the term is 0 as long as
     * prob > 0, which JJ >= 1e-100 guarantees */
    if (x[i] != 0.0) r += x[i] * log(prob[i]) - lgamma(x[i] + 1);
  }
  return Log ? r : exp(r);
}

/* in-sample fit of a saved iteration, src/dm_ppmx_ct.cpp:1060-1075.
 * pi, ispred: nobs x D col-major; like: [nobs].  rmultinom(1, pii) is one inverse-CDF uniform. */
int oracle_insample(const oracle_problem *pb, const tppmx_state *s, uint64_t seed, uint32_t chain_id,
                    int32_t l, double *pi, double *ispred, double *like) {
  const int nobs = pb->dims.nobs, dim = pb->dims.D;
  o_rng g = {seed, chain_id};
  for (int i = 0; i < nobs; i++) {
    double TT = 0.0, pii[TPPMX_MAXD], yrow[TPPMX_MAXD];
    for (int k = 0; k < dim; k++) TT += s->JJ[i + nobs * k];
    for (int k = 0; k < dim; k++) {
      pii[k] = s->JJ[i + nobs * k] / TT;
      yrow[k] = pb->data.y[i + nobs * k];
      if (pi) pi[i + nobs * k] = pii[k];
    }
    double u0, u1, cw = 0.0;
    o_rng_u2(&g, (uint32_t)l, TPPMX_SITE_ISPRED, 0, (uint32_t)i, 0, &u0, &u1);
    int pick = dim - 1;
    for (int k = 0; k < dim; k++) {
      cw += pii[k];
      if (u0 < cw) { pick = k; break; }
    }
    if (ispred)
      for (int k = 0; k < dim; k++) ispred[i + nobs * k] = (k == pick) ? 1.0 : 0.0;
    if (like) like[i] = o_dmultinom(yrow, 1, pii, dim, 0);
  }
  return 0;
}
