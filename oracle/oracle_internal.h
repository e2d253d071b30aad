/*
 * oracle_internal.h — shared internals of the CPU oracle (TEST INFRASTRUCTURE ONLY):
 * the per-call context and the Armadillo-layout accessors used by ppmx_oracle.c (sweep,
 * predictive) and ppmx_chain.c (the remaining per-iteration updates).
 */
#ifndef ORACLE_INTERNAL_H
#define ORACLE_INTERNAL_H
#include "philox_ref.h"
#include "ppmx_oracle.h"

typedef struct o_ctx {
  const oracle_problem *pb;
  tppmx_state *s;
  int nobs, T, D, Q, P, ncat, maxC, npred, G, ntmax, CC;
  int num_treat[TPPMX_MAXT];
  double *dirweights; /* [maxC] */
  double *njctmp;     /* [maxC] */
  double *weight, *pweight, *gtilN, *gtilY, *lgtilN, *lgtilY; /* [ntmax+CC] per call arm */
  o_rng rng;
} o_ctx;

#define LAB(t, it) (c->s->labels[(t) + c->T * (it)])
#define NJ(t, j) (c->s->nj[(t) + c->T * (j)])
#define SUMX(t, idx) (c->s->sumx[(t) + c->T * (idx)])
#define SUMX2(t, idx) (c->s->sumx2[(t) + c->T * (idx)])
#define NJC(t, idx) (c->s->njc[(t) + c->T * (idx)])
#define ETAS(k, j, t) (c->s->eta_star[(k) + c->D * ((j) + c->ntmax * (t))])
#define ETAE(k, m, t) (c->s->eta_empty[(k) + c->D * ((m) + c->CC * (t))])
#define JJm(i, k) (c->s->JJ[(i) + c->nobs * (k)])
#define LG(i, k) (c->s->loggamma[(i) + c->nobs * (k)])
#define Ym(i, k) (c->pb->data.y[(i) + c->nobs * (k)])
#define Zm(i, q) (c->pb->data.z[(i) + c->nobs * (q)])
#define XCON(idx) (c->pb->data.xcon[(idx)])
#define XCAT(idx) (c->pb->data.xcat[(idx)])


int o_ctx_open(o_ctx *c, const oracle_problem *pb, tppmx_state *s, uint64_t seed, uint32_t chain_id);
void o_ctx_close(o_ctx *c);
double o_gcon(const o_ctx *c, double sx, double sx2, int n);
void o_chol_lower(const double *A, int n, double *L);
void o_sweep_iter(o_ctx *c, int l);

#endif
