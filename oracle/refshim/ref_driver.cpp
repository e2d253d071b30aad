/*
 * oracle/refshim/ref_driver.cpp — TEST INFRASTRUCTURE ONLY.
 *
 * extern "C" access to the reference's OWN functions, compiled unmodified from
 * /root/reference/src/{utils_ct,dm_ppmx_ct,prior_ppmx}.cpp against the stand-in headers of this
 * directory (oracle/Makefile target `ref` -> oracle/_ref/libref.so).  Used by
 * tests/test_oracle_vs_reference.py and scripts/make_reference_golden.py to pin the oracle
 * against the reference's statements executed here.
 *
 * Nothing of the reference is restated in this file: it only marshals flat arrays into the
 * argument types the reference declares (src/utils_ct.h, src/dm_ppmx_ct.cpp:10-23,
 * src/prior_ppmx.cpp:9-16) and copies the returned list out.
 */
#include "RcppArmadillo.h"

#include <cstdint>
#include <cstdio>


/* ---- the reference's declarations (src/utils_ct.h; the two samplers have no header) ----------
 * REFDRV_SHIM_ONLY (oracle/Makefile target `shim`): the same driver over the CUDA library's Rcpp shim
 * (treatppmx_b200/csrc/rcpp_shim/treatppmx_shim.cpp), which defines the two samplers with the
 * reference's signatures; the helper-by-helper entry points below need the reference and are left out. */
#ifndef REFDRV_SHIM_ONLY
#include "utils_ct.h"
arma::vec dmvnrm(arma::mat x, arma::rowvec mean, arma::mat sigma, bool logd);
#endif
Rcpp::List dm_ppmx_ct(int iter, int burn, int thin, int nobs, arma::vec treatments, int PPMx, int ncon,
                      int ncat, arma::vec catvec, double alpha, arma::mat grid, arma::cube Vwm,
                      int cohesion, int CC, int reuse, int consim, int similarity, int calibration,
                      int coardegree, arma::mat y, arma::mat z, arma::mat zpred, int noprog,
                      arma::vec xcon, arma::vec xcat, arma::vec xconp, arma::vec xcatp, int npred,
                      arma::vec similparam, arma::vec hP0_mu0, double hP0_nu0, double hP0_s0,
                      double hP0_Lambda0, int upd_hier, arma::vec initbeta, int hsp,
                      arma::vec mhtunepar, int A, arma::vec n_a, arma::mat curr_cluster,
                      arma::mat card_cluster, arma::vec ncluster_curr);
Rcpp::List prior_ppmx_core(int iter, int burn, int thin, int nobs, int PPMx, int ncon, int ncat,
                           double alpha, double sigma, arma::mat Vwm, int cohesion, int CC, int consim,
                           int similarity, int calibration, int coardegree, arma::vec xcon,
                           arma::vec similparam, arma::vec curr_cluster, arma::vec card_cluster,
                           int ncluster_curr);

/* ---- tape (per thread: bench.py times one chain per host thread) ------------------------------- */
namespace {
thread_local const double *g_kind = nullptr, *g_val = nullptr, *g_aux = nullptr;
thread_local long g_n = 0, g_pos = 0;
thread_local std::string g_err;
/* With no tape set (ref_rng_free) the variates come from a free-running generator instead: the
 * timing legs of bench.py need R's RNG replaced by SOMETHING, not by particular values.
 * splitmix64 uniforms, Box-Muller normals, Marsaglia-Tsang gammas. */
thread_local uint64_t g_rng = 0x9E3779B97F4A7C15ull;
double fr_unif() {
  uint64_t z = (g_rng += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return ((double)(z >> 11) + 0.5) * 0x1.0p-53;
}
double fr_norm() { return std::sqrt(-2.0 * std::log(fr_unif())) * std::cos(6.283185307179586 * fr_unif()); }
double fr_gamma(double a) {
  double boost = 1.0;
  if (a < 1.0) {
    boost = std::exp(std::log(fr_unif()) / a);
    a += 1.0;
  }
  const double d = a - 1.0 / 3.0, c = 1.0 / std::sqrt(9.0 * d);
  for (;;) {
    const double x = fr_norm();
    double v = 1.0 + c * x;
    if (v <= 0.0) continue;
    v = v * v * v;
    const double u = fr_unif(), x2 = x * x;
    if (u < 1.0 - 0.0331 * x2 * x2 || std::log(u) < 0.5 * x2 + d * (1.0 - v + std::log(v))) return boost * d * v;
  }
}
}  // namespace

namespace refshim {
double tape_pop(int kind, double aux) {
  if (!g_kind) return kind == TAPE_U ? fr_unif() : kind == TAPE_N ? fr_norm() : fr_gamma(aux);
  char buf[256];
  if (g_pos >= g_n) {
    snprintf(buf, sizeof buf, "tape exhausted at draw %ld (wanted kind %d)", g_pos, kind);
    throw std::runtime_error(buf);
  }
  if ((int)g_kind[g_pos] != kind) {
    snprintf(buf, sizeof buf, "tape draw %ld: the reference asks for kind %d, the oracle recorded kind %d",
             g_pos, kind, (int)g_kind[g_pos]);
    throw std::runtime_error(buf);
  }
  if (kind == TAPE_G && !(std::fabs(g_aux[g_pos] - aux) <= 1e-9 * std::fabs(aux))) {
    snprintf(buf, sizeof buf, "tape draw %ld: gamma shape %.17g asked, %.17g recorded", g_pos, aux, g_aux[g_pos]);
    throw std::runtime_error(buf);
  }
  return g_val[g_pos++];
}
/* R::pgamma / R::qgamma stand-ins for the REFERENCE build.  Deliberately independent of the product's
 * csrc/incgamma.h (closed forms / continued fraction / Newton from Wilson-Hilferty, shared by the device
 * kernel and the oracle): here P(a, x) is Kummer's all-positive series
 *     P(a, x) = x^a e^-x sum_{k>=0} x^k / Gamma(a + k + 1)
 * summed in extended precision for every x (no cancellation, so it is accurate relative to P everywhere the
 * path goes), and the inverse is plain bisection on it to the last bit.  The whole-run pin of the oracle
 * against the reference (tests/test_oracle_vs_reference.py) therefore checks the product's incomplete-gamma
 * pair against a second implementation, not against itself. */
static long double gammainc_series_ld(long double a, long double x) {
  if (!(x > 0.0L)) return 0.0L;
  long double term = 1.0L / a, sum = term;
  for (int k = 1; k < 100000; k++) {
    term *= x / (a + (long double)k);
    sum += term;
    if (term < sum * 1e-21L) break;
  }
  const long double lp = a * logl(x) - x - lgammal(a) + logl(sum);
  const long double p = expl(lp);
  return p > 1.0L ? 1.0L : p;
}
double gammainc_lower(double a, double x) { return (double)gammainc_series_ld((long double)a, (long double)x); }
double gammainc_lower_inv(double a, double p) {
  if (!(p > 0.0)) return 0.0;
  if (p >= 1.0) return INFINITY;
  long double lo = 0.0L, hi = 1.0L;
  while (gammainc_series_ld((long double)a, hi) < (long double)p) hi *= 2.0L;
  for (int it = 0; it < 200; it++) {
    const long double mid = 0.5L * (lo + hi);
    if (mid <= lo || mid >= hi) break;
    if (gammainc_series_ld((long double)a, mid) < (long double)p) lo = mid; else hi = mid;
  }
  return (double)(0.5L * (lo + hi));
}
}  // namespace refshim

static arma::vec mk_vec(const double *p, long n) {
  arma::vec v((arma::uword)n);
  for (long i = 0; i < n; i++) v.d[i] = p[i];
  return v;
}
static arma::mat mk_mat(const double *p, long r, long c) {
  arma::mat m((arma::uword)r, (arma::uword)c);
  for (long i = 0; i < r * c; i++) m.d[i] = p[i];
  return m;
}
static void put(double *dst, const arma::mat &m) {
  if (dst) std::copy(m.d.begin(), m.d.end(), dst);
}
static void put(double *dst, const arma::cube &c) {
  if (!dst) return;
  for (arma::uword k = 0; k < c.n_slices; k++)
    std::copy(c.s[k].d.begin(), c.s[k].d.end(), dst + k * c.n_rows * c.n_cols);
}

extern "C" {

/* kind[i] in {1 uniform, 2 normal, 3 gamma(shape aux[i], scale 1)}, val[i] the variate */
void ref_tape_set(const double *kind, const double *val, const double *aux, long n) {
  g_kind = kind; g_val = val; g_aux = aux; g_n = n; g_pos = 0;
}
/* no tape: free-running stand-in generator, seeded per thread (timing only) */
void ref_rng_free(uint64_t seed) {
  g_kind = g_val = g_aux = nullptr;
  g_n = g_pos = 0;
  g_rng = seed * 0x9E3779B97F4A7C15ull + 1;
}
long ref_tape_pos(void) { return g_pos; }
const char *ref_last_error(void) { return g_err.c_str(); }

/* ---- the stand-in containers themselves (checked against numpy in tests/test_oracle_vs_reference.py):
 * inverse, upper Cholesky factor, covariance (rows = observations, divided by N), row means, a
 * matrix product through sub-views, for an n x n SPD matrix A and an m x n data matrix X */
int ref_shim_linalg(const double *A, int n, const double *X, int m, double *inv_out, double *chol_out, double *cov_out,
                    double *mean_out, double *prod_out) {
  try {
    arma::mat a = mk_mat(A, n, n), x = mk_mat(X, m, n);
    put(inv_out, arma::inv(a));
    put(chol_out, arma::chol(a));
    put(cov_out, arma::cov(x, 1));
    put(mean_out, arma::mean(x.t(), 1));
    arma::mat p(n, n);
    for (int j = 0; j < n; j++) p.col(j) = a * a.col(j);   /* sub-view in, sub-view out */
    p(arma::span::all, arma::span(0, 0)) = p.col(0) * 1.0;
    put(prod_out, p);
    return 0;
  } catch (const std::exception &e) {
    g_err = e.what();
    return 1;
  }
}

#ifndef REFDRV_SHIM_ONLY
/* ---- src/utils_ct.cpp, function by function ---------------------------------------------------- */
double ref_dinvgamma(double y, double a, double b, int lg) { return dinvgamma(y, a, b, lg); }
double ref_dN_IG(double mu, double s2, double mu0, double k0, double a0, double b0, int lg) {
  return dN_IG(mu, s2, mu0, k0, a0, b0, lg);
}
double ref_gsimconNN(double m0, double v2, double s20, double sumx, double sumx2, int n, int DD, int lg) {
  return gsimconNN(m0, v2, s20, sumx, sumx2, n, DD, lg);
}
double ref_gsimconNNIG(double m0, double k0, double nu0, double s20, double sumx, double sumx2, int n, int DD,
                       int lg) {
  return gsimconNNIG(m0, k0, nu0, s20, sumx, sumx2, n, DD, lg);
}
double ref_gsimcatDM(const double *nobsj, const double *dirw, int C, int DD, int lg) {
  return gsimcatDM(mk_vec(nobsj, C), mk_vec(dirw, C), C, DD, lg);
}
double ref_calculate_gamma(const double *eta, int D, int ncols, const double *Z, int nobs, int Q,
                           const double *beta, int clu, int k, int i, int Log) {
  return calculate_gamma(mk_mat(eta, D, ncols), mk_mat(Z, nobs, Q), mk_vec(beta, (long)Q * D), clu, k, i, Log);
}
double ref_dmvnrm_log(const double *x, const double *mean, const double *sigma, int n) {
  return dmvnrm(mk_mat(x, 1, n), arma::rowvec(mk_mat(mean, 1, n)), mk_mat(sigma, n, n), true)(0);
}
#endif /* REFDRV_SHIM_ONLY */
/* ---- src/dm_ppmx_ct.cpp:10-1471, the whole run -------------------------------------------------
 * Inputs are the 42 arguments in the reference's order, matrices column-major (Armadillo).  Vwm is
 * the dense cube (vr x vc x G).  Outputs are the 18 list elements, flattened column-major; fields
 * are laid out element after element: eta [D x ntmax] for (ll, tt) with ll fastest (Armadillo field
 * order), pipred / ypred [npred x D x T] per saved iteration. */
int ref_dm_ppmx_ct(int iter, int burn, int thin, int nobs, const double *treatments, int PPMx, int ncon, int ncat,
                   const double *catvec, int ncatvec, double alpha, const double *grid, int G, const double *Vwm,
                   int vr, int vc, int cohesion, int CC, int reuse, int consim, int similarity, int calibration,
                   int coardegree, const double *y, int D, const double *z, int Q, const double *zpred, int noprog,
                   const double *xcon, const double *xcat, const double *xconp, const double *xcatp, int npred,
                   const double *similparam, const double *hP0_mu0, double hP0_nu0, double hP0_s0,
                   double hP0_Lambda0, int upd_hier, const double *initbeta, int hsp, const double *mhtunepar, int A,
                   const double *n_a, const double *curr_cluster, const double *card_cluster,
                   const double *ncluster_curr, int ntmax,
                   /* outputs */
                   double *eta, double *beta, double *sigma_ngg, double *kappa_ngg, double *eta_acc, double *beta_acc,
                   double *cl_lab, double *num_treat, double *pi, double *pipred, double *yispred, double *ypred,
                   double *clupred, double *nclus, double *WAIC, double *lpml, double *theta_prior,
                   double *sigma_prior) {
  try {
    arma::cube V((arma::uword)vr, (arma::uword)vc, (arma::uword)(Vwm ? G : 0));
    if (Vwm)
      for (int g = 0; g < G; g++)
        for (long i = 0; i < (long)vr * vc; i++) V.s[g].d[i] = Vwm[(long)g * vr * vc + i];
    Rcpp::List o = dm_ppmx_ct(
        iter, burn, thin, nobs, mk_vec(treatments, nobs), PPMx, ncon, ncat, mk_vec(catvec, ncatvec), alpha,
        mk_mat(grid, 2, G), V, cohesion, CC, reuse, consim, similarity, calibration, coardegree, mk_mat(y, nobs, D),
        mk_mat(z, nobs, Q), mk_mat(zpred, npred, Q), noprog, mk_vec(xcon, (long)nobs * ncon),
        mk_vec(xcat, (long)nobs * ncat), mk_vec(xconp, (long)npred * ncon), mk_vec(xcatp, (long)npred * ncat), npred,
        mk_vec(similparam, 7), mk_vec(hP0_mu0, D), hP0_nu0, hP0_s0, hP0_Lambda0, upd_hier,
        mk_vec(initbeta, (long)Q * D), hsp, mk_vec(mhtunepar, 2), A, mk_vec(n_a, A), mk_mat(curr_cluster, A, ntmax),
        mk_mat(card_cluster, A, ntmax), mk_vec(ncluster_curr, A));
    const int nout = (iter - burn) / thin;
    const arma::field<arma::mat> &fe = o["eta"].fm;
    if (eta)
      for (arma::uword e = 0; e < fe.n_elem; e++) put(eta + e * (long)D * ntmax, fe.e[e]);
    put(beta, o["beta"].m);
    put(sigma_ngg, o["sigma_ngg"].m);
    put(kappa_ngg, o["kappa_ngg"].m);
    put(eta_acc, o["eta_acc"].m);
    put(beta_acc, o["beta_acc"].m);
    put(cl_lab, o["cl_lab"].c);
    put(num_treat, o["num_treat"].m);
    put(pi, o["pi"].c);
    put(yispred, o["yispred"].c);
    for (int ll = 0; ll < nout; ll++) {
      if (pipred) put(pipred + (long)ll * npred * D * A, o["pipred"].fc.e[ll]);
      if (ypred) put(ypred + (long)ll * npred * D * A, o["ypred"].fc.e[ll]);
    }
    put(clupred, o["clupred"].m);
    put(nclus, o["nclus"].m);
    if (WAIC) *WAIC = o["WAIC"].num;
    put(lpml, o["lpml"].m);
    put(theta_prior, o["theta_prior"].m);
    put(sigma_prior, o["sigma_prior"].c);
    return 0;
  } catch (const std::exception &e) {
    g_err = e.what();
    return 1;
  }
}

/* ---- src/prior_ppmx.cpp:9-466 -------------------------------------------------------------------
 * outputs: nclu[nout], njout[nobs x nout] (col-major) */
int ref_prior_ppmx_core(int iter, int burn, int thin, int nobs, int PPMx, int ncon, int ncat, double alpha,
                        double sigma, const double *Vwm, int vr, int vc, int cohesion, int CC, int consim,
                        int similarity, int calibration, int coardegree, const double *xcon,
                        const double *similparam, const double *curr_cluster, const double *card_cluster,
                        int ncluster_curr, double *nclu, double *njout) {
  try {
    Rcpp::List o = prior_ppmx_core(iter, burn, thin, nobs, PPMx, ncon, ncat, alpha, sigma, mk_mat(Vwm, vr, vc),
                                   cohesion, CC, consim, similarity, calibration, coardegree,
                                   mk_vec(xcon, (long)nobs * ncon), mk_vec(similparam, 7), mk_vec(curr_cluster, nobs),
                                   mk_vec(card_cluster, nobs), ncluster_curr);
    put(nclu, o["nclu"].m);
    put(njout, o["njout"].m);
    return 0;
  } catch (const std::exception &e) {
    g_err = e.what();
    return 1;
  }
}

} /* extern "C" */
