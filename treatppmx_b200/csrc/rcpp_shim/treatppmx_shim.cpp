// treatppmx_shim.cpp — the Rcpp side of the drop-in: the package's two sampler entry points with
// the reference's OWN C++ signatures, forwarding to the CUDA library's C ABI (include/tppmx.h).
//
//   dm_ppmx_ct(42 arguments)       reference src/dm_ppmx_ct.cpp:10-23 (body :24-1472), .Call entry
//                                  _treatppmx_dm_ppmx_ct, src/RcppExports.cpp:16-65, :187
//   prior_ppmx_core(21 arguments)  reference src/prior_ppmx.cpp:9-16 (body :17-466), .Call entry
//                                  _treatppmx_prior_ppmx_core, src/RcppExports.cpp:189
//
// A maintainer drops this file into the package's src/ IN PLACE OF the two function bodies: the
// generated glue (src/RcppExports.cpp, R/RcppExports.R), the R wrappers ppmxct() / prior_ppmx() and
// the returned lists do not change.  src/Makevars gains -I<include> and -ltppmx.
// set.seed() keeps governing a run: the Philox key is drawn from R's RNG (two unif_rand() draws
// under the RNGScope of the generated glue, src/RcppExports.cpp:19).
//
// No R / Rcpp / RcppArmadillo exists in this repository's build image, so the file is also
// compiled, with -DTPPMX_SHIM_STANDIN, against the stand-in headers of oracle/refshim/ (the ones the
// reference's own sources are compiled against for the oracle pin) into oracle/_ref/libshim.so and
// driven through the reference's signatures on the GPU (tests/test_gpu_shim.py).  The stand-in
// containers have no contiguous cubes and no pointer constructors; the few lines that differ are
// behind that macro, everything else is the code a real build compiles.
#include <RcppArmadillo.h>

#include <cstdint>
#include <vector>

#include "tppmx.h"

#if !defined(TPPMX_SHIM_STANDIN) && defined(TPPMX_SHIM_R_INTERRUPT)
extern "C" int R_interrupts_pending;  // Rinterface.h (Unix): Ctrl-C sets it; the library polls it
#endif

namespace {

tppmx_ctx *g_ctx = nullptr;  // one device context per R session
uint64_t g_last_seed = 0;

tppmx_ctx *shim_ctx() {
  if (!g_ctx && tppmx_create(&g_ctx, 0) != TPPMX_OK) Rcpp::stop("tppmx: no usable CUDA device");
#if !defined(TPPMX_SHIM_STANDIN) && defined(TPPMX_SHIM_R_INTERRUPT)
  tppmx_set_interrupt_flag(g_ctx, (const volatile int32_t *)&R_interrupts_pending, 64);
#endif
  return g_ctx;
}

uint64_t seed_from_r_rng() {
  const uint64_t hi = (uint64_t)(R::unif_rand() * 4294967296.0), lo = (uint64_t)(R::unif_rand() * 4294967296.0);
  return g_last_seed = (hi << 32) | (lo & 0xffffffffull);
}

#ifdef TPPMX_SHIM_STANDIN
std::vector<double> g_cube_tmp;
const double *cube_ptr(const arma::cube &c) {  // stand-in cubes are lists of slices: flatten
  g_cube_tmp.resize((size_t)c.n_rows * c.n_cols * c.n_slices);
  for (arma::uword k = 0; k < c.n_slices; k++)
    for (arma::uword e = 0; e < c.n_rows * c.n_cols; e++)
      g_cube_tmp[(size_t)k * c.n_rows * c.n_cols + e] = c.slice(k).memptr()[e];
  return c.n_slices ? g_cube_tmp.data() : nullptr;
}
arma::mat mat_of(const double *p, arma::uword r, arma::uword c) {
  arma::mat m(r, c);
  for (arma::uword e = 0; e < r * c; e++) m.memptr()[e] = p[e];
  return m;
}
arma::vec vec_of(const double *p, arma::uword n) {
  arma::vec v(n);
  for (arma::uword e = 0; e < n; e++) v.memptr()[e] = p[e];
  return v;
}
arma::cube cube_of(const double *p, arma::uword r, arma::uword c, arma::uword n) {
  arma::cube q(r, c, n);
  for (arma::uword k = 0; k < n; k++)
    for (arma::uword e = 0; e < r * c; e++) q.slice(k).memptr()[e] = p[(size_t)k * r * c + e];
  return q;
}
#else
const double *cube_ptr(const arma::cube &c) { return c.n_elem ? c.memptr() : nullptr; }
arma::mat mat_of(const double *p, arma::uword r, arma::uword c) { return arma::mat(p, r, c); }
arma::vec vec_of(const double *p, arma::uword n) { return arma::vec(p, n); }
arma::cube cube_of(const double *p, arma::uword r, arma::uword c, arma::uword n) { return arma::cube(p, r, c, n); }
#endif

}  // namespace

extern "C" uint64_t tppmx_shim_last_seed(void) { return g_last_seed; }  // test hook

// [[Rcpp::export]]
Rcpp::List dm_ppmx_ct(int iter, int burn, int thin, int nobs, arma::vec treatments, int PPMx, int ncon, int ncat,
                      arma::vec catvec, double alpha, arma::mat grid, arma::cube Vwm, int cohesion, int CC, int reuse,
                      int consim, int similarity, int calibration, int coardegree, arma::mat y, arma::mat z,
                      arma::mat zpred, int noprog, arma::vec xcon, arma::vec xcat, arma::vec xconp, arma::vec xcatp,
                      int npred, arma::vec similparam, arma::vec hP0_mu0, double hP0_nu0, double hP0_s0,
                      double hP0_Lambda0, int upd_hier, arma::vec initbeta, int hsp, arma::vec mhtunepar, int A,
                      arma::vec n_a, arma::mat curr_cluster, arma::mat card_cluster, arma::vec ncluster_curr) {
  tppmx_ctx *ctx = shim_ctx();
  const int dim = (int)y.n_cols, Q = (int)z.n_cols, ntmax = (int)curr_cluster.n_cols;
  const int nout = thin > 0 ? (iter - burn) / thin : 0, no = nout > 0 ? nout : 0;
  tppmx_ppmxct_args in = {};
  in.iter = iter; in.burn = burn; in.thin = thin; in.nobs = nobs; in.treatments = treatments.memptr();
  in.PPMx = PPMx; in.ncon = ncon; in.ncat = ncat; in.catvec = catvec.memptr(); in.alpha = alpha;
  in.grid = grid.memptr(); in.G = (int)grid.n_cols;
  in.Vwm = cube_ptr(Vwm); in.Vwm_rows = (int)Vwm.n_rows; in.Vwm_cols = (int)Vwm.n_cols;
  in.cohesion = cohesion; in.CC = CC; in.reuse = reuse; in.consim = consim; in.similarity = similarity;
  in.calibration = calibration; in.coardegree = coardegree;
  in.y = y.memptr(); in.D = dim; in.z = z.memptr(); in.Q = Q; in.zpred = zpred.memptr(); in.noprog = noprog;
  in.xcon = xcon.memptr(); in.xcat = xcat.memptr(); in.xconp = xconp.memptr(); in.xcatp = xcatp.memptr();
  in.npred = npred; in.similparam = similparam.memptr(); in.hP0_mu0 = hP0_mu0.memptr();
  in.hP0_nu0 = hP0_nu0; in.hP0_s0 = hP0_s0; in.hP0_Lambda0 = hP0_Lambda0; in.upd_hier = upd_hier;
  in.initbeta = initbeta.memptr(); in.hsp = hsp; in.mhtunepar = mhtunepar.memptr(); in.A = A; in.n_a = n_a.memptr();
  in.curr_cluster = curr_cluster.memptr(); in.card_cluster = card_cluster.memptr(); in.ntmax = ntmax;
  in.ncluster_curr = ncluster_curr.memptr();
  in.seed = seed_from_r_rng();

  // flat, column-major stores for the 18 outputs (the layouts of the reference's containers, :296-335)
  const size_t n1 = (size_t)npred * dim * A;
  std::vector<double> eta((size_t)no * A * dim * ntmax + 1), beta((size_t)Q * dim * no + 1), sig((size_t)A * no + 1),
      kap((size_t)A * no + 1), eta_acc((size_t)A * ntmax), beta_acc((size_t)(Q > 0 ? Q : 1) * dim),
      cl_lab((size_t)A * ntmax * no + 1), num_treat(A), pi((size_t)nobs * dim * no + 1), pipred(n1 * no + 1),
      yispred((size_t)nobs * dim * no + 1), ypred(n1 * no + 1), clupred((size_t)no * npred * A + 1),
      nclus((size_t)A * no + 1), lpml(no + 1), Theta((size_t)dim * A), Sigma((size_t)dim * dim * A);
  double WAIC = 0;
  tppmx_ppmxct_out out = {eta.data(), beta.data(), sig.data(), kap.data(), eta_acc.data(), beta_acc.data(),
                          cl_lab.data(), num_treat.data(), pi.data(), pipred.data(), yispred.data(), ypred.data(),
                          clupred.data(), nclus.data(), &WAIC, lpml.data(), Theta.data(), Sigma.data()};
  if (tppmx_dm_ppmx_ct(ctx, &in, &out) != TPPMX_OK) Rcpp::stop(tppmx_last_error(ctx));

  // wrap the blocks in the reference's return types (src/dm_ppmx_ct.cpp:1453-1471); no arithmetic
  arma::field<arma::mat> eta_out(no, A);
  arma::field<arma::cube> pigreco_pred(no, 1), ppred_out(no, 1);
  for (int ll = 0; ll < no; ll++) {
    for (int t = 0; t < A; t++) eta_out(ll, t) = mat_of(eta.data() + ((size_t)ll + (size_t)no * t) * dim * ntmax, dim, ntmax);
    pigreco_pred(ll, 0) = cube_of(pipred.data() + (size_t)ll * n1, npred, dim, A);
    ppred_out(ll, 0) = cube_of(ypred.data() + (size_t)ll * n1, npred, dim, A);
  }
  return Rcpp::List::create(
      Rcpp::Named("eta") = eta_out, Rcpp::Named("beta") = mat_of(beta.data(), Q * dim, no),
      Rcpp::Named("sigma_ngg") = mat_of(sig.data(), A, no), Rcpp::Named("kappa_ngg") = mat_of(kap.data(), A, no),
      Rcpp::Named("eta_acc") = mat_of(eta_acc.data(), A, ntmax), Rcpp::Named("beta_acc") = mat_of(beta_acc.data(), Q, dim),
      Rcpp::Named("cl_lab") = cube_of(cl_lab.data(), A, ntmax, no), Rcpp::Named("num_treat") = vec_of(num_treat.data(), A),
      Rcpp::Named("pi") = cube_of(pi.data(), nobs, dim, no), Rcpp::Named("pipred") = pigreco_pred,
      Rcpp::Named("yispred") = cube_of(yispred.data(), nobs, dim, no), Rcpp::Named("ypred") = ppred_out,
      Rcpp::Named("clupred") = mat_of(clupred.data(), (arma::uword)no * npred, A),
      Rcpp::Named("nclus") = mat_of(nclus.data(), A, no), Rcpp::Named("WAIC") = WAIC,
      Rcpp::Named("lpml") = vec_of(lpml.data(), no), Rcpp::Named("theta_prior") = mat_of(Theta.data(), dim, A),
      Rcpp::Named("sigma_prior") = cube_of(Sigma.data(), dim, dim, A));
}

// [[Rcpp::export]]
Rcpp::List prior_ppmx_core(int iter, int burn, int thin, int nobs, int PPMx, int ncon, int ncat, double alpha,
                           double sigma, arma::mat Vwm, int cohesion, int CC, int consim, int similarity,
                           int calibration, int coardegree, arma::vec xcon, arma::vec similparam,
                           arma::vec curr_cluster, arma::vec card_cluster, int ncluster_curr) {
  tppmx_ctx *ctx = shim_ctx();
  const int nout = thin > 0 ? (iter - burn) / thin : 0, no = nout > 0 ? nout : 0;
  tppmx_prior_args in = {};
  in.iter = iter; in.burn = burn; in.thin = thin; in.nobs = nobs; in.PPMx = PPMx; in.ncon = ncon; in.ncat = ncat;
  in.alpha = alpha; in.sigma = sigma;
  in.Vwm = Vwm.n_elem ? Vwm.memptr() : nullptr; in.Vwm_rows = (int)Vwm.n_rows; in.Vwm_cols = (int)Vwm.n_cols;
  in.cohesion = cohesion; in.CC = CC; in.consim = consim; in.similarity = similarity;
  in.calibration = calibration; in.coardegree = coardegree;
  in.xcon = xcon.memptr(); in.similparam = similparam.memptr();
  in.curr_cluster = curr_cluster.memptr(); in.card_cluster = card_cluster.memptr(); in.ncluster_curr = ncluster_curr;
  in.seed = seed_from_r_rng();
  in.n_chains = 1;
  std::vector<double> nclu(no + 1), njout((size_t)nobs * no + 1);
  tppmx_prior_out out = {nclu.data(), njout.data()};
  if (tppmx_prior_ppmx_core(ctx, &in, &out) != TPPMX_OK) Rcpp::stop(tppmx_last_error(ctx));
  return Rcpp::List::create(Rcpp::Named("nclu") = vec_of(nclu.data(), no),       // :463-465
                            Rcpp::Named("njout") = mat_of(njout.data(), nobs, no));
}
