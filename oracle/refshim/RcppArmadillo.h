/*
 * oracle/refshim/RcppArmadillo.h — TEST INFRASTRUCTURE ONLY.
 *
 * A small stand-in for the parts of <RcppArmadillo.h>, <R.h>, <Rmath.h> that the reference's own
 * sources (src/utils_ct.cpp, src/dm_ppmx_ct.cpp, src/prior_ppmx.cpp under /root/reference) use, so
 * that those files compile UNMODIFIED, from where they lie, into oracle/_ref/libref.so with plain
 * g++ (oracle/Makefile target `ref`).  R, Rcpp, Armadillo and libRmath are absent from this image;
 * nothing here is copied from them: dense column-major containers with eager arithmetic, a
 * heterogeneous list, and the handful of R math functions the path calls, restated from their
 * documented definitions.
 *
 * What this buys: the oracle (oracle/ppmx_oracle.c, ppmx_chain.c) is checked against the
 * reference's own statements executed here, not only against closed-form known answers.
 *
 * Random numbers.  The reference draws from R's global RNG (R::runif, R::rnorm, R::rgamma,
 * arma::mvnrnd, arma::wishrnd, fill::randu, Rcpp sample, rmultinom).  Here every one of those pops
 * the next value of a TAPE (ref_tape_*): the oracle records the variates it consumes, in call
 * order, and the reference is then driven by exactly those variates.  A pop of the wrong kind
 * (uniform / normal / gamma) or past the end throws: the draw ORDER of the oracle is thereby pinned
 * to the reference's too.  Constructions on top of the primitive variates (Cholesky-based mvnrnd,
 * Bartlett wishrnd, inverse-CDF sample / rmultinom) are library stand-ins, stated the same way the
 * oracle states them.
 */
#ifndef REFSHIM_RCPPARMADILLO_H
#define REFSHIM_RCPPARMADILLO_H

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstring>
#include <initializer_list>
#include <iostream>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#ifndef M_LN_SQRT_2PI
#define M_LN_SQRT_2PI 0.918938533204672741780329736406 /* log(sqrt(2*pi)), Rmath.h */
#endif
#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ------------------------------------------------------------------------------------------------
 * the tape of random variates (defined in oracle/refshim/ref_driver.cpp)
 * ---------------------------------------------------------------------------------------------- */
namespace refshim {
enum { TAPE_U = 1, TAPE_N = 2, TAPE_G = 3 };
double tape_pop(int kind, double aux); /* throws std::runtime_error on kind mismatch / exhaustion */
double gammainc_lower(double a, double x);      /* regularised P(a, x) */
double gammainc_lower_inv(double a, double p);  /* x with P(a, x) = p */
inline double unif() { return tape_pop(TAPE_U, 0.0); }
inline double norm() { return tape_pop(TAPE_N, 0.0); }
inline double gamma1(double shape) { return tape_pop(TAPE_G, shape); } /* Gamma(shape, scale 1) */
}  // namespace refshim

/* ------------------------------------------------------------------------------------------------
 * arma:: — dense double matrices, column-major, eager evaluation
 * ---------------------------------------------------------------------------------------------- */
namespace arma {

typedef unsigned long long uword;

namespace fill {
struct zeros_t {};
struct ones_t {};
struct randu_t {};
static const zeros_t zeros = zeros_t();
static const ones_t ones = ones_t();
static const randu_t randu = randu_t();
}  // namespace fill

struct span {
  bool whole;
  uword a, b;
  span() : whole(true), a(0), b(0) {}
  span(uword a_, uword b_) : whole(false), a(a_), b(b_) {}
  static const span all;
};
inline const span span::all = span();

class mat;
class vec;
class rowvec;

static inline void shim_check(bool ok, const char *what) {
  if (!ok) throw std::logic_error(std::string("refshim: ") + what);
}

/* rectangular window of a mat; assignable, converts to mat */
class subview {
 public:
  mat &m;
  uword r0, c0, n_rows, n_cols, n_elem;
  subview(mat &m_, uword r0_, uword c0_, uword nr, uword nc);
  subview(const subview &) = default;
  double &at(uword i, uword j) const;
  double &operator()(uword i) const;
  double &operator()(uword i, uword j) const { return at(i, j); }
  subview &operator=(const subview &o);
  subview &operator=(const mat &o);
  subview &operator=(double v);
  void fill(double v) const;
  mat t() const;
  subview row(uword i) const { return subview(m, r0 + i, c0, 1, n_cols); }
  subview col(uword j) const { return subview(m, r0, c0 + j, n_rows, 1); }
};

class mat {
 public:
  uword n_rows, n_cols, n_elem;
  std::vector<double> d;

  mat() : n_rows(0), n_cols(0), n_elem(0) {}
  mat(uword r, uword c) : n_rows(r), n_cols(c), n_elem(r * c), d(r * c, 0.0) {}
  mat(uword r, uword c, fill::zeros_t) : n_rows(r), n_cols(c), n_elem(r * c), d(r * c, 0.0) {}
  mat(uword r, uword c, fill::ones_t) : n_rows(r), n_cols(c), n_elem(r * c), d(r * c, 1.0) {}
  mat(uword r, uword c, fill::randu_t) : n_rows(r), n_cols(c), n_elem(r * c), d(r * c, 0.0) {
    for (uword i = 0; i < n_elem; i++) d[i] = refshim::unif();
  }
  mat(const subview &s) : n_rows(s.n_rows), n_cols(s.n_cols), n_elem(s.n_elem), d(s.n_elem) {
    for (uword j = 0; j < n_cols; j++)
      for (uword i = 0; i < n_rows; i++) d[i + n_rows * j] = s.at(i, j);
  }
  void set_size(uword r, uword c) {
    n_rows = r; n_cols = c; n_elem = r * c;
    d.assign(n_elem, 0.0);
  }
  double &operator()(uword i) {
    shim_check(i < n_elem, "mat(i): index out of bounds");
    return d[i];
  }
  const double &operator()(uword i) const {
    shim_check(i < n_elem, "mat(i): index out of bounds");
    return d[i];
  }
  double &operator[](uword i) { return d[i]; }
  const double &operator[](uword i) const { return d[i]; }
  double &operator()(uword i, uword j) {
    shim_check(i < n_rows && j < n_cols, "mat(i,j): index out of bounds");
    return d[i + n_rows * j];
  }
  const double &operator()(uword i, uword j) const {
    shim_check(i < n_rows && j < n_cols, "mat(i,j): index out of bounds");
    return d[i + n_rows * j];
  }
  subview operator()(const span &rs, const span &cs) {
    const uword ra = rs.whole ? 0 : rs.a, rb = rs.whole ? n_rows - 1 : rs.b;
    const uword ca = cs.whole ? 0 : cs.a, cb = cs.whole ? n_cols - 1 : cs.b;
    return subview(*this, ra, ca, rb - ra + 1, cb - ca + 1);
  }
  subview row(uword i) { return subview(*this, i, 0, 1, n_cols); }
  subview col(uword j) { return subview(*this, 0, j, n_rows, 1); }
  subview cols(const span &cs) { return (*this)(span(), cs); }
  subview rows(const span &rs) { return (*this)(rs, span()); }
  /* const access returns copies */
  mat row(uword i) const { return mat(subview(const_cast<mat &>(*this), i, 0, 1, n_cols)); }
  mat col(uword j) const { return mat(subview(const_cast<mat &>(*this), 0, j, n_rows, 1)); }
  uword size() const { return n_elem; }
  double *memptr() { return d.data(); }
  const double *memptr() const { return d.data(); }
  double *begin() { return d.data(); }
  double *end() { return d.data() + n_elem; }
  void fill(double v) { std::fill(d.begin(), d.end(), v); }
  void zeros() { fill(0.0); }
  void ones() { fill(1.0); }
  double max() const {
    shim_check(n_elem > 0, "max(): empty object");
    return *std::max_element(d.begin(), d.end());
  }
  double min() const {
    shim_check(n_elem > 0, "min(): empty object");
    return *std::min_element(d.begin(), d.end());
  }
  mat t() const {
    mat o(n_cols, n_rows);
    for (uword j = 0; j < n_cols; j++)
      for (uword i = 0; i < n_rows; i++) o.d[j + n_cols * i] = d[i + n_rows * j];
    return o;
  }
  mat diag() const {
    const uword n = std::min(n_rows, n_cols);
    mat o(n, 1);
    for (uword i = 0; i < n; i++) o.d[i] = d[i + n_rows * i];
    return o;
  }
  const mat &eval() const { return *this; }
  mat &eval() { return *this; }
};

class vec : public mat {
 public:
  vec() : mat(0, 1) {}
  explicit vec(uword n) : mat(n, 1) {}
  vec(uword n, fill::zeros_t f) : mat(n, 1, f) {}
  vec(uword n, fill::ones_t f) : mat(n, 1, f) {}
  vec(uword n, fill::randu_t f) : mat(n, 1, f) {}
  vec(const mat &m) : mat(m) { reshape_col(); }
  vec(const subview &s) : mat(s) { reshape_col(); }
  vec(std::initializer_list<double> l) : mat(l.size(), 1) { std::copy(l.begin(), l.end(), d.begin()); }
  vec &operator=(const mat &m) {
    mat::operator=(m);
    reshape_col();
    return *this;
  }
  vec &operator=(const subview &s) { return (*this) = mat(s); }

 private:
  void reshape_col() {
    shim_check(n_rows == 1 || n_cols == 1 || n_elem == 0, "vec: not a vector");
    n_rows = n_elem;
    n_cols = 1;
  }
};

class rowvec : public mat {
 public:
  rowvec() : mat(1, 0) {}
  explicit rowvec(uword n) : mat(1, n) {}
  rowvec(const mat &m) : mat(m) { reshape_row(); }
  rowvec(const subview &s) : mat(s) { reshape_row(); }
  rowvec &operator=(const mat &m) {
    mat::operator=(m);
    reshape_row();
    return *this;
  }
  rowvec &operator=(const subview &s) { return (*this) = mat(s); }

 private:
  void reshape_row() {
    shim_check(n_rows == 1 || n_cols == 1 || n_elem == 0, "rowvec: not a vector");
    n_cols = n_elem;
    n_rows = 1;
  }
};

/* subview members */
inline subview::subview(mat &m_, uword r0_, uword c0_, uword nr, uword nc)
    : m(m_), r0(r0_), c0(c0_), n_rows(nr), n_cols(nc), n_elem(nr * nc) {
  shim_check(r0 + nr <= m.n_rows && c0 + nc <= m.n_cols, "subview: indices out of bounds");
}
inline double &subview::at(uword i, uword j) const {
  shim_check(i < n_rows && j < n_cols, "subview(i,j): index out of bounds");
  return m.d[(r0 + i) + m.n_rows * (c0 + j)];
}
inline double &subview::operator()(uword i) const {
  shim_check(i < n_elem, "subview(i): index out of bounds");
  return at(i % n_rows, i / n_rows);
}
inline subview &subview::operator=(const mat &o) {
  shim_check(o.n_rows == n_rows && o.n_cols == n_cols, "subview = mat: incompatible sizes");
  for (uword j = 0; j < n_cols; j++)
    for (uword i = 0; i < n_rows; i++) at(i, j) = o.d[i + o.n_rows * j];
  return *this;
}
inline subview &subview::operator=(const subview &o) { return (*this) = mat(o); }
inline subview &subview::operator=(double v) {
  shim_check(n_elem == 1, "subview = scalar: the window must be 1x1");
  at(0, 0) = v;
  return *this;
}
inline void subview::fill(double v) const {
  for (uword j = 0; j < n_cols; j++)
    for (uword i = 0; i < n_rows; i++) at(i, j) = v;
}
inline mat subview::t() const { return mat(*this).t(); }

/* element-wise and matrix arithmetic (non-template on purpose: subviews convert implicitly) */
#define REFSHIM_EW(NAME, EXPR)                                                                   \
  inline mat NAME(const mat &a, const mat &b) {                                                  \
    shim_check(a.n_rows == b.n_rows && a.n_cols == b.n_cols, #NAME ": incompatible sizes");      \
    mat o(a.n_rows, a.n_cols);                                                                   \
    for (uword i = 0; i < a.n_elem; i++) {                                                       \
      const double x = a.d[i], y = b.d[i];                                                       \
      o.d[i] = (EXPR);                                                                           \
    }                                                                                            \
    return o;                                                                                    \
  }                                                                                              \
  inline mat NAME(const mat &a, double y) {                                                      \
    mat o(a.n_rows, a.n_cols);                                                                   \
    for (uword i = 0; i < a.n_elem; i++) {                                                       \
      const double x = a.d[i];                                                                   \
      o.d[i] = (EXPR);                                                                           \
    }                                                                                            \
    return o;                                                                                    \
  }                                                                                              \
  inline mat NAME(double x, const mat &b) {                                                      \
    mat o(b.n_rows, b.n_cols);                                                                   \
    for (uword i = 0; i < b.n_elem; i++) {                                                       \
      const double y = b.d[i];                                                                   \
      o.d[i] = (EXPR);                                                                           \
    }                                                                                            \
    return o;                                                                                    \
  }
REFSHIM_EW(operator+, x + y)
REFSHIM_EW(operator-, x - y)
REFSHIM_EW(operator/, x / y)
REFSHIM_EW(operator%, x *y)
#undef REFSHIM_EW
inline mat operator*(const mat &a, double y) {
  mat o(a.n_rows, a.n_cols);
  for (uword i = 0; i < a.n_elem; i++) o.d[i] = a.d[i] * y;
  return o;
}
inline mat operator*(double x, const mat &b) {
  mat o(b.n_rows, b.n_cols);
  for (uword i = 0; i < b.n_elem; i++) o.d[i] = x * b.d[i];
  return o;
}
inline mat operator*(const mat &a, const mat &b) {
  shim_check(a.n_cols == b.n_rows, "matrix product: incompatible sizes");
  mat o(a.n_rows, b.n_cols);
  for (uword j = 0; j < b.n_cols; j++)
    for (uword i = 0; i < a.n_rows; i++) {
      double acc = 0.0;
      for (uword k = 0; k < a.n_cols; k++) acc += a.d[i + a.n_rows * k] * b.d[k + b.n_rows * j];
      o.d[i + o.n_rows * j] = acc;
    }
  return o;
}
inline mat operator-(const mat &a) {
  mat o(a.n_rows, a.n_cols);
  for (uword i = 0; i < a.n_elem; i++) o.d[i] = -a.d[i];
  return o;
}
#define REFSHIM_MAP(NAME, EXPR)                      \
  inline mat NAME(const mat &a) {                    \
    mat o(a.n_rows, a.n_cols);                       \
    for (uword i = 0; i < a.n_elem; i++) {           \
      const double x = a.d[i];                       \
      o.d[i] = (EXPR);                               \
    }                                                \
    return o;                                        \
  }
REFSHIM_MAP(log, std::log(x))
REFSHIM_MAP(exp, std::exp(x))
REFSHIM_MAP(sqrt, std::sqrt(x))
REFSHIM_MAP(lgamma, ::lgamma(x))
#undef REFSHIM_MAP
inline mat pow(const mat &a, double e) {
  mat o(a.n_rows, a.n_cols);
  for (uword i = 0; i < a.n_elem; i++) o.d[i] = std::pow(a.d[i], e);
  return o;
}
inline std::ostream &operator<<(std::ostream &os, const mat &a) {
  for (uword i = 0; i < a.n_rows; i++) {
    for (uword j = 0; j < a.n_cols; j++) os << ' ' << a.d[i + a.n_rows * j];
    os << '\n';
  }
  return os;
}

inline double sum(const mat &a) { /* every call site sums a vector */
  double s = 0.0;
  for (uword i = 0; i < a.n_elem; i++) s += a.d[i];
  return s;
}
inline double accu(const mat &a) { return sum(a); }
inline mat trans(const mat &a) { return a.t(); }
/* mean over dimension dim (0: of each column -> row vector, 1: of each row -> column vector) */
inline mat mean(const mat &a, int dim) {
  if (dim == 0) {
    mat o(1, a.n_cols);
    for (uword j = 0; j < a.n_cols; j++) {
      double s = 0.0;
      for (uword i = 0; i < a.n_rows; i++) s += a(i, j);
      o.d[j] = s / (double)a.n_rows;
    }
    return o;
  }
  mat o(a.n_rows, 1);
  for (uword i = 0; i < a.n_rows; i++) {
    double s = 0.0;
    for (uword j = 0; j < a.n_cols; j++) s += a(i, j);
    o.d[i] = s / (double)a.n_cols;
  }
  return o;
}
/* covariance of the columns of X (rows = observations); norm_type 0: /(N-1), 1: /N */
inline mat cov(const mat &X, int norm_type = 0) {
  const uword N = X.n_rows, p = X.n_cols;
  const mat mu = mean(X, 0);
  const double den = norm_type == 1 ? (double)N : (double)(N > 1 ? N - 1 : 1);
  mat o(p, p);
  for (uword a = 0; a < p; a++)
    for (uword b = 0; b < p; b++) {
      double s = 0.0;
      for (uword i = 0; i < N; i++) s += (X(i, a) - mu.d[a]) * (X(i, b) - mu.d[b]);
      o(a, b) = s / den;
    }
  return o;
}
/* lower Cholesky factor L (A = L L') */
inline mat chol_lower(const mat &A) {
  shim_check(A.n_rows == A.n_cols, "chol(): matrix must be square");
  const uword n = A.n_rows;
  mat L(n, n);
  for (uword j = 0; j < n; j++) {
    double s = A(j, j);
    for (uword k = 0; k < j; k++) s -= L(j, k) * L(j, k);
    if (!(s > 0.0)) throw std::runtime_error("chol(): decomposition failed");
    const double dj = std::sqrt(s);
    L(j, j) = dj;
    for (uword i = j + 1; i < n; i++) {
      double t = A(i, j);
      for (uword k = 0; k < j; k++) t -= L(i, k) * L(j, k);
      L(i, j) = t / dj;
    }
  }
  return L;
}
inline mat chol(const mat &A) { return chol_lower(A).t(); } /* upper: A = R'R */
inline mat trimatu(const mat &A) {
  mat o(A.n_rows, A.n_cols);
  for (uword j = 0; j < A.n_cols; j++)
    for (uword i = 0; i <= j && i < A.n_rows; i++) o(i, j) = A(i, j);
  return o;
}
/* general inverse, Gauss-Jordan with partial pivoting */
inline mat inv(const mat &A) {
  shim_check(A.n_rows == A.n_cols, "inv(): matrix must be square");
  const uword n = A.n_rows;
  mat a(A), o(n, n);
  for (uword i = 0; i < n; i++) o(i, i) = 1.0;
  for (uword c = 0; c < n; c++) {
    uword piv = c;
    for (uword r = c + 1; r < n; r++)
      if (std::fabs(a(r, c)) > std::fabs(a(piv, c))) piv = r;
    if (a(piv, c) == 0.0) throw std::runtime_error("inv(): matrix is singular");
    if (piv != c)
      for (uword j = 0; j < n; j++) {
        std::swap(a(c, j), a(piv, j));
        std::swap(o(c, j), o(piv, j));
      }
    const double p = a(c, c);
    for (uword j = 0; j < n; j++) {
      a(c, j) /= p;
      o(c, j) /= p;
    }
    for (uword r = 0; r < n; r++)
      if (r != c) {
        const double f = a(r, c);
        if (f != 0.0)
          for (uword j = 0; j < n; j++) {
            a(r, j) -= f * a(c, j);
            o(r, j) -= f * o(c, j);
          }
      }
  }
  return o;
}
/* N(M, C): M + chol_lower(C) z, z_k = the next tape normals in order k = 0..D-1 */
inline vec mvnrnd(const mat &M, const mat &C) {
  const uword n = M.n_elem;
  const mat L = chol_lower(C);
  std::vector<double> z(n);
  for (uword k = 0; k < n; k++) z[k] = refshim::norm();
  vec o(n);
  for (uword k = 0; k < n; k++) {
    double acc = M.d[k];
    for (uword r = 0; r <= k; r++) acc += L(k, r) * z[r];
    o.d[k] = acc;
  }
  return o;
}
/* Wishart(S, df) by Bartlett: W = (L A)(L A)', L = chol_lower(S), A lower triangular with
 * A_kk = sqrt(2 Gamma((df - k)/2)) and A_kr ~ N(0,1), r < k; row by row, diagonal first */
inline mat wishrnd(const mat &S, double df) {
  const uword n = S.n_rows;
  const mat L = chol_lower(S);
  mat A(n, n);
  for (uword k = 0; k < n; k++) {
    A(k, k) = std::sqrt(2.0 * refshim::gamma1(0.5 * (df - (double)k)));
    for (uword r = 0; r < k; r++) A(k, r) = refshim::norm();
  }
  const mat LA = L * A;
  return LA * LA.t();
}

class cube {
 public:
  uword n_rows, n_cols, n_slices, n_elem;
  std::vector<mat> s;
  cube() : n_rows(0), n_cols(0), n_slices(0), n_elem(0) {}
  cube(uword r, uword c, uword n) { init(r, c, n, 0.0); }
  cube(uword r, uword c, uword n, fill::zeros_t) { init(r, c, n, 0.0); }
  cube(uword r, uword c, uword n, fill::ones_t) { init(r, c, n, 1.0); }
  cube(uword r, uword c, uword n, fill::randu_t) {
    init(r, c, n, 0.0);
    for (uword k = 0; k < n; k++)
      for (uword i = 0; i < r * c; i++) s[k].d[i] = refshim::unif();
  }
  mat &slice(uword k) {
    shim_check(k < n_slices, "cube.slice(): index out of bounds");
    return s[k];
  }
  const mat &slice(uword k) const {
    shim_check(k < n_slices, "cube.slice(): index out of bounds");
    return s[k];
  }
  double &operator()(uword i, uword j, uword k) { return slice(k)(i, j); }
  const double &operator()(uword i, uword j, uword k) const { return slice(k)(i, j); }
  void fill(double v) {
    for (uword k = 0; k < n_slices; k++) s[k].fill(v);
  }

 private:
  void init(uword r, uword c, uword n, double v) {
    n_rows = r; n_cols = c; n_slices = n; n_elem = r * c * n;
    s.assign(n, mat(r, c));
    if (v != 0.0) fill(v);
  }
};

template <class T>
class field {
 public:
  uword n_rows, n_cols, n_elem;
  std::vector<T> e;
  field() : n_rows(0), n_cols(0), n_elem(0) {}
  explicit field(uword r) : n_rows(r), n_cols(1), n_elem(r), e(r) {}
  field(uword r, uword c) : n_rows(r), n_cols(c), n_elem(r * c), e(r * c) {}
  T &operator()(uword i) { return e.at(i); }
  const T &operator()(uword i) const { return e.at(i); }
  T &operator()(uword i, uword j) {
    shim_check(i < n_rows && j < n_cols, "field(i,j): index out of bounds");
    return e[i + n_rows * j];
  }
  const T &operator()(uword i, uword j) const {
    shim_check(i < n_rows && j < n_cols, "field(i,j): index out of bounds");
    return e[i + n_rows * j];
  }
};

class uvec {
 public:
  std::vector<uword> d;
  uword n_elem;
  uvec() : n_elem(0) {}
  explicit uvec(uword n) : d(n, 0), n_elem(n) {}
  uword &operator()(uword i) { return d.at(i); }
  const uword &operator()(uword i) const { return d.at(i); }
  uword size() const { return n_elem; }
};

template <class T>
inline T linspace(double a, double b, uword n);
template <>
inline uvec linspace<uvec>(double a, double b, uword n) {
  uvec o(n);
  for (uword i = 0; i < n; i++) o.d[i] = (uword)(n > 1 ? a + (b - a) * (double)i / (double)(n - 1) : b);
  return o;
}
template <>
inline vec linspace<vec>(double a, double b, uword n) {
  vec o(n);
  for (uword i = 0; i < n; i++) o.d[i] = n > 1 ? a + (b - a) * (double)i / (double)(n - 1) : b;
  return o;
}
template <class T>
struct conv_to;
template <>
struct conv_to<uword> {
  static uword from(const uvec &v) {
    shim_check(v.n_elem == 1, "conv_to<uword>::from(): object must have exactly one element");
    return v.d[0];
  }
};

}  // namespace arma

/* ------------------------------------------------------------------------------------------------
 * R:: — the Rmath functions the sources call
 * ---------------------------------------------------------------------------------------------- */
namespace R {
inline double lgammafn(double x) { return ::lgamma(x); }
/* dnorm(x, mu, sigma, give_log): -(log sqrt(2 pi) + log sigma + z^2 / 2) */
inline double dnorm(double x, double mu, double sigma, int give_log) {
  const double z = (x - mu) / sigma;
  const double ld = -(M_LN_SQRT_2PI + 0.5 * z * z + std::log(sigma));
  return give_log ? ld : std::exp(ld);
}
inline double dnorm4(double x, double mu, double sigma, int give_log) { return dnorm(x, mu, sigma, give_log); }
inline double runif(double a, double b) { return a + (b - a) * refshim::unif(); }
inline double rnorm(double mu, double sigma) { return mu + sigma * refshim::norm(); }
inline double norm_rand() { return refshim::norm(); }
inline double unif_rand() { return refshim::unif(); }
inline double rgamma(double shape, double scale) { return refshim::gamma1(shape) * scale; }
inline double pgamma(double x, double shape, double scale, int lower, int give_log) {
  double p = refshim::gammainc_lower(shape, x / scale);
  if (!lower) p = 1.0 - p;
  return give_log ? std::log(p) : p;
}
inline double qgamma(double p, double shape, double scale, int lower, int give_log) {
  if (give_log) p = std::exp(p);
  if (!lower) p = 1.0 - p;
  return scale * refshim::gammainc_lower_inv(shape, p);
}
}  // namespace R

/* rmultinom(n, prob, K, rN) of Rmath.h (global, as the macro Rf_rmultinom -> rmultinom).  Only
 * n = 1 is used by the sources: one inverse-CDF uniform picks the category (the oracle's form). */
inline void rmultinom(int n, double *prob, int K, int *rN) {
  for (int k = 0; k < K; k++) rN[k] = 0;
  for (int r = 0; r < n; r++) {
    const double u = refshim::unif();
    double cw = 0.0;
    int pick = K - 1;
    for (int k = 0; k < K; k++) {
      cw += prob[k];
      if (u < cw) {
        pick = k;
        break;
      }
    }
    rN[pick] += 1;
  }
}

/* ------------------------------------------------------------------------------------------------
 * Rcpp:: — vectors, the generic list, wrap / as
 * ---------------------------------------------------------------------------------------------- */
namespace Rcpp {

static std::ostream &Rcout = std::cout;
static std::ostream &Rcerr = std::cerr;

class IntegerVector {
 public:
  std::vector<int> d;
  IntegerVector() {}
  explicit IntegerVector(int n) : d(n, 0) {}
  int *begin() { return d.data(); }
  int length() const { return (int)d.size(); }
  int size() const { return (int)d.size(); }
  int &operator[](int i) { return d[i]; }
  int &operator()(int i) { return d.at(i); }
};
class NumericVector {
 public:
  std::vector<double> d;
  NumericVector() {}
  explicit NumericVector(int n) : d(n, 0.0) {}
  NumericVector(const IntegerVector &v) : d(v.d.begin(), v.d.end()) {}
  double *begin() { return d.data(); }
  int length() const { return (int)d.size(); }
  int size() const { return (int)d.size(); }
  double &operator[](int i) { return d[i]; }
  double &operator()(int i) { return d.at(i); }
};

/* one value of any of the kinds the sources put into a list */
struct Value {
  enum Kind { NONE, NUM, MAT, CUBE, FIELD_MAT, FIELD_CUBE } kind;
  std::string name;
  double num;
  arma::mat m;
  bool m_is_col;
  arma::cube c;
  arma::field<arma::mat> fm;
  arma::field<arma::cube> fc;
  Value() : kind(NONE), num(0.0), m_is_col(false) {}
  Value(double v) : kind(NUM), num(v), m_is_col(false) {}
  Value(int v) : kind(NUM), num(v), m_is_col(false) {}
  Value(const arma::vec &v) : kind(MAT), num(0.0), m(v), m_is_col(true) {}
  Value(const arma::rowvec &v) : kind(MAT), num(0.0), m(v), m_is_col(false) {}
  Value(const arma::mat &v) : kind(MAT), num(0.0), m(v), m_is_col(false) {}
  Value(const arma::subview &v) : kind(MAT), num(0.0), m(v), m_is_col(false) {}
  Value(const arma::cube &v) : kind(CUBE), num(0.0), m_is_col(false), c(v) {}
  Value(const arma::field<arma::mat> &v) : kind(FIELD_MAT), num(0.0), m_is_col(false), fm(v) {}
  Value(const arma::field<arma::cube> &v) : kind(FIELD_CUBE), num(0.0), m_is_col(false), fc(v) {}
  Value(const NumericVector &v) : kind(MAT), num(0.0), m(v.d.size(), 1), m_is_col(true) {
    std::copy(v.d.begin(), v.d.end(), m.d.begin());
  }
};
typedef Value RObject;

template <class T>
inline Value wrap(const T &x) { return Value(x); }

template <class T>
struct as_impl;
template <>
struct as_impl<arma::mat> {
  static arma::mat get(const Value &v) {
    arma::shim_check(v.kind == Value::MAT, "as<mat>: not a numeric matrix");
    return v.m;
  }
};
template <>
struct as_impl<arma::vec> {
  static arma::vec get(const Value &v) {
    arma::shim_check(v.kind == Value::MAT, "as<vec>: not numeric");
    return arma::vec(v.m);
  }
};
template <>
struct as_impl<arma::rowvec> {
  static arma::rowvec get(const Value &v) {
    arma::shim_check(v.kind == Value::MAT, "as<rowvec>: not numeric");
    return arma::rowvec(v.m);
  }
};
template <>
struct as_impl<NumericVector> {
  static NumericVector get(const Value &v) {
    arma::shim_check(v.kind == Value::MAT, "as<NumericVector>: not numeric");
    NumericVector o((int)v.m.n_elem);
    std::copy(v.m.d.begin(), v.m.d.end(), o.d.begin());
    return o;
  }
};
template <>
struct as_impl<double> {
  static double get(const Value &v) {
    arma::shim_check(v.kind == Value::NUM, "as<double>: not a scalar");
    return v.num;
  }
};
template <class T>
inline T as(const Value &v) { return as_impl<T>::get(v); }
template <class T>
inline T as(const NumericVector &v) { return as_impl<T>::get(Value(v)); }
template <class T>
inline T as(const IntegerVector &v) { return as_impl<T>::get(Value(NumericVector(v))); }

struct Named {
  std::string name;
  explicit Named(const std::string &n) : name(n) {}
  template <class T>
  Value operator=(const T &x) const {
    Value v(x);
    v.name = name;
    return v;
  }
};

class List {
 public:
  std::vector<Value> e;
  List() {}
  explicit List(int n) : e(n) {}
  Value &operator[](int i) { return e.at(i); }
  const Value &operator[](int i) const { return e.at(i); }
  const Value &operator[](const std::string &name) const {
    for (size_t i = 0; i < e.size(); i++)
      if (e[i].name == name) return e[i];
    throw std::out_of_range("List: no element named " + name);
  }
  int size() const { return (int)e.size(); }
  int length() const { return (int)e.size(); }
  template <class... A>
  static List create(const A &...a) {
    List l;
    const Value vs[] = {Value(a)...};
    l.e.assign(vs, vs + sizeof...(A));
    return l;
  }
};

namespace RcppArmadillo {
/* sample(x, size = 1, replace, prob): prob is normalised, then ONE inverse-CDF uniform in the
 * order of x (R walks the probabilities sorted in decreasing order with the same uniform; the law
 * is the same, the oracle uses the plain order) */
inline arma::uvec sample(const arma::uvec &x, int size, bool /*replace*/, const arma::vec &prob) {
  arma::shim_check(size == 1, "sample(): only size = 1 is used by the sources");
  arma::shim_check(prob.n_elem == x.n_elem, "sample(): prob and x differ in length");
  const double den = arma::sum(prob);
  const double u = refshim::unif();
  arma::uword pick = x.n_elem - 1;
  double cw = 0.0;
  for (arma::uword g = 0; g < x.n_elem; g++) {
    cw += prob.d[g] / den;
    if (u < cw) {
      pick = g;
      break;
    }
  }
  arma::uvec o(1);
  o.d[0] = x.d[pick];
  return o;
}
}  // namespace RcppArmadillo

inline void checkUserInterrupt() {}
inline void stop(const std::string &msg) { throw std::runtime_error(msg); }

}  // namespace Rcpp

#endif
