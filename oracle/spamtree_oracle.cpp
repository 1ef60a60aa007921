// oracle/spamtree_oracle.cpp
//
// TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's (mkln/spamtree
// v0.2.1) per-iteration MCMC hot path, used (a) as the checker in tests/,
// (b) by __graft_entry__.smoke(), (c) as bench.py's `cpu_baseline` / `--impl
// reference` arm.  The product (spamtree_b200/) never links, imports or calls
// anything in this directory.
//
// PINNED AGAINST THE REFERENCE ITSELF: the reference ships no tests, golden
// vectors or fixtures and cannot be built with its own
// toolchain here (needs R + Rcpp + RcppArmadillo + BLAS/LAPACK), but its model
// layer compiles UNMODIFIED from /root/reference/src against the Armadillo/Rcpp
// stand-in header oracle/refshim/ (`make -C oracle ref` -> oracle/_ref/), and
// tests/test_oracle_vs_reference.py checks this restatement against it: integer
// bookkeeping bit-exact, H / Ri / Kxx_inv / log-density / Gibbs draws / predict /
// beta / tausq to <= 1e-10 for q = 1, 2, 3, 5.  (The stand-in's dpotrf/dtrtri/dgemm
// are plain loops, so the rounding of the reference's real BLAS is not pinned.)
// Since round 2 the reference's DRIVER (spamtree_fit.cpp: spamtree_mv_mcmc itself), its
// DAG builders (make_edges, make_edges_limited, part_axis_parallel_lmt) and its
// Metropolis glue (mh_adapt.h) are part of that build as well, and
// tests/test_reference_driver.py compares whole chains of or_mcmc with the reference
// driver's on the shared host random stream (<= 2e-10, typically 1e-14) and the DAG
// builders bit-exactly.
// Further pins: analytic known-answer tests derived from the reference source
// and man pages, and a dense numpy statement of the model's math
// (tests/test_oracle_pinning.py, tests/dense_twin.py).
//
// Arithmetic follows the reference operation by operation (same formulas, same
// operand order at the matrix level); Armadillo/LAPACK calls are replaced by the
// small column-major routines below (dpotrf -> chol_lower, dtrtri -> inv_lower,
// dgemm/dsyrk -> loops).  Every function cites the reference file:line it
// follows (paths relative to /root/reference/).
//
// Build: see oracle/Makefile (g++ -O3 -march=x86-64-v3 -fopenmp -shared; the library travels to the GPU box prebuilt, hence no
// -march=native).  or_use_blas() optionally routes the dense kernels through the host's OpenBLAS for the timing arms.

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#include <dlfcn.h>
#endif

namespace {

typedef std::vector<int64_t> ivec;
typedef std::vector<double> dvec;

// ---------------------------------------------------------------- dense helpers
// column-major matrix, like arma::mat
struct Mat {
  int r = 0, c = 0;
  dvec a;
  Mat() {}
  Mat(int r_, int c_) : r(r_), c(c_), a((size_t)r_ * c_, 0.0) {}
  inline double& operator()(int i, int j) { return a[(size_t)i + (size_t)j * r]; }
  inline double operator()(int i, int j) const { return a[(size_t)i + (size_t)j * r]; }
  bool empty() const { return a.empty(); }
  void clear() { r = c = 0; dvec().swap(a); }
};

// ---- optional BLAS/LAPACK backend ("port+blas" CPU baseline): the same algorithm with its dense kernels routed through an
// OpenBLAS the host already has (scipy's bundled libscipy_openblas: scipy_dgemm_, scipy_dsyrk_, scipy_dpotrf_, scipy_dtrtri_,
// scipy_dgemv_), the way the reference's Armadillo calls reach R's BLAS/LAPACK (SURVEY §8c lists the call sites).
// Off by default: the parity tests use the plain loops below (bit-reproducible); or_use_blas(path) switches it on.
struct BlasApi {
  bool on = false;
  void (*dgemm)(const char*, const char*, const int*, const int*, const int*, const double*, const double*, const int*, const double*,
                const int*, const double*, double*, const int*) = nullptr;
  void (*dsyrk)(const char*, const char*, const int*, const int*, const double*, const double*, const int*, const double*, double*,
                const int*) = nullptr;
  void (*dgemv)(const char*, const int*, const int*, const double*, const double*, const int*, const double*, const int*, const double*,
                double*, const int*) = nullptr;
  void (*dpotrf)(const char*, const int*, double*, const int*, int*) = nullptr;
  void (*dtrtri)(const char*, const char*, const int*, double*, const int*, int*) = nullptr;
  void (*set_threads)(int) = nullptr;
};
static BlasApi g_blas;

// C = A * B
static Mat mm(const Mat& A, const Mat& B) {
  Mat C(A.r, B.c);
  if (g_blas.on && A.r && A.c && B.c) {
    const double one = 1.0, zero = 0.0;
    g_blas.dgemm("N", "N", &A.r, &B.c, &A.c, &one, A.a.data(), &A.r, B.a.data(), &B.r, &zero, C.a.data(), &C.r);
    return C;
  }
  for (int j = 0; j < B.c; j++)
    for (int k = 0; k < A.c; k++) {
      const double b = B(k, j);
      const double* ap = &A.a[(size_t)k * A.r];
      double* cp = &C.a[(size_t)j * C.r];
#pragma omp simd
      for (int i = 0; i < A.r; i++) cp[i] += ap[i] * b;
    }
  return C;
}
// C = A' * B
static Mat mtm(const Mat& A, const Mat& B) {
  Mat C(A.c, B.c);
  if (g_blas.on && A.r && A.c && B.c) {
    const double one = 1.0, zero = 0.0;
    g_blas.dgemm("T", "N", &A.c, &B.c, &A.r, &one, A.a.data(), &A.r, B.a.data(), &B.r, &zero, C.a.data(), &C.r);
    return C;
  }
  for (int j = 0; j < B.c; j++)
    for (int i = 0; i < A.c; i++) {
      const double* ap = &A.a[(size_t)i * A.r];
      const double* bp = &B.a[(size_t)j * B.r];
      double s = 0;
#pragma omp simd reduction(+ : s)
      for (int k = 0; k < A.r; k++) s += ap[k] * bp[k];
      C(i, j) = s;
    }
  return C;
}
// C = L' * L for a lower-triangular L (zeros above the diagonal are skipped:
// identical result to the full product since the skipped terms are exact zeros)
static Mat ltl(const Mat& L) {
  const int n = L.r;
  Mat C(n, n);
  if (g_blas.on && n) {  // arma: A.t() * A -> dsyrk (spamtree_model.cpp:867,906,912,948)
    const double one = 1.0, zero = 0.0;
    g_blas.dsyrk("U", "T", &n, &n, &one, L.a.data(), &n, &zero, C.a.data(), &n);
    for (int j = 0; j < n; j++)
      for (int i = j + 1; i < n; i++) C(i, j) = C(j, i);
    return C;
  }
  for (int j = 0; j < n; j++)
    for (int i = 0; i <= j; i++) {
      const double* ap = &L.a[(size_t)i * n];
      const double* bp = &L.a[(size_t)j * n];
      double s = 0;
#pragma omp simd reduction(+ : s)
      for (int k = j; k < n; k++) s += ap[k] * bp[k];
      C(i, j) = s;
      C(j, i) = s;
    }
  return C;
}
static dvec mv(const Mat& A, const dvec& x) {
  dvec y(A.r, 0.0);
  if (g_blas.on && A.r && A.c) {
    const double one = 1.0, zero = 0.0;
    const int inc = 1;
    g_blas.dgemv("N", &A.r, &A.c, &one, A.a.data(), &A.r, x.data(), &inc, &zero, y.data(), &inc);
    return y;
  }
  for (int k = 0; k < A.c; k++) {
    const double xk = x[k];
    const double* ap = &A.a[(size_t)k * A.r];
    for (int i = 0; i < A.r; i++) y[i] += ap[i] * xk;
  }
  return y;
}
static dvec mtv(const Mat& A, const dvec& x) {
  dvec y(A.c, 0.0);
  if (g_blas.on && A.r && A.c) {
    const double one = 1.0, zero = 0.0;
    const int inc = 1;
    g_blas.dgemv("T", &A.r, &A.c, &one, A.a.data(), &A.r, x.data(), &inc, &zero, y.data(), &inc);
    return y;
  }
  for (int j = 0; j < A.c; j++) {
    const double* ap = &A.a[(size_t)j * A.r];
    double s = 0;
    for (int k = 0; k < A.r; k++) s += ap[k] * x[k];
    y[j] = s;
  }
  return y;
}
// arma::symmatu: reflect upper triangle into lower
static void symmatu(Mat& A) {
  for (int j = 0; j < A.c; j++)
    for (int i = j + 1; i < A.r; i++) A(i, j) = A(j, i);
}
// arma::chol(A, "lower") -> LAPACK dpotrf('L'): false if a pivot is <= 0 or NaN
static bool chol_lower(Mat& A) {
  const int n = A.r;
  if (g_blas.on && n) {
    int info = 0;
    g_blas.dpotrf("L", &n, A.a.data(), &n, &info);
    if (info != 0) return false;
    for (int j = 0; j < n; j++)
      for (int i = 0; i < j; i++) A(i, j) = 0.0;
    return true;
  }
  for (int j = 0; j < n; j++) {
    double ajj = A(j, j);
    for (int k = 0; k < j; k++) ajj -= A(j, k) * A(j, k);
    if (!(ajj > 0.0) || !std::isfinite(ajj)) return false;
    ajj = std::sqrt(ajj);
    A(j, j) = ajj;
    for (int i = j + 1; i < n; i++) {
      double s = A(i, j);
      for (int k = 0; k < j; k++) s -= A(i, k) * A(j, k);
      A(i, j) = s / ajj;
    }
  }
  for (int j = 0; j < n; j++)
    for (int i = 0; i < j; i++) A(i, j) = 0.0;
  return true;
}
// arma::inv(arma::trimatl(L)) -> LAPACK dtrtri('L','N')
static Mat inv_lower(const Mat& L) {
  const int n = L.r;
  Mat X(n, n);
  if (g_blas.on && n) {
    X = L;
    int info = 0;
    g_blas.dtrtri("L", "N", &n, X.a.data(), &n, &info);
    return X;
  }
  for (int j = 0; j < n; j++) {
    X(j, j) = 1.0 / L(j, j);
    for (int i = j + 1; i < n; i++) {
      double s = 0;
      for (int k = j; k < i; k++) s += L(i, k) * X(k, j);
      X(i, j) = -s / L(i, i);
    }
  }
  return X;
}

// ---------------------------------------------------------------- RNG
// Host random stream used by BOTH the oracle chain and the product's "host" RNG
// mode so that the two chains can be run in lock-step (R's RNG, which the
// reference uses through Rcpp, is not available here).  xoshiro256++ seeded by
// splitmix64; normals by Box-Muller (both outputs used); gamma by
// Marsaglia-Tsang.  The product carries its own, independently written copy.
struct Rng {
  uint64_t s[4];
  bool have = false;
  double spare = 0;
  static uint64_t splitmix(uint64_t& x) {
    uint64_t z = (x += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
  }
  void seed(uint64_t sd) {
    for (int i = 0; i < 4; i++) s[i] = splitmix(sd);
    have = false;
  }
  static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
  uint64_t next() {
    const uint64_t result = rotl(s[0] + s[3], 23) + s[0];
    const uint64_t t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3];
    s[2] ^= t; s[3] = rotl(s[3], 45);
    return result;
  }
  double unif() { return ((next() >> 11) + 0.5) * (1.0 / 9007199254740992.0); }
  double norm() {
    if (have) { have = false; return spare; }
    const double u1 = unif(), u2 = unif();
    const double rad = std::sqrt(-2.0 * std::log(u1));
    const double ang = 6.283185307179586476925286766559 * u2;
    spare = rad * std::sin(ang);
    have = true;
    return rad * std::cos(ang);
  }
  double gamma(double shape, double scale) {  // shape >= 1 on this path (2.01 + n/2)
    const double d = shape - 1.0 / 3.0, c = 1.0 / std::sqrt(9.0 * d);
    for (;;) {
      double x, v;
      do { x = norm(); v = 1.0 + c * x; } while (v <= 0);
      v = v * v * v;
      const double u = unif();
      if (u < 1.0 - 0.0331 * x * x * x * x) return d * v * scale;
      if (std::log(u) < 0.5 * x * x + d * (1.0 - v + std::log(v))) return d * v * scale;
    }
  }
};

// ---------------------------------------------------------------- covariance
// covariance_functions.h:7-31, covariance_functions.cpp:10-92
struct CovPars {
  int q = 0, n_cbase = 1, npars = 0;
  dvec ai1, ai2, phi_i, thetamv;
  Mat Dmat;
  void init(int q_in) {  // covariance_functions.cpp:10-32 with dd==2 -> model 0
    q = q_in;
    n_cbase = q > 2 ? 3 : 1;
    npars = 3 * q + n_cbase;
  }
  // covariance_functions.cpp:34-52 and vec_to_symmat :77-92
  void transform(const dvec& theta) {
    int k = (int)theta.size() - npars;
    ai1.assign(theta.begin(), theta.begin() + q);
    ai2.assign(theta.begin() + q, theta.begin() + 2 * q);
    phi_i.assign(theta.begin() + 2 * q, theta.begin() + 3 * q);
    thetamv.assign(theta.begin() + 3 * q, theta.begin() + 3 * q + n_cbase);
    if (k > 0) {
      int pp = (int)((1 + std::sqrt(1.0 + 8.0 * k)) / 2);
      Dmat = Mat(pp, pp);
      int start_i = 1, ix = 0;
      for (int j = 0; j < pp; j++) {
        for (int i = start_i; i < pp; i++) { Dmat(i, j) = theta[npars + ix]; ix++; }
        start_i++;
      }
      for (int j = 0; j < pp; j++)
        for (int i = 0; i < j; i++) Dmat(i, j) = Dmat(j, i);  // symmatl
    } else {
      Dmat = Mat(1, 1);
    }
  }
};
// covariance_functions.h:40-48
static inline double fphi(double x, double c) { return std::exp(-c * x); }
static inline double sqrt_fpsi(double x, double a, double beta) { return std::exp(0.5 * beta * std::log1p(a * x)); }
// covariance_functions.cpp:113-135 (u, dim unused there as well)
static inline double C_base(double h, double v, const dvec& params, int q) {
  if (q > 2) {
    double psi1_sqrt = sqrt_fpsi(v, params[0], params[1]);
    return fphi(h / psi1_sqrt, params[2]) / (psi1_sqrt * psi1_sqrt);
  } else if (q == 2) {
    double psi1_sqrt = std::sqrt(v + 1);
    return fphi(h / psi1_sqrt, params[0]) / (v + 1.0);
  }
  return fphi(h, params[0]);
}

struct Model;
// mvCovAG20107_inplace covariance_functions.cpp:213-286 (+ cexpcov :95-111 for q==1)
static void covariancef(Mat& res, const Model& M, const int64_t* ind1, int n1, const int64_t* ind2, int n2,
                        const CovPars& cp, bool same);

// ---------------------------------------------------------------- model state
// tree_utils.h:63-102
struct Data {
  dvec theta;
  dvec wcore, logdetCi_comps, loglik_w_comps;
  double logdetCi = 0, loglik_w = 0;
  std::vector<Mat> Kxc, Kxx_inv, Kxx_invchol, Rcc_invchol, H /*w_cond_mean_K*/, w_cond_prec;
  std::vector<dvec> w_cond_prec_noref, ccholprecdiag, Sigi_chol_noref;
  std::vector<int> has_updated;
  std::vector<Mat> Sigi_children;  // m x (m*C): slice c at columns [c*m, (c+1)*m)
  std::vector<Mat> Smu_children;   // m x C
  std::vector<Mat> AK_uP_all, AK_uP_u_all, Sigi_chol;
  // test probes (not in the reference): last Sigi_tot / Smu_tot seen by the Gibbs step
  std::vector<Mat> probe_Sigi_tot;
  std::vector<dvec> probe_Smu_tot;
};

struct Model {
  // inputs (spamtree_model.cpp:8-37)
  int64_t n_all = 0;
  int p = 0, q = 0, dd = 2, n_blocks = 0;
  dvec y, X, coords;  // col-major
  ivec mv_id, qv;
  ivec res_is_ref;
  std::vector<ivec> parents, children, indexing;
  bool limited_tree = false;
  dvec block_names, block_groups;
  // flags
  bool lean = false;             // Kxx_inv / Kxx_invchol shared between slots, AK_uP_u_all transient
  bool q1_norm_expansion = false;  // App. D #1: cexpcov's |x|^2+|y|^2-2xy distance instead of direct difference
  bool faithful_beta_index = true; // App. D #12
  bool probes = false;
  // App. D #13: predict(theta_update = false) on a slot whose prediction weights were never computed uses the zeros they
  // were allocated as (spamtree_model.cpp:472-473, :1256-1296): H = 0, Kxc = 0 -> a draw from the marginal N(0, K_ii).
  // The state is per theta-slot and survives accept_make_change's swap.  false = recompute instead (what the product does)
  bool faithful_predict_cache = true;
  // derived
  int64_t n = 0;
  ivec na_ix_all;
  std::vector<ivec> ix_by_q, ix_by_q_a;
  dvec y_available, X_available;
  std::vector<Mat> XtX;
  dvec block_groups_labels;
  int n_gibbs_groups = 0, n_actual_groups = 0;
  std::vector<ivec> parents_indexing, children_indexing, this_is_jth_child, u_by_block_groups;
  std::vector<ivec> dim_by_parent;
  std::vector<std::vector<std::pair<int64_t, int64_t>>> u_is_which_col;  // (firstcol,lastcol) per child
  ivec block_ct_obs, blocks_not_empty, blocks_predicting, block_is_reference;
  // params
  dvec bigrnorm, w, XB, tausq_inv, tausq_inv_long;
  Mat Bcoeff;
  Data data[2];
  int cur = 0;  // param_data = data[cur], alter_data = data[1-cur] (std::swap in the reference)
  std::vector<Mat> scratch_Kxx_inv, scratch_Kxx_invchol;
  CovPars covpars;
  Rng rng;
  std::string err;
  Data& param() { return data[cur]; }
  Data& alter() { return data[1 - cur]; }
};

static void covariancef(Mat& res, const Model& M, const int64_t* ind1, int n1, const int64_t* ind2, int n2,
                        const CovPars& cp, bool same) {
  const double* cx = M.coords.data();
  const double* cy = M.coords.data() + M.n_all;
  const int pdim = cp.Dmat.c;
  if (res.r != n1 || res.c != n2) res = Mat(n1, n2);
  if (pdim < 2) {
    // cexpcov covariance_functions.cpp:95-111 with sigmasq = ai1(0) (unsquared), phi = thetamv(0) (:221)
    const double sigmasq = cp.ai1[0], phi = cp.thetamv[0];
    for (int j = 0; j < n2; j++)
      for (int i = 0; i < n1; i++) {
        const double x1 = cx[ind1[i]], y1 = cy[ind1[i]], x2 = cx[ind2[j]], y2 = cy[ind2[j]];
        double d;
        if (M.q1_norm_expansion) {
          const double pm = x1 * x1 + y1 * y1, qm = x2 * x2 + y2 * y2;
          d = std::sqrt(std::fabs(qm + pm - 2 * (x1 * x2 + y1 * y2)));
        } else {
          const double dx = x1 - x2, dy = y1 - y2;
          d = std::sqrt(dx * dx + dy * dy);
        }
        res(i, j) = sigmasq * std::exp(-phi * d);
      }
    return;
  }
  for (int i = 0; i < n1; i++) {
    const int vi = (int)M.qv[ind1[i]];
    const double ai1_sq = cp.ai1[vi] * cp.ai1[vi], ai2_sq = cp.ai2[vi] * cp.ai2[vi];
    const double xi = cx[ind1[i]], yi = cy[ind1[i]];
    for (int j = (same ? i : 0); j < n2; j++) {
      const double dx = xi - cx[ind2[j]], dy = yi - cy[ind2[j]];
      const double h = std::sqrt(dx * dx + dy * dy);
      const int vj = (int)M.qv[ind2[j]];
      const double v = cp.Dmat(vi, vj);
      if (v == 0) {
        res(i, j) = ai1_sq * C_base(h, 0, cp.thetamv, pdim) + ai2_sq * fphi(h, cp.phi_i[vi]);
      } else {
        res(i, j) = cp.ai1[vi] * cp.ai1[vj] * C_base(h, v, cp.thetamv, pdim);
      }
    }
  }
  if (same) symmatu(res);
}

static dvec rows(const dvec& v, const ivec& ix) {
  dvec o(ix.size());
  for (size_t i = 0; i < ix.size(); i++) o[i] = v[ix[i]];
  return o;
}

// ---------------------------------------------------------------- ctor bookkeeping
// spamtree_model.cpp:8-192 + :194-301 + :303-313 + :315-353 + :355-420 + :422-503
static bool model_init(Model& M, const dvec& theta, const dvec& beta, double tausq_inv_in) {
  const int nb = M.n_blocks;
  M.qv.resize(M.n_all);
  for (int64_t i = 0; i < M.n_all; i++) M.qv[i] = M.mv_id[i] - 1;
  // block_groups_labels = unique(block_groups) (sorted)
  M.block_groups_labels = M.block_groups;
  std::sort(M.block_groups_labels.begin(), M.block_groups_labels.end());
  M.block_groups_labels.erase(std::unique(M.block_groups_labels.begin(), M.block_groups_labels.end()),
                              M.block_groups_labels.end());
  M.n_gibbs_groups = (int)M.block_groups_labels.size();
  // :80-96
  for (int64_t i = 0; i < M.n_all; i++)
    if (std::isfinite(M.y[i])) M.na_ix_all.push_back(i);
  M.n = (int64_t)M.na_ix_all.size();
  M.y_available.resize(M.n);
  M.X_available.resize((size_t)M.n * M.p);
  for (int64_t i = 0; i < M.n; i++) {
    M.y_available[i] = M.y[M.na_ix_all[i]];
    for (int j = 0; j < M.p; j++) M.X_available[i + (size_t)j * M.n] = M.X[M.na_ix_all[i] + (size_t)j * M.n_all];
  }
  M.ix_by_q.assign(M.q, ivec());
  M.ix_by_q_a.assign(M.q, ivec());
  for (int64_t i = 0; i < M.n_all; i++) M.ix_by_q[M.qv[i]].push_back(i);
  for (int64_t i = 0; i < M.n; i++) M.ix_by_q_a[M.qv[M.na_ix_all[i]]].push_back(i);
  // :118-130
  M.tausq_inv.assign(M.q, tausq_inv_in);
  M.tausq_inv_long.assign(M.n_all, tausq_inv_in);
  M.XB.assign(M.n_all, 0.0);
  M.Bcoeff = Mat(M.p, M.q);
  for (int j = 0; j < M.q; j++) {
    for (int64_t r : M.ix_by_q[j]) {
      double s = 0;
      for (int k = 0; k < M.p; k++) s += M.X[r + (size_t)k * M.n_all] * beta[k];
      M.XB[r] = s;
    }
    for (int k = 0; k < M.p; k++) M.Bcoeff(k, j) = beta[k];
  }
  M.w.assign(M.n_all, 0.0);
  // init_indexing :315-353
  M.parents_indexing.assign(nb, ivec());
  M.children_indexing.assign(nb, ivec());
  for (int i = 0; i < nb; i++) {
    int u = (int)M.block_names[i] - 1;
    for (int64_t pa : M.parents[u]) M.parents_indexing[u].insert(M.parents_indexing[u].end(), M.indexing[pa].begin(), M.indexing[pa].end());
    if (!M.lean)
      for (int64_t ch : M.children[u]) M.children_indexing[u].insert(M.children_indexing[u].end(), M.indexing[ch].begin(), M.indexing[ch].end());
  }
  // na_study :303-313
  M.block_ct_obs.assign(nb, 0);
  for (int i = 0; i < nb; i++)
    for (int64_t r : M.indexing[i])
      if (std::isfinite(M.y[r])) M.block_ct_obs[i]++;
  for (int64_t i = 0; i < M.n_all; i++)
    if (!std::isfinite(M.y[i])) M.y[i] = 0;  // :146
  // XtX :151-155
  M.XtX.assign(M.q, Mat(M.p, M.p));
  for (int j = 0; j < M.q; j++)
    for (int a = 0; a < M.p; a++)
      for (int b = 0; b < M.p; b++) {
        double s = 0;
        for (int64_t r : M.ix_by_q_a[j]) s += M.X_available[r + (size_t)a * M.n] * M.X_available[r + (size_t)b * M.n];
        M.XtX[j](a, b) = s;
      }
  // make_gibbs_groups :194-301 -- checks :201-226
  // block -> group label lookup
  for (int i = 0; i < nb; i++) {
    int u = (int)M.block_names[i] - 1;
    if (M.indexing[u].empty()) continue;
    for (int64_t pa : M.parents[u])
      if (M.block_groups[pa] == M.block_groups[u]) { M.err = "parent in same group"; return false; }
    for (int64_t ch : M.children[u])
      if (M.block_groups[ch] == M.block_groups[u]) { M.err = "child in same group"; return false; }
  }
  std::vector<ivec> temp(M.n_gibbs_groups);
  {
    // label -> group index
    for (int i = 0; i < nb; i++) {
      int u = (int)M.block_names[i] - 1;
      int g = (int)(std::lower_bound(M.block_groups_labels.begin(), M.block_groups_labels.end(), M.block_groups[u]) -
                    M.block_groups_labels.begin());
      if (M.block_ct_obs[u] > 0) temp[g].push_back(u);
    }
  }
  M.n_actual_groups = 0;
  for (int g = 0; g < M.n_gibbs_groups; g++)
    if (!temp[g].empty()) M.n_actual_groups++;
  M.u_by_block_groups.assign(M.n_actual_groups, ivec());
  for (int g = 0; g < M.n_actual_groups; g++) M.u_by_block_groups[g] = temp[g];  // :257-260 (App. D #3)
  M.block_is_reference.assign(nb, 1);
  ivec which_not_reference;
  for (size_t r = 0; r < M.res_is_ref.size(); r++)
    if (M.res_is_ref[r] == 0) which_not_reference.push_back((int64_t)r);
  std::vector<char> in_nonref(nb, 0);
  for (int64_t r : which_not_reference)
    if (r < (int64_t)M.u_by_block_groups.size())
      for (int64_t u : M.u_by_block_groups[r]) in_nonref[u] = 1;
  for (int i = 0; i < nb; i++) {
    int u = (int)M.block_names[i] - 1;
    if (M.block_ct_obs[u] > 0) {
      M.blocks_not_empty.push_back(u);
      if (in_nonref[u]) M.block_is_reference[u] = 0;
    } else {
      M.blocks_predicting.push_back(u);
      M.block_is_reference[u] = 0;
    }
  }
  // init_finalize :355-420
  M.dim_by_parent.assign(nb, ivec());
  for (int i = 0; i < nb; i++) {
    int u = (int)M.block_names[i] - 1;
    if (!M.indexing[u].empty()) {
      M.dim_by_parent[u].assign(M.parents[u].size() + 1, 0);
      for (size_t j = 0; j < M.parents[u].size(); j++)
        M.dim_by_parent[u][j + 1] = M.dim_by_parent[u][j] + (int64_t)M.indexing[M.parents[u][j]].size();
    }
  }
  M.u_is_which_col.assign(nb, {});
  M.this_is_jth_child.assign(nb, ivec());
  for (int i = 0; i < nb; i++) {
    int u = (int)M.block_names[i] - 1;
    M.u_is_which_col[u].resize(M.children[u].size());
    M.this_is_jth_child[u].assign(M.parents[u].size(), 0);
    for (size_t c = 0; c < M.children[u].size(); c++) {
      int child = (int)M.children[u][c];
      size_t which = std::find(M.parents[child].begin(), M.parents[child].end(), (int64_t)u) - M.parents[child].begin();
      if (which >= M.parents[child].size() || M.dim_by_parent[child].empty()) { M.err = "child does not list block as parent"; return false; }
      M.u_is_which_col[u][c] = std::make_pair(M.dim_by_parent[child][which], M.dim_by_parent[child][which + 1]);
    }
  }
  // this_is_jth_child: arma::find(children(up)==u,1,"first") -- done with one pass per parent
  {
    for (int i = 0; i < nb; i++) {
      int u = (int)M.block_names[i] - 1;
      if (M.block_ct_obs[u] == 0) continue;
      for (size_t pp = 0; pp < M.parents[u].size(); pp++) {
        const ivec& ch = M.children[M.parents[u][pp]];
        // children lists are sorted ascending (arma::intersect) -> binary search is equivalent to find-first
        auto it = std::lower_bound(ch.begin(), ch.end(), (int64_t)u);
        if (it == ch.end() || *it != u) {
          it = std::find(ch.begin(), ch.end(), (int64_t)u);
          if (it == ch.end()) { M.err = "block not among its parent's children"; return false; }
        }
        M.this_is_jth_child[u][pp] = (int64_t)(it - ch.begin());
      }
    }
  }
  // init_model_data :422-503
  for (int s = 0; s < 2; s++) {
    Data& d = M.data[s];
    d.theta = theta;
    d.wcore.assign(nb, 0.0);
    d.logdetCi_comps.assign(nb, 0.0);
    d.loglik_w_comps.assign(nb, 0.0);
    d.has_updated.assign(nb, 0);
    d.Kxc.assign(nb, Mat()); d.Kxx_inv.assign(nb, Mat()); d.Kxx_invchol.assign(nb, Mat());
    d.Rcc_invchol.assign(nb, Mat()); d.H.assign(nb, Mat()); d.w_cond_prec.assign(nb, Mat());
    d.w_cond_prec_noref.assign(nb, dvec()); d.ccholprecdiag.assign(nb, dvec()); d.Sigi_chol_noref.assign(nb, dvec());
    d.Sigi_children.assign(nb, Mat()); d.Smu_children.assign(nb, Mat());
    d.AK_uP_all.assign(nb, Mat()); d.AK_uP_u_all.assign(nb, Mat()); d.Sigi_chol.assign(nb, Mat());
    d.probe_Sigi_tot.assign(nb, Mat()); d.probe_Smu_tot.assign(nb, dvec());
    for (int i = 0; i < nb; i++) d.ccholprecdiag[i].assign(M.indexing[i].size(), 0.0);
  }
  if (M.lean) { M.scratch_Kxx_inv.assign(nb, Mat()); M.scratch_Kxx_invchol.assign(nb, Mat()); }
  M.covpars.init(M.q);
  if ((int)theta.size() < M.covpars.npars) { M.err = "theta too short"; return false; }
  return true;
}

static inline std::vector<Mat>& KXI(Model& M, Data& d) { return M.lean ? M.scratch_Kxx_inv : d.Kxx_inv; }
static inline std::vector<Mat>& KXC(Model& M, Data& d) { return M.lean ? M.scratch_Kxx_invchol : d.Kxx_invchol; }

const double hl2pi = -.5 * std::log(2 * M_PI);  // spamtree_model.h:20

// ---------------------------------------------------------------- BUILD
// get_loglik_comps_w_std spamtree_model.cpp:834-998 ; invchol_block_inplace_direct tree_utils.cpp:194-208
static bool build(Model& M, Data& d) {
  M.covpars.transform(d.theta);
  int errtype = -1;
  std::vector<Mat>& Kxx_inv = KXI(M, d);
  std::vector<Mat>& Kxx_invchol = KXC(M, d);
  for (int g = 0; g < M.n_actual_groups; g++) {
    const ivec& us = M.u_by_block_groups[g];
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t ii = 0; ii < (int64_t)us.size(); ii++) {
      const int u = (int)us[ii];
      const ivec& ix = M.indexing[u];
      const int m = (int)ix.size();
      dvec w_x = rows(M.w, ix);
      if (M.parents[u].empty()) {  // :861-878
        Mat Kcc;
        covariancef(Kcc, M, ix.data(), m, ix.data(), m, M.covpars, true);
        if (chol_lower(Kcc)) {
          Mat Li = inv_lower(Kcc);
          Kxx_invchol[u] = Li;
          Kxx_inv[u] = ltl(Li);
          d.Rcc_invchol[u] = Li;
          d.w_cond_prec[u] = Kxx_inv[u];
          dvec t = mv(d.w_cond_prec[u], w_x);
          double s = 0;
          for (int i = 0; i < m; i++) s += w_x[i] * t[i];
          d.wcore[u] = s;
          for (int i = 0; i < m; i++) d.ccholprecdiag[u][i] = Li(i, i);
        } else {
          errtype = 1;
        }
        d.has_updated[u] = 1;
      } else {  // :880-963
        const int last_par = (int)M.parents[u].back();
        const ivec& pix = M.parents_indexing[u];
        const int P = (int)pix.size();
        Mat Kxc_local;
        Mat& Kxc = M.lean ? Kxc_local : d.Kxc[u];
        covariancef(Kxc, M, pix.data(), P, ix.data(), m, M.covpars, false);  // :885
        dvec w_pars = rows(M.w, pix);
        d.H[u] = mtm(Kxc, Kxx_inv[last_par]);  // :887  (m x P)
        const Mat& H = d.H[u];
        {
          dvec t = mv(H, w_pars);
          for (int i = 0; i < m; i++) w_x[i] -= t[i];  // :888
        }
        if (M.res_is_ref[g] == 1) {  // :890-920
          Mat Kcc;
          covariancef(Kcc, M, ix.data(), m, ix.data(), m, M.covpars, true);
          Mat S = mm(H, Kxc);
          for (size_t k = 0; k < S.a.size(); k++) S.a[k] = Kcc.a[k] - S.a[k];
          symmatu(S);
          if (chol_lower(S)) {
            d.Rcc_invchol[u] = inv_lower(S);
            const Mat& Ri = d.Rcc_invchol[u];
            if (!M.children[u].empty()) {
              if (M.limited_tree) {  // :902 inv_sympd(Kcc)
                Mat Kc = Kcc;
                if (chol_lower(Kc)) { Mat Li = inv_lower(Kc); Kxx_inv[u] = ltl(Li); } else errtype = 2;
              } else {  // :904-906
                const Mat& LAi = Kxx_invchol[last_par];
                Mat L(P + m, P + m);
                for (int j = 0; j < P; j++)
                  for (int i = j; i < P; i++) L(i, j) = LAi(i, j);
                Mat RH = mm(Ri, H);
                for (int j = 0; j < P; j++)
                  for (int i = 0; i < m; i++) L(P + i, j) = -RH(i, j);
                for (int j = 0; j < m; j++)
                  for (int i = 0; i < m; i++) L(P + i, P + j) = Ri(i, j);
                Kxx_inv[u] = ltl(L);
                Kxx_invchol[u].a.swap(L.a);
                Kxx_invchol[u].r = Kxx_invchol[u].c = P + m;
              }
              d.has_updated[u] = 1;
            }
            d.w_cond_prec[u] = ltl(Ri);  // :912
            dvec t = mv(d.w_cond_prec[u], w_x);
            double s = 0;
            for (int i = 0; i < m; i++) s += w_x[i] * t[i];
            d.wcore[u] = s;  // :913
            for (int i = 0; i < m; i++) d.ccholprecdiag[u][i] = Ri(i, i);
          } else {
            errtype = 2;
          }
        } else {  // :923-962 non-reference: every row conditionally independent
          d.wcore[u] = 0;
          if ((int)d.w_cond_prec_noref[u].size() != m) d.w_cond_prec_noref[u].assign(m, 0.0);
          for (int r = 0; r < m; r++) {
            Mat Kcc;
            covariancef(Kcc, M, &ix[r], 1, &ix[r], 1, M.covpars, true);
            double hk = 0;
            for (int k = 0; k < P; k++) hk += H(r, k) * Kxc(k, r);
            const double rr = Kcc(0, 0) - hk;
            if (rr > 0.0 && std::isfinite(rr)) {
              const double Rinvchol = 1.0 / std::sqrt(rr);
              d.ccholprecdiag[u][r] = Rinvchol;
              d.w_cond_prec_noref[u][r] = Rinvchol * Rinvchol;
              d.wcore[u] += w_x[r] * d.w_cond_prec_noref[u][r] * w_x[r];
            } else {
              errtype = 3;
            }
          }
        }
      }
      double ld = 0;
      for (int i = 0; i < m; i++) ld += std::log(d.ccholprecdiag[u][i]);
      d.logdetCi_comps[u] = ld;                          // :966
      d.loglik_w_comps[u] = (m + .0) * hl2pi - .5 * d.wcore[u];  // :967-968
    }
    if (errtype > 0) return false;  // :971-982
  }
  double a = 0, b = 0;
  for (int i = 0; i < M.n_blocks; i++) { a += d.logdetCi_comps[i]; b += d.loglik_w_comps[i]; }
  d.logdetCi = a;
  d.loglik_w = a + b;  // :987-988
  return true;
}

// ---------------------------------------------------------------- LLW
// get_loglik_w_std spamtree_model.cpp:781-826
static void loglik_w(Model& M, Data& d) {
  const int64_t nne = (int64_t)M.blocks_not_empty.size();
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < nne; i++) {
    const int u = (int)M.blocks_not_empty[i];
    const int m = (int)M.indexing[u].size();
    dvec w_x = rows(M.w, M.indexing[u]);
    if (!M.parents[u].empty()) {
      dvec t = mv(d.H[u], rows(M.w, M.parents_indexing[u]));
      for (int k = 0; k < m; k++) w_x[k] -= t[k];
    }
    if (M.block_is_reference[u] == 1) {
      dvec t = mv(d.w_cond_prec[u], w_x);
      double s = 0;
      for (int k = 0; k < m; k++) s += w_x[k] * t[k];
      d.wcore[u] = s;
    } else {
      d.wcore[u] = 0;
      for (int k = 0; k < m; k++) d.wcore[u] += w_x[k] * d.w_cond_prec_noref[u][k] * w_x[k];
    }
    d.loglik_w_comps[u] = (m + .0) * hl2pi - .5 * d.wcore[u];
  }
  double a = 0, b = 0;
  for (int i = 0; i < M.n_blocks; i++) { a += d.logdetCi_comps[i]; b += d.loglik_w_comps[i]; }
  d.logdetCi = a;
  d.loglik_w = a + b;
}

// ---------------------------------------------------------------- GIBBS
// gibbs_sample_w_std spamtree_model.cpp:1011-1226.  z = bigrnorm (length n_all)
static bool gibbs_w(Model& M, bool need_update) {
  Data& d = M.param();
  int errtype = -1;
  // spamtree_model.cpp:457-463 allocates these in init_model_data; done lazily (per slot) here, before the parallel loops
  for (int i = 0; i < M.n_blocks; i++) {
    const int C = (int)M.children[i].size(), m = (int)M.indexing[i].size();
    if (C > 0 && d.Sigi_children[i].empty()) { d.Sigi_children[i] = Mat(m, m * C); d.Smu_children[i] = Mat(m, C); }
  }
  for (int g = M.n_actual_groups - 1; g >= 0; g--) {
    const ivec& us = M.u_by_block_groups[g];
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t ii = 0; ii < (int64_t)us.size(); ii++) {
      const int u = (int)us[ii];
      const ivec& ix = M.indexing[u];
      const int m = (int)ix.size();
      const ivec& pix = M.parents_indexing[u];
      const int P = (int)pix.size();
      const int C = (int)M.children[u].size();
      if (M.res_is_ref[g] == 1) {  // :1037-1089
        dvec Smu_tot(m, 0.0);
        Mat Sigi_tot = d.w_cond_prec[u];
        if (P > 0) {  // :1046  AK_uP_all = H' prec  (P x m)
          d.AK_uP_all[u] = mtm(d.H[u], d.w_cond_prec[u]);
        }
        if (C > 0) {  // :1049 sum over slices
          const Mat& SC = d.Sigi_children[u];
          for (int c = 0; c < C; c++)
            for (int j = 0; j < m; j++)
              for (int i = 0; i < m; i++) Sigi_tot(i, j) += SC(i, c * m + j);
        }
        for (int i = 0; i < m; i++) Sigi_tot(i, i) += M.tausq_inv_long[ix[i]];  // :1051
        if (M.probes) d.probe_Sigi_tot[u] = Sigi_tot;
        Mat Sc = Sigi_tot;
        symmatu(Sc);
        if (chol_lower(Sc)) d.Sigi_chol[u] = inv_lower(Sc); else errtype = 10;  // :1054
        if (P > 0) {  // :1063
          dvec t = mtv(d.AK_uP_all[u], rows(M.w, pix));
          for (int i = 0; i < m; i++) Smu_tot[i] += t[i];
        }
        if (C > 0) {  // :1072
          const Mat& SM = d.Smu_children[u];
          for (int c = 0; c < C; c++)
            for (int i = 0; i < m; i++) Smu_tot[i] += SM(i, c);
        }
        for (int i = 0; i < m; i++) Smu_tot[i] += M.tausq_inv_long[ix[i]] * (M.y[ix[i]] - M.XB[ix[i]]);  // :1075-1077
        if (M.probes) d.probe_Smu_tot[u] = Smu_tot;
        if (!d.Sigi_chol[u].empty()) {
          const Mat& Sigi_chol = d.Sigi_chol[u];
          dvec t = mv(Sigi_chol, Smu_tot);
          for (int i = 0; i < m; i++) t[i] += M.bigrnorm[ix[i]];
          dvec wn = mtv(Sigi_chol, t);  // :1086
          for (int i = 0; i < m; i++) M.w[ix[i]] = wn[i];
        }
      } else {  // :1091-1155
        dvec cond_mean_K_wpar = mv(d.H[u], rows(M.w, pix));  // :1103
        if (d.AK_uP_all[u].empty()) d.AK_uP_all[u] = Mat(P, m);
        if ((int)d.Sigi_chol_noref[u].size() != m) d.Sigi_chol_noref[u].assign(m, 0.0);
        if (M.probes) { d.probe_Sigi_tot[u] = Mat(m, 1); d.probe_Smu_tot[u].assign(m, 0.0); }
        for (int r = 0; r < m; r++) {
          const double tsqi = M.tausq_inv_long[ix[r]];
          const double prec = d.w_cond_prec_noref[u][r];
          const double Sigi_tot = prec + tsqi;  // :1123
          const double Smu_tot = prec * cond_mean_K_wpar[r] + tsqi * (M.y[ix[r]] - M.XB[ix[r]]);  // :1125-1127
          if (M.probes) { d.probe_Sigi_tot[u](r, 0) = Sigi_tot; d.probe_Smu_tot[u][r] = Smu_tot; }
          if (Sigi_tot > 0 && std::isfinite(Sigi_tot)) d.Sigi_chol_noref[u][r] = 1.0 / std::sqrt(Sigi_tot); else errtype = 11;
          const double Sc = d.Sigi_chol_noref[u][r];
          M.w[ix[r]] = Sc * Sc * Smu_tot + Sc * M.bigrnorm[ix[r]];  // :1139-1140
          for (int k = 0; k < P; k++) d.AK_uP_all[u](k, r) = d.H[u](r, k) * prec;  // :1144-1147
        }
      }
      // messages to every ancestor :1158-1207
      if (P > 0) {
        Mat AKu_local;
        Mat& AKu = M.lean ? AKu_local : d.AK_uP_u_all[u];
        if (need_update || AKu.empty()) AKu = mm(d.AK_uP_all[u], d.H[u]);  // :1162 (P x P)
        dvec w_par = rows(M.w, pix);  // :1165
        dvec w_u = rows(M.w, ix);
        for (size_t pp = 0; pp < M.parents[u].size(); pp++) {
          const int up = (int)M.parents[u][pp];
          const int c_ix = (int)M.this_is_jth_child[u][pp];
          const int f = (int)M.u_is_which_col[up][c_ix].first, l = (int)M.u_is_which_col[up][c_ix].second;
          const int mu = l - f;
          if (need_update) {  // :1190-1192
            Mat& SC = d.Sigi_children[up];
            for (int j = 0; j < mu; j++)
              for (int i = 0; i < mu; i++) SC(i, c_ix * mu + j) = AKu(f + i, f + j);
          }
          // :1200-1203
          for (int i = 0; i < mu; i++) {
            double s1 = 0;
            for (int k = 0; k < m; k++) s1 += d.AK_uP_all[u](f + i, k) * w_u[k];
            double s2 = 0;
            for (int k = 0; k < f; k++) s2 += AKu(f + i, k) * w_par[k];
            for (int k = l; k < P; k++) s2 += AKu(f + i, k) * w_par[k];
            d.Smu_children[up](i, c_ix) = s1 - s2;
          }
        }
      }
    }
  }
  return errtype <= 0;  // :1215-1217 Rcpp::stop
}

// ---------------------------------------------------------------- PREDICT
// predict_std spamtree_model.cpp:1234-1358 (sampling = true)
static void predict(Model& M, bool theta_update) {
  Data& d = M.param();
  M.covpars.transform(d.theta);
  std::vector<Mat>& Kxx_inv = KXI(M, d);
  std::vector<Mat>& Kxx_invchol = KXC(M, d);
  // the reference rebuilds Kxx_inv(u_par) inside the parallel loop (benign race); do it first, serially
  if (theta_update)
    for (int64_t u : M.blocks_predicting) {
      const int u_par = (int)M.parents[u].back();
      if (d.has_updated[u_par] == 0 && Kxx_inv[u_par].empty()) {  // :1274-1286
        const ivec& ixp = M.indexing[u_par];
        if (M.limited_tree) {
          Mat Kxx;
          covariancef(Kxx, M, ixp.data(), (int)ixp.size(), ixp.data(), (int)ixp.size(), M.covpars, true);
          if (chol_lower(Kxx)) { Mat Li = inv_lower(Kxx); Kxx_inv[u_par] = ltl(Li); }
        } else {
          const int u_gp = (int)M.parents[u_par].back();
          const Mat& LAi = Kxx_invchol[u_gp];
          const Mat& H = d.H[u_par];
          const Mat& Ri = d.Rcc_invchol[u_par];
          const int P = LAi.r, m = Ri.r;
          Mat L(P + m, P + m);
          for (int j = 0; j < P; j++)
            for (int i = j; i < P; i++) L(i, j) = LAi(i, j);
          Mat RH = mm(Ri, H);
          for (int j = 0; j < P; j++)
            for (int i = 0; i < m; i++) L(P + i, j) = -RH(i, j);
          for (int j = 0; j < m; j++)
            for (int i = 0; i < m; i++) L(P + i, P + j) = Ri(i, j);
          Kxx_inv[u_par] = ltl(L);
          Kxx_invchol[u_par] = L;
        }
      }
    }
  const int64_t np = (int64_t)M.blocks_predicting.size();
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t i = 0; i < np; i++) {
    const int u = (int)M.blocks_predicting[i];
    const ivec& ix = M.indexing[u];
    const ivec& pix = M.parents_indexing[u];
    const int m = (int)ix.size(), P = (int)pix.size();
    if (!theta_update && d.H[u].empty() && M.faithful_predict_cache) {  // never computed in this slot: the allocated zeros (:472-473)
      d.Kxc[u] = Mat(P, m);
      d.H[u] = Mat(m, P);
    } else if (theta_update || d.H[u].empty()) {
      covariancef(d.Kxc[u], M, pix.data(), P, ix.data(), m, M.covpars, false);  // :1258
      const int u_par = (int)M.parents[u].back();
      d.H[u] = mtm(d.Kxc[u], Kxx_inv[u_par]);  // :1296
    }
    dvec w_par = rows(M.w, pix);
    dvec hw = mv(d.H[u], w_par);
    for (int r = 0; r < m; r++) {  // :1306-1326
      Mat Kcc;
      covariancef(Kcc, M, &ix[r], 1, &ix[r], 1, M.covpars, true);
      double hk = 0;
      for (int k = 0; k < P; k++) hk += d.H[u](r, k) * d.Kxc[u](k, r);
      const double Ktemp = Kcc(0, 0) - hk;
      const double Rchol = (Ktemp > 0 && std::isfinite(Ktemp)) ? std::sqrt(Ktemp) : 0.0;  // :1316-1322
      M.w[ix[r]] = hw[r] + Rchol * M.bigrnorm[ix[r]];
    }
  }
}

// ---------------------------------------------------------------- beta / tausq
// gibbs_sample_beta spamtree_model.cpp:1364-1391 (zb: p*q normals or NULL -> rng)
static void sample_beta(Model& M, const double* zb) {
  const int p = M.p;
  dvec w_avail;
  if (!M.faithful_beta_index) w_avail = rows(M.w, M.na_ix_all);
  for (int j = 0; j < M.q; j++) {
    Mat Si(p, p);
    for (int a = 0; a < p; a++)
      for (int b = 0; b < p; b++) Si(a, b) = M.tausq_inv[j] * M.XtX[j](a, b) + (a == b ? .01 : 0.0);
    symmatu(Si);
    chol_lower(Si);
    Mat Sc = inv_lower(Si);  // Sigma_chol_Bcoeff
    dvec Xprecy(p, 0.0);
    for (int64_t r : M.ix_by_q_a[j]) {
      // App. D #12: the reference indexes the FULL-length w with positions inside the observed subset
      const double wr = M.faithful_beta_index ? M.w[r] : w_avail[r];
      const double res = M.y_available[r] - wr;
      for (int a = 0; a < p; a++) Xprecy[a] += M.X_available[r + (size_t)a * M.n] * res;
    }
    for (int a = 0; a < p; a++) Xprecy[a] = 0.0 /*Vim*/ + M.tausq_inv[j] * Xprecy[a];
    dvec Bmu = mtv(Sc, mv(Sc, Xprecy));
    dvec z(p);
    for (int a = 0; a < p; a++) z[a] = zb ? zb[a + (size_t)j * p] : M.rng.norm();
    dvec Sz = mtv(Sc, z);
    for (int a = 0; a < p; a++) M.Bcoeff(a, j) = Bmu[a] + Sz[a];
    for (int64_t r : M.ix_by_q[j]) {
      double s = 0;
      for (int a = 0; a < p; a++) s += M.X[r + (size_t)a * M.n_all] * M.Bcoeff(a, j);
      M.XB[r] = s;
    }
  }
}
// gibbs_sample_tausq spamtree_model.cpp:1393-1417
static void sample_tausq(Model& M, const double* fixed /*q values or NULL*/) {
  for (int j = 0; j < M.q; j++) {
    double bcore = 0;
    for (int64_t r : M.ix_by_q_a[j]) {
      const int64_t g = M.na_ix_all[r];
      const double yrr = M.y_available[r] - M.XB[g] - M.w[g];
      bcore += yrr * yrr;
    }
    const double aparam = 2.01 + M.ix_by_q_a[j].size() / 2.0;
    const double bparam = 1.0 / (1.0 + .5 * bcore);
    M.tausq_inv[j] = fixed ? fixed[j] : M.rng.gamma(aparam, bparam);
    for (int64_t r : M.ix_by_q[j]) M.tausq_inv_long[r] = M.tausq_inv[j];
  }
}

// ---------------------------------------------------------------- MH helpers
// mh_adapt.h:150-156, mh_adapt.cpp:3-15, mh_adapt.h:188-202, :230-239
static inline double logistic(double x, double l, double u) { return l + (u - l) / (1.0 + std::exp(-x)); }
static inline double logit(double x, double l, double u) { return -std::log((u - l) / (x - l) - 1.0); }

// RAMAdapt mh_adapt.h:40-135
struct RAMAdapt {
  int p = 0, g0 = 50, c = 0;
  double alpha_star = .234, gamma = 0.5 + 1e-6;
  Mat paramsd, prodparam, S;
  bool started = false, flag_accepted = false;
  double propos_count = 0, accept_count = 0;
  void init(int npars, const Mat& sd) {
    p = npars; S = sd;
    paramsd = S; symmatu(paramsd);  // arma::chol(S,"lower") reads the lower triangle; S is symmetric here
    chol_lower(paramsd);
    prodparam = paramsd;
    for (auto& v : prodparam.a) v /= (g0 + 1.0);
  }
  void adapt(const dvec& U, double alpha, int mc) {
    if (mc < g0) {
      for (int j = 0; j < p; j++)
        for (int i = 0; i < p; i++) prodparam(i, j) += U[i] * U[j] / (mc + 1.0);
    } else {
      if (!started) { paramsd = prodparam; started = true; }
      const int i0 = mc - g0;
      const double eta = std::min(1.0, (p + .0) * std::pow(i0 + 1.0, -gamma));
      alpha = std::min(1.0, alpha);
      double uu = 0;
      for (int i = 0; i < p; i++) uu += U[i] * U[i];
      Mat Sigma(p, p);
      for (int j = 0; j < p; j++)
        for (int i = 0; i < p; i++) Sigma(i, j) = (i == j ? 1.0 : 0.0) + eta * (alpha - alpha_star) * U[i] * U[j] / uu;
      Mat t = mm(paramsd, Sigma);
      Mat St(p, p);  // paramsd * Sigma * paramsd'
      for (int j = 0; j < p; j++)
        for (int i = 0; i < p; i++) {
          double s = 0;
          for (int k = 0; k < p; k++) s += t(i, k) * paramsd(j, k);
          St(i, j) = s;
        }
      S = St;
      Mat L = St;
      // arma::chol on a (numerically) symmetric S; use the lower triangle
      if (chol_lower(L)) paramsd = L;
    }
  }
};

}  // namespace

// ================================================================== C interface (ctypes)
extern "C" {

void* or_create(int64_t n_all, int p, int q, const double* y, const double* X, const double* coords,
                const int64_t* mv_id, int n_blocks, const int64_t* idx_ptr, const int64_t* idx,
                const int64_t* par_ptr, const int64_t* par, const int64_t* chi_ptr, const int64_t* chi,
                const double* block_names, const double* block_groups, const int64_t* res_is_ref, int n_res,
                int limited_tree, const double* theta, int n_theta, const double* beta, double tausq, int flags) {
  Model* M = new Model();
  M->n_all = n_all; M->p = p; M->q = q; M->n_blocks = n_blocks;
  M->y.assign(y, y + n_all);
  M->X.assign(X, X + (size_t)n_all * p);
  M->coords.assign(coords, coords + (size_t)n_all * 2);
  M->mv_id.assign(mv_id, mv_id + n_all);
  M->res_is_ref.assign(res_is_ref, res_is_ref + n_res);
  M->parents.resize(n_blocks); M->children.resize(n_blocks); M->indexing.resize(n_blocks);
  for (int i = 0; i < n_blocks; i++) {
    M->indexing[i].assign(idx + idx_ptr[i], idx + idx_ptr[i + 1]);
    M->parents[i].assign(par + par_ptr[i], par + par_ptr[i + 1]);
    M->children[i].assign(chi + chi_ptr[i], chi + chi_ptr[i + 1]);
  }
  M->limited_tree = limited_tree != 0;
  M->block_names.assign(block_names, block_names + n_blocks);
  M->block_groups.assign(block_groups, block_groups + n_blocks);
  M->lean = (flags & 1) != 0;
  M->q1_norm_expansion = (flags & 2) != 0;
  M->faithful_beta_index = (flags & 4) == 0;
  M->probes = (flags & 8) != 0;
  M->faithful_predict_cache = (flags & 16) == 0;
  dvec th(theta, theta + n_theta), be(beta, beta + p);
  if (!model_init(*M, th, be, 1.0 / tausq)) {
    fprintf(stderr, "oracle: %s\n", M->err.c_str());
    delete M;
    return nullptr;
  }
  M->rng.seed(1);
  return M;
}
void or_destroy(void* h) { delete (Model*)h; }
void or_seed(void* h, uint64_t s) { ((Model*)h)->rng.seed(s); }
// switches the dense kernels to the OpenBLAS at `path` (single-threaded per call: the parallelism stays OpenMP over the
// blocks of a level, like the reference's); returns 0, or 1 when the library or a symbol is missing.  path == NULL: off.
int or_use_blas(const char* path) {
  g_blas.on = false;
  if (!path) return 0;
  void* h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
  if (!h) return 1;
  BlasApi b;
  b.dgemm = (decltype(b.dgemm))dlsym(h, "scipy_dgemm_");
  b.dsyrk = (decltype(b.dsyrk))dlsym(h, "scipy_dsyrk_");
  b.dgemv = (decltype(b.dgemv))dlsym(h, "scipy_dgemv_");
  b.dpotrf = (decltype(b.dpotrf))dlsym(h, "scipy_dpotrf_");
  b.dtrtri = (decltype(b.dtrtri))dlsym(h, "scipy_dtrtri_");
  b.set_threads = (decltype(b.set_threads))dlsym(h, "scipy_openblas_set_num_threads");
  if (!b.dgemm || !b.dsyrk || !b.dgemv || !b.dpotrf || !b.dtrtri) return 1;
  if (b.set_threads) b.set_threads(1);
  b.on = true;
  g_blas = b;
  return 0;
}
void or_set_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n);
#endif
}
int or_max_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
// slot: 0 = param_data, 1 = alter_data
void or_theta_update(void* h, int slot, const double* theta) {
  Model& M = *(Model*)h;
  Data& d = slot ? M.alter() : M.param();
  std::copy(theta, theta + d.theta.size(), d.theta.begin());
}
int or_build(void* h, int slot, double* out3) {
  Model& M = *(Model*)h;
  Data& d = slot ? M.alter() : M.param();
  bool ok = build(M, d);
  out3[0] = d.loglik_w; out3[1] = d.logdetCi; out3[2] = ok ? 1 : 0;
  return ok ? 1 : 0;
}
void or_loglik_w(void* h, int slot, double* out2) {
  Model& M = *(Model*)h;
  Data& d = slot ? M.alter() : M.param();
  loglik_w(M, d);
  out2[0] = d.loglik_w; out2[1] = d.logdetCi;
}
int or_gibbs(void* h, const double* z) {
  Model& M = *(Model*)h;
  M.bigrnorm.resize(M.n_all);
  if (z) std::copy(z, z + M.n_all, M.bigrnorm.begin());
  else for (int64_t i = 0; i < M.n_all; i++) M.bigrnorm[i] = M.rng.norm();
  return gibbs_w(M, true) ? 1 : 0;
}
void or_swap(void* h) { Model& M = *(Model*)h; M.cur = 1 - M.cur; }
void or_predict(void* h, int theta_changed) { predict(*(Model*)h, theta_changed != 0); }
void or_sample_beta(void* h, const double* zb) { sample_beta(*(Model*)h, zb); }
void or_sample_tausq(void* h, const double* fixed) { sample_tausq(*(Model*)h, fixed); }
void or_get_w(void* h, double* out) { Model& M = *(Model*)h; std::copy(M.w.begin(), M.w.end(), out); }
void or_set_w(void* h, const double* in) { Model& M = *(Model*)h; std::copy(in, in + M.n_all, M.w.begin()); }
void or_get_params(void* h, double* beta_pq, double* tausq_inv_q, double* xb_n) {
  Model& M = *(Model*)h;
  if (beta_pq) std::copy(M.Bcoeff.a.begin(), M.Bcoeff.a.end(), beta_pq);
  if (tausq_inv_q) std::copy(M.tausq_inv.begin(), M.tausq_inv.end(), tausq_inv_q);
  if (xb_n) std::copy(M.XB.begin(), M.XB.end(), xb_n);
}
void or_set_tausq_inv(void* h, const double* t) {
  Model& M = *(Model*)h;
  for (int j = 0; j < M.q; j++) { M.tausq_inv[j] = t[j]; for (int64_t r : M.ix_by_q[j]) M.tausq_inv_long[r] = t[j]; }
}

static int64_t put(const dvec& v, double* out, int64_t cap) {
  if (out) for (int64_t i = 0; i < (int64_t)v.size() && i < cap; i++) out[i] = v[i];
  return (int64_t)v.size();
}
static int64_t puti(const ivec& v, double* out, int64_t cap) {
  if (out) for (int64_t i = 0; i < (int64_t)v.size() && i < cap; i++) out[i] = (double)v[i];
  return (int64_t)v.size();
}
// Generic probe: returns the element count; writes at most cap doubles (ints are converted exactly).
int64_t or_get(void* h, const char* name, int slot, int u, int c, double* out, int64_t cap) {
  Model& M = *(Model*)h;
  Data& d = slot ? M.alter() : M.param();
  std::string s(name);
  if (s == "H") return put(d.H[u].a, out, cap);
  if (s == "Ri") return put(d.Rcc_invchol[u].a, out, cap);
  if (s == "prec") return put(d.w_cond_prec[u].a, out, cap);
  if (s == "prec_noref") return put(d.w_cond_prec_noref[u], out, cap);
  if (s == "ccholprecdiag") return put(d.ccholprecdiag[u], out, cap);
  if (s == "Kxx_inv") return put(KXI(M, d)[u].a, out, cap);
  if (s == "Kxx_invchol") return put(KXC(M, d)[u].a, out, cap);
  if (s == "Kxc") return put(d.Kxc[u].a, out, cap);
  if (s == "Sigi_tot") return put(d.probe_Sigi_tot[u].a, out, cap);
  if (s == "Smu_tot") return put(d.probe_Smu_tot[u], out, cap);
  if (s == "Sigi_chol") return put(d.Sigi_chol[u].a, out, cap);
  if (s == "Sigi_chol_noref") return put(d.Sigi_chol_noref[u], out, cap);
  if (s == "Sigi_children") return put(d.Sigi_children[u].a, out, cap);
  if (s == "Smu_children") return put(d.Smu_children[u].a, out, cap);
  if (s == "logdetCi_comps") return put(d.logdetCi_comps, out, cap);
  if (s == "loglik_w_comps") return put(d.loglik_w_comps, out, cap);
  if (s == "wcore") return put(d.wcore, out, cap);
  if (s == "theta") return put(d.theta, out, cap);
  if (s == "parents_indexing") return puti(M.parents_indexing[u], out, cap);
  if (s == "children_indexing") return puti(M.children_indexing[u], out, cap);
  if (s == "dim_by_parent") return puti(M.dim_by_parent[u], out, cap);
  if (s == "this_is_jth_child") return puti(M.this_is_jth_child[u], out, cap);
  if (s == "u_by_block_groups") return puti(M.u_by_block_groups[u], out, cap);
  if (s == "blocks_not_empty") return puti(M.blocks_not_empty, out, cap);
  if (s == "blocks_predicting") return puti(M.blocks_predicting, out, cap);
  if (s == "block_is_reference") return puti(M.block_is_reference, out, cap);
  if (s == "block_ct_obs") return puti(M.block_ct_obs, out, cap);
  if (s == "n_actual_groups") { if (out && cap > 0) out[0] = M.n_actual_groups; return 1; }
  if (s == "u_is_which_col") {  // expanded like spamtree_model.cpp:401-408; c = child index; slot: 0 local, 1 other
    const auto& fl = M.u_is_which_col[u][c];
    const int64_t dimen = (int64_t)M.parents_indexing[M.children[u][c]].size();
    ivec r;
    for (int64_t k = 0; k < dimen; k++) {
      bool local = (k >= fl.first && k < fl.second);
      if ((slot == 0) == local) r.push_back(k);
    }
    return puti(r, out, cap);
  }
  return -1;
}

// ---- standalone R-exported helpers
// CrossCovarianceAG10 covariance_functions.cpp:301-355 (mv ids 1-based; Dmat q x q col-major)
void or_cross_covariance_ag10(const double* c1, const int64_t* mv1, int64_t n1, const double* c2, const int64_t* mv2,
                              int64_t n2, const double* ai1, const double* ai2, const double* phi_i,
                              const double* thetamv, int n_thetamv, const double* Dmat, int q, double* out) {
  dvec tm(thetamv, thetamv + n_thetamv);
  for (int64_t i = 0; i < n1; i++) {
    const int vi = (int)mv1[i] - 1;
    const double ai1_sq = ai1[vi] * ai1[vi], ai2_sq = ai2[vi] * ai2[vi];
    for (int64_t j = 0; j < n2; j++) {
      const double dx = c1[i] - c2[j], dy = c1[i + n1] - c2[j + n2];
      const double hh = std::sqrt(dx * dx + dy * dy);
      const int vj = (int)mv2[j] - 1;
      const double v = Dmat[vi + (size_t)vj * q];
      out[i + (size_t)j * n1] = (v == 0) ? ai1_sq * C_base(hh, 0, tm, q) + ai2_sq * fphi(hh, phi_i[vi])
                                         : ai1[vi] * ai1[vj] * C_base(hh, v, tm, q);
    }
  }
}
// kthresholds tree_dep.cpp:16-27 (x is copied: nth_element permutes it across calls exactly as there)
void or_kthresholds(const double* x, int64_t n, int k, double* res) {
  dvec xx(x, x + n);
  for (unsigned int i = 1; i < (unsigned)k; i++) {
    unsigned int Q1 = (unsigned int)(i * (unsigned int)n / k);
    std::nth_element(xx.begin(), xx.begin() + Q1, xx.end());
    res[i - 1] = xx[Q1];
  }
}
// part_axis_parallel_lmt / column_threshold tree_dep.cpp:42-67 ; thresholds for axis j at thr[thr_ptr[j]..thr_ptr[j+1])
void or_part_axis_parallel_lmt(const double* coords, int64_t n, int d, const double* thr, const int64_t* thr_ptr,
                               double* out) {
  for (int j = 0; j < d; j++)
    for (int64_t i = 0; i < n; i++) {
      int over = 1;
      for (int64_t t = thr_ptr[j]; t < thr_ptr[j + 1]; t++)
        if (coords[i + (size_t)j * n] >= thr[t]) over += 1;
      out[i + (size_t)j * n] = over;
    }
}
// number_revalue tree_dep.cpp:240-259
void or_number_revalue(const int64_t* orig, int64_t nr, int nc, const int64_t* from_val, const int64_t* to_val,
                       int64_t nfrom, int64_t* out) {
  int64_t maxval = 0;
  for (int64_t j = 0; j < nfrom; j++) maxval = std::max(maxval, to_val[j]);
  for (int64_t i = 0; i < nr; i++)
    for (int c = 0; c < nc; c++) {
      int64_t v = orig[i + (size_t)c * nr];
      out[i + (size_t)c * nr] = v;
      for (int64_t j = 0; j < nfrom; j++)
        if (v == from_val[j]) { out[i + (size_t)c * nr] = to_val[j]; break; }
      if (out[i + (size_t)c * nr] > maxval) out[i + (size_t)c * nr] = 0;
    }
}
// make_edges tree_dep.cpp:75-130 and make_edges_limited :133-186.  parchimat: nr x L col-major doubles, NaN = NA.
// Output CSR is written into caller buffers sized by a first call with NULL outputs (returns total counts).
static ivec unique_finite(const dvec& v) {
  ivec o;
  for (double x : v) if (std::isfinite(x)) o.push_back((int64_t)x);
  std::sort(o.begin(), o.end());
  o.erase(std::unique(o.begin(), o.end()), o.end());
  return o;
}
void or_make_edges(const double* parchimat, int64_t nr, int L, const int64_t* non_empty_blocks, int64_t n_ne,
                   const int64_t* res_is_ref, int limited, int64_t* par_ptr, int64_t* par_idx, int64_t* chi_ptr,
                   int64_t* chi_idx, int64_t* n_blocks_out) {
  int64_t n_blocks = 0;
  for (int64_t i = 0; i < nr; i++) {
    double v = parchimat[i + (size_t)(L - 1) * nr];
    if (std::isfinite(v)) n_blocks = std::max(n_blocks, (int64_t)v);
  }
  *n_blocks_out = n_blocks;
  std::vector<ivec> parents(n_blocks), children(n_blocks);
  ivec reference_res;
  for (int l = 0; l < L; l++) if (res_is_ref[l] == 1) reference_res.push_back(l);
  ivec ne(non_empty_blocks, non_empty_blocks + n_ne);
  for (auto& v : ne) v -= 1;
  std::sort(ne.begin(), ne.end());
  for (int lev = 0; lev < L; lev++) {
    dvec col(parchimat + (size_t)lev * nr, parchimat + (size_t)(lev + 1) * nr);
    ivec blocks_this_lev = unique_finite(col);
    for (size_t b = 0; b < blocks_this_lev.size(); b++) {
      const int64_t u = blocks_this_lev[b] - 1;
      ivec rowsel;
      for (int64_t i = 0; i < nr; i++) if (col[i] == (double)blocks_this_lev[b]) rowsel.push_back(i);
      if (res_is_ref[lev] == 1 && lev < L - 1) {
        dvec vals;
        const int lastc = limited ? lev + 1 : L - 1;
        for (int c = lev + 1; c <= lastc; c++)
          for (int64_t i : rowsel) vals.push_back(parchimat[i + (size_t)c * nr]);
        ivec pc = unique_finite(vals);
        for (auto& v : pc) v -= 1;
        ivec inter;
        std::set_intersection(pc.begin(), pc.end(), ne.begin(), ne.end(), std::back_inserter(inter));
        children[u] = inter;
      }
      if (lev > 0) {
        ivec colselect;
        if (!reference_res.empty()) { for (int64_t r : reference_res) if (r < lev) colselect.push_back(r); }
        else for (int c = 0; c < lev; c++) colselect.push_back(c);
        if (limited) { int64_t lastcol = colselect.back(); colselect.assign(1, lastcol); }
        dvec vals;
        for (int64_t c : colselect)
          for (int64_t i : rowsel) vals.push_back(parchimat[i + (size_t)c * nr]);
        ivec pp = unique_finite(vals);
        for (auto& v : pp) v -= 1;
        parents[u] = pp;
      }
    }
  }
  int64_t np = 0, nc = 0;
  for (int64_t i = 0; i < n_blocks; i++) {
    if (par_ptr) par_ptr[i] = np;
    if (chi_ptr) chi_ptr[i] = nc;
    for (int64_t v : parents[i]) { if (par_idx) par_idx[np] = v; np++; }
    for (int64_t v : children[i]) { if (chi_idx) chi_idx[nc] = v; nc++; }
  }
  if (par_ptr) par_ptr[n_blocks] = np;
  if (chi_ptr) chi_ptr[n_blocks] = nc;
  n_blocks_out[1] = np;
  n_blocks_out[2] = nc;
}

// ---- the MCMC loop: spamtree_mv_mcmc spamtree_fit.cpp:5-430 (printing / interrupt handling omitted)
// bounds: npar x 2 col-major; mcmcsd: npar x npar col-major; outputs sized by the caller:
// beta_mcmc p*keep*q (cube p x keep x q), tausq_mcmc q*keep, theta_mcmc npar*keep, w_mcmc/yhat_mcmc n_all*keep (or NULL)
// stats: [0] accepted count, [1] chol-fail (unacceptable) count, [2] seconds in the loop
int or_mcmc(void* h, const double* bounds, const double* mcmcsd, int keep, int burn, int thin, int adapting,
            int sample_beta_f, int sample_tausq_f, int sample_theta_f, int sample_w_f, int sample_predicts_f,
            uint64_t seed, double* beta_mcmc, double* tausq_mcmc, double* theta_mcmc, double* w_mcmc,
            double* yhat_mcmc, double* paramsd_out, double* stats) {
  Model& M = *(Model*)h;
  M.rng.seed(seed);
  double o3[3];
  or_build(h, 0, o3);  // :110-111
  or_build(h, 1, o3);
  const int npar = (int)M.param().theta.size();
  dvec param = M.param().theta, predict_param = param;
  double current_loglik = M.param().loglik_w;
  const int mcmc = thin * keep + burn;
  Mat sd(npar, npar);
  std::copy(mcmcsd, mcmcsd + (size_t)npar * npar, sd.a.begin());
  RAMAdapt ad;
  ad.init(npar, sd);
  int msaved = 0;
  double nacc = 0, nfail = 0;
  double t0 = 0;
#ifdef _OPENMP
  t0 = omp_get_wtime();
#endif
  auto lo = [&](int j) { return bounds[j]; };
  auto hi = [&](int j) { return bounds[j + npar]; };
  for (int m = 0; m < mcmc; m++) {
    bool predicting = false;
    const int mx = m - burn;
    if (mx >= 0 && mx % thin == 0) predicting = true;
    if (sample_w_f) {  // :183-187
      if (!or_gibbs(h, nullptr)) return -10;
      loglik_w(M, M.param());
      current_loglik = M.param().loglik_w;
    }
    if (sample_theta_f) {  // :203-289
      ad.propos_count++; ad.c++; ad.flag_accepted = false;
      dvec U(npar);
      for (int j = 0; j < npar; j++) U[j] = M.rng.norm();
      dvec new_param(npar);
      for (int j = 0; j < npar; j++) {
        double s = logit(param[j], lo(j), hi(j));
        for (int k = 0; k < npar; k++) s += ad.paramsd(j, k) * U[k];
        new_param[j] = logistic(s, lo(j), hi(j));
      }
      bool out_unif_bounds = false;  // unif_bounds mh_adapt.h:188-202
      for (int j = 0; j < npar; j++) {
        if (new_param[j] < lo(j)) { out_unif_bounds = true; new_param[j] = lo(j) + 1e-10; }
        if (new_param[j] > hi(j)) { out_unif_bounds = true; new_param[j] = hi(j) - 1e-10; }
      }
      (void)out_unif_bounds;
      M.alter().theta = new_param;
      const bool acceptable = build(M, M.alter());
      const double new_loglik = M.alter().loglik_w;
      current_loglik = M.param().loglik_w;
      if (std::isnan(current_loglik)) return -1;
      double jac = 0;  // calc_jacobian mh_adapt.h:230-239
      for (int j = 0; j < npar; j++)
        jac += (-std::log(hi(j) - param[j]) - std::log(param[j] - lo(j))) -
               (-std::log(hi(j) - new_param[j]) - std::log(new_param[j] - lo(j)));
      const double logaccept = new_loglik - current_loglik + jac;
      // do_I_accept mh_adapt.h:20-36
      double acceptj = 1.0;
      if (!std::isfinite(logaccept)) acceptj = 0.0; else if (logaccept < 0) acceptj = std::exp(logaccept);
      const double uu = M.rng.unif();
      const bool accepted = (uu < acceptj) && acceptable;
      if (!acceptable) nfail++;
      if (accepted) {
        nacc++; ad.accept_count++; ad.flag_accepted = true;
        current_loglik = new_loglik;
        M.cur = 1 - M.cur;  // accept_make_change
        param = new_param;
      }
      if (adapting) ad.adapt(U, (acceptable ? 1.0 : 0.0) * std::exp(logaccept), m);  // :285
    }
    bool need_update = false;  // :300
    for (int j = 0; j < npar; j++) if (std::fabs(param[j] - predict_param[j]) > 1e-05) need_update = true;
    if (predicting && sample_predicts_f && sample_w_f) { predict(M, need_update); predict_param = param; }
    if (sample_tausq_f) sample_tausq(M, nullptr);
    if (sample_beta_f) sample_beta(M, nullptr);
    if (mx >= 0 && mx % thin == 0) {  // :376-389
      for (int j = 0; j < M.q; j++) tausq_mcmc[j + (size_t)msaved * M.q] = 1.0 / M.tausq_inv[j];
      for (int j = 0; j < M.q; j++)
        for (int a = 0; a < M.p; a++) beta_mcmc[a + (size_t)msaved * M.p + (size_t)j * M.p * keep] = M.Bcoeff(a, j);
      for (int j = 0; j < npar; j++) theta_mcmc[j + (size_t)msaved * npar] = M.param().theta[j];
      for (int64_t i = 0; i < M.n_all; i++) {
        const double e = M.rng.norm();
        if (w_mcmc) w_mcmc[i + (size_t)msaved * M.n_all] = M.w[i];
        if (yhat_mcmc) yhat_mcmc[i + (size_t)msaved * M.n_all] = M.XB[i] + M.w[i] + std::pow(M.tausq_inv_long[i], -.5) * e;
      }
      msaved++;
    }
  }
  if (paramsd_out) std::copy(ad.paramsd.a.begin(), ad.paramsd.a.end(), paramsd_out);
  if (stats) {
    stats[0] = nacc; stats[1] = nfail;
#ifdef _OPENMP
    stats[2] = omp_get_wtime() - t0;
#else
    stats[2] = 0;
#endif
  }
  return 0;
}

// one timed iteration for the CPU baseline: GIBBS + LLW + BUILD(alter) [+ swap] + tausq + beta, no predict
// (spamtree_fit.cpp:167-330 without :302-306); theta_prop supplied by the caller
double or_timed_iteration(void* h, const double* theta_prop, int do_swap) {
  Model& M = *(Model*)h;
  double t0 = 0, t1 = 0;
#ifdef _OPENMP
  t0 = omp_get_wtime();
#endif
  or_gibbs(h, nullptr);
  loglik_w(M, M.param());
  std::copy(theta_prop, theta_prop + M.alter().theta.size(), M.alter().theta.begin());
  bool ok = build(M, M.alter());
  if (ok && do_swap) M.cur = 1 - M.cur;
  sample_tausq(M, nullptr);
  sample_beta(M, nullptr);
#ifdef _OPENMP
  t1 = omp_get_wtime();
#endif
  return t1 - t0;
}

}  // extern "C"
