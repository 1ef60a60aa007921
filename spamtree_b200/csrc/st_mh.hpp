// st_mh.hpp — the Metropolis-Hastings glue of the reference (mh_adapt.h:20-36, :40-135, :150-156, :188-202, :230-239;
// mh_adapt.cpp:3-15; spamtree_fit.cpp:203-289) on plain arrays, compiled for the host AND the device: the host path
// (rng_mode 0, lock-step with the oracle / the reference's own driver) and the device-resident chain (rng_mode 1, one
// thread of mh_step_kernel) run the same code.  Column-major npar x npar matrices, npar <= kMaxPar.
#pragma once
#include <cmath>

#ifdef __CUDACC__
#define ST_HD __host__ __device__
#else
#define ST_HD
#endif

namespace st {

constexpr int kMaxPar = 64;  // q = 8: 3q + 3 + q(q-1)/2 = 55

// mh_adapt.h:150-156
ST_HD inline double mh_logistic(double x, double l, double u) { return l + (u - l) / (1.0 + exp(-x)); }
ST_HD inline double mh_logit(double x, double l, double u) { return -log((u - l) / (x - l) - 1.0); }

// lower Cholesky in place, column-major n x n with leading dimension n (reads the lower triangle); false if not
// positive definite (arma::chol(S, "lower") throws there, mh_adapt.h:86,133).  The strict upper triangle is zeroed.
ST_HD inline bool mh_chol_lower(double* A, int n) {
  for (int j = 0; j < n; j++) {
    double d = A[j + j * n];
    for (int k = 0; k < j; k++) d -= A[j + k * n] * A[j + k * n];
    if (!(d > 0.0) || !isfinite(d)) return false;
    d = sqrt(d);
    A[j + j * n] = d;
    for (int i = j + 1; i < n; i++) {
      double s = A[i + j * n];
      for (int k = 0; k < j; k++) s -= A[i + k * n] * A[j + k * n];
      A[i + j * n] = s / d;
    }
  }
  for (int j = 0; j < n; j++)
    for (int i = 0; i < j; i++) A[i + j * n] = 0.0;
  return true;
}

// RAMAdapt::RAMAdapt (mh_adapt.h:78-96): paramsd = chol(metropolis_sd), prodparam = paramsd / (g0 + 1)
ST_HD inline bool ram_init(int p, const double* metropolis_sd, double* paramsd, double* prodparam, int g0 = 50) {
  for (int e = 0; e < p * p; e++) paramsd[e] = metropolis_sd[e];
  if (!mh_chol_lower(paramsd, p)) return false;
  for (int e = 0; e < p * p; e++) prodparam[e] = paramsd[e] / (g0 + 1.0);
  return true;
}

// RAMAdapt::adapt (mh_adapt.h:117-135), Vihola's robust adaptive Metropolis.  scratch: 2 p^2 doubles.
// Returns false when chol(S) fails: the reference's arma::chol throws and the run ends with its "None" list
// (spamtree_fit.cpp:416-427); here the previous factor is kept and the caller decides (st_mcmc_run counts it).
ST_HD inline bool ram_adapt(int p, double* paramsd, double* prodparam, int* started, const double* U, double alpha, int mc,
                            double* scratch, int g0 = 50, double alpha_star = .234, double gamma = 0.5 + 1e-6) {
  if (mc < g0) {
    for (int j = 0; j < p; j++)
      for (int i = 0; i < p; i++) prodparam[i + j * p] += U[i] * U[j] / (mc + 1.0);
    return true;
  }
  if (!*started) {
    for (int e = 0; e < p * p; e++) paramsd[e] = prodparam[e];
    *started = 1;
  }
  const int i0 = mc - g0;
  const double eta = fmin(1.0, (p + .0) * pow(i0 + 1.0, -gamma));
  alpha = fmin(1.0, alpha);
  double uu = 0;
  for (int i = 0; i < p; i++) uu += U[i] * U[i];
  double* T = scratch;          // paramsd * Sigma, Sigma = I + eta (alpha - alpha*) U U' / |U|^2
  double* S = scratch + p * p;  // T * paramsd'
  const double f = eta * (alpha - alpha_star);
  for (int j = 0; j < p; j++)
    for (int i = 0; i < p; i++) T[i + j * p] = 0.0;
  for (int j = 0; j < p; j++)
    for (int k = 0; k < p; k++) {  // accumulation order of a column-major axpy product
      const double b = (k == j ? 1.0 : 0.0) + f * U[k] * U[j] / uu;
      for (int i = 0; i < p; i++) T[i + j * p] += paramsd[i + k * p] * b;
    }
  for (int j = 0; j < p; j++)
    for (int i = 0; i < p; i++) {
      double s = 0;
      for (int k = 0; k < p; k++) s += T[i + k * p] * paramsd[j + k * p];
      S[i + j * p] = s;
    }
  if (!mh_chol_lower(S, p)) return false;
  for (int e = 0; e < p * p; e++) paramsd[e] = S[e];
  return true;
}

// spamtree_fit.cpp:211-215: new = back(fwd(param) + paramsd U), then unif_bounds (mh_adapt.h:188-202).
// bounds: npar x 2 column-major.  Returns out_unif_bounds.
ST_HD inline bool mh_propose(int p, const double* param, const double* bounds, const double* paramsd, const double* U, double* new_param) {
  bool oob = false;
  for (int j = 0; j < p; j++) {
    const double lo = bounds[j], hi = bounds[j + p];
    double s = mh_logit(param[j], lo, hi);
    double t = 0;
    for (int k = 0; k < p; k++) t += paramsd[j + k * p] * U[k];
    s += t;
    double v = mh_logistic(s, lo, hi);
    if (v < lo) { oob = true; v = lo + 1e-10; }
    if (v > hi) { oob = true; v = hi - 1e-10; }
    new_param[j] = v;
  }
  return oob;
}

// calc_jacobian (mh_adapt.h:230-239) with normal_proposal_logitscale (:210-213)
ST_HD inline double mh_jacobian(int p, const double* new_param, const double* param, const double* bounds) {
  double jac = 0;
  for (int j = 0; j < p; j++) {
    const double lo = bounds[j], hi = bounds[j + p];
    jac += (-log(hi - param[j]) - log(param[j] - lo)) - (-log(hi - new_param[j]) - log(new_param[j] - lo));
  }
  return jac;
}

// do_I_accept (mh_adapt.h:20-36): the acceptance probability; the caller compares its uniform draw with it (u < p)
ST_HD inline double mh_accept_prob(double logaccept) {
  if (!isfinite(logaccept)) return 0.0;
  return logaccept < 0 ? exp(logaccept) : 1.0;
}

}  // namespace st
