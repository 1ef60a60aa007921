// st_chain.hpp — the chain state that lives in device memory.  The reference keeps it in the host objects of its driver
// (spamtree_fit.cpp:118-166: param, predict_param, current_loglik, adaptivemc ...) and in SpamTreeMV's public fields
// (param_data / alter_data, tausq_inv, Bcoeff); here the kernels read it from HBM, so that an MCMC iteration — proposal,
// accept decision, slot swap, RAM adaptation, tausq and beta draws — needs no host round trip.
// The host-driven path (rng_mode 0, lock-step with the oracle) keeps `cur`, the thetas and the covariance tables of
// this struct in sync with its own copies; the device-resident path (rng_mode 1) owns all of it.
#pragma once
#include "st_mh.hpp"

namespace st {

constexpr int kMaxQ = 8;
constexpr int kMaxStats = 40;  // q * (p + 1)

// theta -> per outcome-pair coefficients: K(h; i, j) = c1*exp(-r1*h) + c2*exp(-r2*h)
// (covariance_functions.cpp:113-135, :213-286; q == 1 -> cexpcov :95-111 with direct-difference distance)
struct CovTab {
  int q;
  double c1[kMaxQ * kMaxQ], r1[kMaxQ * kMaxQ], c2[kMaxQ * kMaxQ], r2[kMaxQ * kMaxQ];
};

// covariance_functions.cpp:34-92 (theta layout) folded with :113-135 (C_base) into per outcome-pair coefficients.
// Returns 0, or 1: q outside 1..8, 2: theta has the wrong length for q.
ST_HD inline int make_covtab_hd(const double* theta, int n_theta, int q, CovTab& tab) {
  if (q < 1 || q > kMaxQ) return 1;
  const int n_cbase = q > 2 ? 3 : 1, npars = 3 * q + n_cbase, kd = n_theta - npars;
  if (kd < 0 || (q >= 2 && kd != q * (q - 1) / 2) || (q == 1 && kd != 0)) return 2;
  const double *ai1 = theta, *ai2 = theta + q, *phi_i = theta + 2 * q, *thetamv = theta + 3 * q;
  tab.q = q;
  if (q == 1) {  // cexpcov(…, sigmasq = ai1(0), phi = thetamv(0)) (:220-221)
    tab.c1[0] = ai1[0]; tab.r1[0] = thetamv[0]; tab.c2[0] = 0; tab.r2[0] = 0;
    return 0;
  }
  double D[kMaxQ * kMaxQ];  // vec_to_symmat: column-major fill of the strict lower triangle
  for (int e = 0; e < kMaxQ * kMaxQ; e++) D[e] = 0.0;
  int ix = 0;
  for (int j = 0; j < q; j++)
    for (int i = j + 1; i < q; i++) { D[i * q + j] = D[j * q + i] = theta[npars + ix]; ix++; }
  for (int i = 0; i < q; i++)
    for (int j = 0; j < q; j++) {
      const double v = D[i * q + j];
      const int e = i * q + j;
      double psi_sqrt, psi2, c;
      if (q > 2) {
        psi_sqrt = exp(0.5 * thetamv[1] * log1p(thetamv[0] * (v == 0 ? 0.0 : v)));  // sqrt_fpsi
        psi2 = psi_sqrt * psi_sqrt;
        c = thetamv[2];
      } else {
        psi_sqrt = sqrt((v == 0 ? 0.0 : v) + 1);
        psi2 = (v == 0 ? 0.0 : v) + 1.0;
        c = thetamv[0];
      }
      if (v == 0) {  // "same outcome" is detected by Dmat(i,j) == 0 (:250)
        tab.c1[e] = ai1[i] * ai1[i] / psi2; tab.r1[e] = c / psi_sqrt;
        tab.c2[e] = ai2[i] * ai2[i];        tab.r2[e] = phi_i[i];
      } else {
        tab.c1[e] = ai1[i] * ai1[j] / psi2; tab.r1[e] = c / psi_sqrt;
        tab.c2[e] = 0;                      tab.r2[e] = 0;
      }
    }
  return 0;
}

struct ChainDev {
  // ---- which theta-slot is param_data (tree_utils.h:63-102; accept_make_change, spamtree_model.cpp:1432-1435)
  int cur;
  int iter;          // MCMC iteration m of the device-resident chain (spamtree_fit.cpp:167)
  int npar, p, q, adapting;
  int accepted_now;  // 1: the last accept step took the proposal (the deferred half of BUILD and the Gram refresh key off it)
  int ram_started;   // RAMAdapt::started (mh_adapt.h:121-124)
  int predict_build; // 1: theta moved since the last predict (need_update, spamtree_fit.cpp:300)
  int gibbs_fail;    // sticky: a conditional precision was not positive definite (Rcpp::stop in the reference, :1215-1217)
  int nan_loglik;    // sticky: spamtree_fit.cpp:234-237
  int msaved;
  long long n_accepted, n_chol_fail, n_ram_fail;
  unsigned long long seed;
  // ---- per physical slot
  double loglik[2], logdet[2];
  double theta[2][kMaxPar];
  CovTab tab[2];
  // ---- log-density reductions: [0..2] replicated blocks (or all of them), [4..6] the rank's own blocks (all-reduced)
  double red_llw[8], red_build[8];
  // ---- Metropolis state
  double bounds[2 * kMaxPar];
  double U[kMaxPar];
  double predict_param[kMaxPar];
  double last_logaccept, last_u;
  double paramsd[kMaxPar * kMaxPar], prodparam[kMaxPar * kMaxPar], scratch[2 * kMaxPar * kMaxPar];
  int pred_valid;      // the prediction weights (Hpred, sd) belong to the current theta
  double nobs[kMaxQ];  // observed rows per outcome, whole problem (gibbs_sample_tausq, spamtree_model.cpp:1401)
};

}  // namespace st
