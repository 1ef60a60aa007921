// st_mcmc.cpp — the MCMC driver: same control flow as spamtree_mv_mcmc (spamtree_fit.cpp:5-430), calling the model
// layer through the reference-named operations.  Printing and R's interrupt polling have no counterpart here.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "../../include/spamtree_b200.h"
#include "st_model.hpp"

namespace st {

// mh_adapt.h:150-156
static inline double logistic(double x, double l, double u) { return l + (u - l) / (1.0 + std::exp(-x)); }
static inline double logit(double x, double l, double u) { return -std::log((u - l) / (x - l) - 1.0); }

// Robust adaptive Metropolis (Vihola 2012) as the reference runs it: mh_adapt.h:40-135
class RAMAdapt {
 public:
  int p = 0, g0 = 50;
  double alpha_star = .234, gamma = 0.5 + 1e-6;
  SmallMat paramsd, prodparam;
  bool started = false;
  void init(int npars, const double* metropolis_sd) {
    p = npars;
    paramsd = SmallMat(p);
    std::copy(metropolis_sd, metropolis_sd + (size_t)p * p, paramsd.a.begin());
    small_chol(paramsd);  // :86
    prodparam = paramsd;
    for (auto& v : prodparam.a) v /= (g0 + 1.0);  // :87
  }
  void adapt(const dvec& U, double alpha, int mc) {  // :117-135
    if (mc < g0) {
      for (int j = 0; j < p; j++)
        for (int i = 0; i < p; i++) prodparam(i, j) += U[i] * U[j] / (mc + 1.0);
      return;
    }
    if (!started) { paramsd = prodparam; started = true; }
    const int i0 = mc - g0;
    const double eta = std::min(1.0, (p + .0) * std::pow(i0 + 1.0, -gamma));
    alpha = std::min(1.0, alpha);
    double uu = 0;
    for (int i = 0; i < p; i++) uu += U[i] * U[i];
    SmallMat Sigma(p), T(p), S(p);
    for (int j = 0; j < p; j++)
      for (int i = 0; i < p; i++) Sigma(i, j) = (i == j ? 1.0 : 0.0) + eta * (alpha - alpha_star) * U[i] * U[j] / uu;
    // mm(paramsd, Sigma) in the same accumulation order as a column-major axpy product
    for (int j = 0; j < p; j++)
      for (int k = 0; k < p; k++) {
        const double b = Sigma(k, j);
        for (int i = 0; i < p; i++) T(i, j) += paramsd(i, k) * b;
      }
    for (int j = 0; j < p; j++)
      for (int i = 0; i < p; i++) {
        double s = 0;
        for (int k = 0; k < p; k++) s += T(i, k) * paramsd(j, k);
        S(i, j) = s;
      }
    SmallMat L = S;
    if (small_chol(L)) paramsd = L;
  }
};

int mcmc_run(Model& M, const st_mcmc_opts& o, st_mcmc_out& out) {
  M.rng.seed(o.seed);
  double o3[3];
  int rc = M.get_loglik_comps_w(0, o3);  // spamtree_fit.cpp:110-111
  if (rc) return rc;
  rc = M.get_loglik_comps_w(1, o3);
  if (rc) return rc;
  const int npar = (int)M.theta[M.cur].size();
  dvec param = M.theta[M.cur], predict_param = param;
  double current_loglik = M.loglik_w[M.cur];
  const int mcmc = o.thin * o.keep + o.burn;
  RAMAdapt ad;
  ad.init(npar, o.mcmcsd);
  int msaved = 0;
  out.n_accepted = out.n_chol_fail = 0;
  auto lo = [&](int j) { return o.set_unif_bounds[j]; };
  auto hi = [&](int j) { return o.set_unif_bounds[j + npar]; };
  dvec zbuf, wbuf, xbbuf, zglob;
  if (o.rng_mode == 0) zbuf.resize(M.n_all);
  const bool pglob = M.part && !M.global_rows.empty();  // partitioned: draw for every row of the problem, keep ours
  if (pglob) zglob.resize(M.n_global_rows);
  // w of a saved iteration goes to the caller's buffer asynchronously (overlapping the next iteration) unless yhat, which
  // is formed on the host from w, is requested too
  const bool async_save = out.w_mcmc && !out.yhat_mcmc;
  if (async_save) { rc = M.save_begin(out.w_mcmc, (size_t)o.keep * M.n_all * sizeof(double)); if (rc) return rc; }
  struct SaveGuard { Model& M; bool on; ~SaveGuard() { if (on) M.save_end(); } } guard{M, async_save};
  // ST_PROFILE_MCMC=1: host wall-clock per phase of the loop, printed by every rank at the end (development aid)
  static const bool prof = getenv("ST_PROFILE_MCMC") != nullptr;
  double tph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  auto tick = [&]() { return std::chrono::steady_clock::now(); };
  auto lap = [&](int i, std::chrono::steady_clock::time_point& t) {
    if (!prof) return;
    const auto n = tick();
    tph[i] += std::chrono::duration<double>(n - t).count();
    t = n;
  };
  const auto t0 = std::chrono::steady_clock::now();
  for (int m = 0; m < mcmc; m++) {
    auto tl = tick();
    bool predicting = false;
    const int mx = m - o.burn;
    if (mx >= 0 && mx % o.thin == 0) predicting = true;
    if (o.sample_w) {  // :183-187
      if (o.rng_mode == 0) {
        if (pglob) {
          for (int64_t i = 0; i < M.n_global_rows; i++) zglob[i] = M.rng.norm();
          for (int64_t i = 0; i < M.n_all; i++) zbuf[i] = zglob[M.global_rows[i]];
        } else {
          for (int64_t i = 0; i < M.n_all; i++) zbuf[i] = M.rng.norm();  // bigrnorm (:1018)
        }
        rc = M.deal_with_w(zbuf.data(), 0);
      } else {
        rc = M.deal_with_w(nullptr, o.seed);
      }
      if (rc) return rc;
      lap(0, tl);
      double o2[2];
      rc = M.get_loglik_w(0, o2);
      if (rc) return rc;
      current_loglik = M.loglik_w[M.cur];
      lap(1, tl);
    }
    if (o.sample_theta) {  // :203-289
      dvec U(npar), new_param(npar);
      for (int j = 0; j < npar; j++) U[j] = M.rng.norm();
      for (int j = 0; j < npar; j++) {
        double s = logit(param[j], lo(j), hi(j));
        for (int k = 0; k < npar; k++) s += ad.paramsd(j, k) * U[k];
        new_param[j] = logistic(s, lo(j), hi(j));
      }
      for (int j = 0; j < npar; j++) {  // unif_bounds, mh_adapt.h:188-202
        if (new_param[j] < lo(j)) new_param[j] = lo(j) + 1e-10;
        if (new_param[j] > hi(j)) new_param[j] = hi(j) - 1e-10;
      }
      M.theta_update(1, new_param.data());
      rc = M.get_loglik_comps_w(1, o3);
      if (rc) return rc;
      lap(2, tl);
      const bool acceptable = o3[2] != 0.0;
      const double new_loglik = M.loglik_w[1 - M.cur];
      current_loglik = M.loglik_w[M.cur];
      if (std::isnan(current_loglik)) { M.err = "At nan loglik: error."; return ST_ERR_NAN; }
      double jac = 0;  // calc_jacobian, mh_adapt.h:230-239
      for (int j = 0; j < npar; j++)
        jac += (-std::log(hi(j) - param[j]) - std::log(param[j] - lo(j))) -
               (-std::log(hi(j) - new_param[j]) - std::log(new_param[j] - lo(j)));
      const double logaccept = new_loglik - current_loglik + jac;
      double acceptj = 1.0;  // do_I_accept, mh_adapt.h:20-36
      if (!std::isfinite(logaccept)) acceptj = 0.0;
      else if (logaccept < 0) acceptj = std::exp(logaccept);
      const double u = M.rng.unif();
      const bool accepted = (u < acceptj) && acceptable;
      if (!acceptable) out.n_chol_fail++;
      if (accepted) {
        out.n_accepted++;
        current_loglik = new_loglik;
        M.accept_make_change();
        param = new_param;
      }
      if (o.adapting) ad.adapt(U, (acceptable ? 1.0 : 0.0) * std::exp(logaccept), m);  // :285
      lap(3, tl);
    }
    bool need_update = false;  // :300
    for (int j = 0; j < npar; j++)
      if (std::fabs(param[j] - predict_param[j]) > 1e-05) need_update = true;
    if (predicting && o.sample_predicts && o.sample_w) {
      rc = M.predict(need_update);
      if (rc) return rc;
      predict_param = param;
    }
    lap(4, tl);
    if (o.sample_tausq) { rc = M.gibbs_sample_tausq(nullptr); if (rc) return rc; }
    if (o.sample_beta) { rc = M.gibbs_sample_beta(nullptr, o.faithful_beta_index != 0); if (rc) return rc; }
    lap(5, tl);
    if (mx >= 0 && mx % o.thin == 0) {  // save, :376-389
      if (out.tausq_mcmc)
        for (int j = 0; j < M.q; j++) out.tausq_mcmc[j + (size_t)msaved * M.q] = 1.0 / M.tausq_inv[j];
      if (out.beta_mcmc)
        for (int j = 0; j < M.q; j++)
          for (int a = 0; a < M.p; a++)
            out.beta_mcmc[a + (size_t)msaved * M.p + (size_t)j * M.p * o.keep] = M.Bcoeff[a + (size_t)j * M.p];
      if (out.theta_mcmc)
        for (int j = 0; j < npar; j++) out.theta_mcmc[j + (size_t)msaved * npar] = M.theta[M.cur][j];
      const bool want_rows = (out.w_mcmc || out.yhat_mcmc) && !async_save;
      if (async_save) { rc = M.save_w_async(out.w_mcmc + (size_t)msaved * M.n_all); if (rc) return rc; }
      if (want_rows) {
        wbuf.resize(M.n_all);
        rc = M.get_w(wbuf.data());
        if (rc) return rc;
        if (out.w_mcmc) std::copy(wbuf.begin(), wbuf.end(), out.w_mcmc + (size_t)msaved * M.n_all);
      }
      if (out.yhat_mcmc) {
        xbbuf.resize(M.n_all);
        rc = M.get_xb(xbbuf.data());
        if (rc) return rc;
      }
      if (pglob && o.rng_mode == 0)
        for (int64_t i = 0; i < M.n_global_rows; i++) zglob[i] = M.rng.norm();
      if (o.rng_mode == 0 || out.yhat_mcmc)
        for (int64_t i = 0; i < M.n_all; i++) {
          const double e = (pglob && o.rng_mode == 0) ? zglob[M.global_rows[i]] : M.rng.norm();  // arma::randn(n) of :384
          if (out.yhat_mcmc)
            out.yhat_mcmc[i + (size_t)msaved * M.n_all] = xbbuf[i] + wbuf[i] + std::pow(M.tausq_inv[M.mv_id[i] - 1], -.5) * e;
        }
      msaved++;
    }
    lap(6, tl);
  }
  if (async_save) { rc = M.save_sync(); if (rc) return rc; }  // the timed region ends when every saved w is on the host
  out.mcmc_time = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();  // (the guard releases the page lock)
  if (prof)
    fprintf(stderr, "[mcmc profile] rank %d: %d iterations, %lld accepted; ms/iteration: gibbs %.3f llw %.3f build %.3f accept+adapt %.3f predict %.3f tausq+beta %.3f save %.3f | total %.3f\n",
            M.rank, mcmc, (long long)out.n_accepted, 1e3 * tph[0] / mcmc, 1e3 * tph[1] / mcmc, 1e3 * tph[2] / mcmc, 1e3 * tph[3] / mcmc,
            1e3 * tph[4] / mcmc, 1e3 * tph[5] / mcmc, 1e3 * tph[6] / mcmc, 1e3 * out.mcmc_time / mcmc);
  if (out.paramsd) std::copy(ad.paramsd.a.begin(), ad.paramsd.a.end(), out.paramsd);
  return 0;
}

}  // namespace st
