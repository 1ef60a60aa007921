// spamtree_fit_b200.cpp — the Rcpp shim a maintainer of mkln/spamtree adds to put libspamtree_b200.so behind the package's
// own entry point: it REPLACES the body of spamtree_mv_mcmc (src/spamtree_fit.cpp:5-430), keeps its 35-argument signature
// and its returned list (src/spamtree_fit.cpp:403-414), and calls the C ABI of include/spamtree_b200.h.
//   PKG_CPPFLAGS += -I$(B200_INCLUDE)     PKG_LIBS += -L$(B200_LIB) -lspamtree_b200
// R/spamtree_fit.R stays as it is (it calls spamtree_mv_mcmc positionally, :327-362).
// In this repository the file is compiled against the Armadillo/Rcpp stand-in of oracle/refshim/ and run on the GPU by
// tests/test_gpu_shim.py (R is not installed here); it uses only API that real RcppArmadillo provides as well.
#include <RcppArmadillo.h>

#include <cstdint>
#include <string>
#include <vector>

#include "spamtree_b200.h"

static void to_csr(const arma::field<arma::uvec>& f, std::vector<int64_t>& ptr, std::vector<int64_t>& idx) {
  ptr.assign(1, 0);
  for (arma::uword i = 0; i < f.n_elem; i++) {
    for (arma::uword k = 0; k < f(i).n_elem; k++) idx.push_back((int64_t)f(i)(k));
    ptr.push_back((int64_t)idx.size());
  }
}

//[[Rcpp::export]]
Rcpp::List spamtree_mv_mcmc(const arma::mat& y, const arma::mat& X, const arma::mat& Z, const arma::mat& coords,
    const arma::uvec& mv_id, const arma::uvec& blocking, const arma::uvec& gix_block, const arma::uvec& res_is_ref,
    const arma::field<arma::uvec>& parents, const arma::field<arma::uvec>& children, bool limited_tree,
    const arma::vec& layer_names, const arma::vec& layer_gibbs_group, const arma::field<arma::uvec>& indexing,
    const arma::mat& set_unif_bounds_in, const arma::mat& start_w, const arma::vec& theta, const arma::vec& beta,
    const double& tausq, const arma::mat& mcmcsd, int mcmc_keep = 100, int mcmc_burn = 100, int mcmc_thin = 1,
    int num_threads = 1, char use_alg = 'S', bool adapting = false, bool main_verbose = true, bool verbose = false,
    bool debug = false, bool printall = false, bool sample_beta = true, bool sample_tausq = true,
    bool sample_theta = true, bool sample_w = true, bool sample_predicts = true) {
  // Z's values, blocking, gix_block, start_w, num_threads, use_alg and the verbosity flags are accepted and ignored, exactly
  // as the reference ignores them
  (void)blocking; (void)gix_block; (void)start_w; (void)num_threads; (void)use_alg; (void)main_verbose; (void)verbose; (void)debug; (void)printall;
  std::vector<int64_t> ip, ii, pp, pi, cp, ci, mv(mv_id.begin(), mv_id.end()), rr(res_is_ref.begin(), res_is_ref.end());
  to_csr(indexing, ip, ii); to_csr(parents, pp, pi); to_csr(children, cp, ci);
  const int npar = (int)theta.n_elem, n = (int)y.n_rows, p = (int)X.n_cols, q = (int)Z.n_cols, nb = (int)layer_names.n_elem;
  st_problem pr{};  // arma::mat is column-major FP64: the pointers pass through unchanged
  pr.n_all = n; pr.p = p; pr.q = q;
  pr.y = y.memptr(); pr.X = X.memptr(); pr.coords = coords.memptr(); pr.mv_id = mv.data();
  pr.n_blocks = nb;
  pr.indexing_ptr = ip.data(); pr.indexing_idx = ii.data(); pr.parents_ptr = pp.data(); pr.parents_idx = pi.data();
  pr.children_ptr = cp.data(); pr.children_idx = ci.data();
  pr.block_names = layer_names.memptr(); pr.block_groups = layer_gibbs_group.memptr();
  pr.res_is_ref = rr.data(); pr.n_res = (int32_t)rr.size(); pr.limited_tree = limited_tree ? 1 : 0;
  pr.theta = theta.memptr(); pr.n_theta = npar; pr.beta = beta.memptr(); pr.tausq = tausq;
  pr.device = 0; pr.keep_H = 0;
  st_handle* h = nullptr;
  if (st_create(&pr, &h) != ST_OK) Rcpp::stop(st_last_error(nullptr));
  std::vector<double> beta_buf((size_t)p * mcmc_keep * q);  // p x keep x q
  arma::mat tausq_mcmc(q, mcmc_keep), theta_mcmc(npar, mcmc_keep), w_all(n, mcmc_keep), yhat_all(n, mcmc_keep), paramsd(npar, npar);
  st_mcmc_opts o{};
  o.set_unif_bounds = set_unif_bounds_in.memptr(); o.mcmcsd = mcmcsd.memptr();
  o.keep = mcmc_keep; o.burn = mcmc_burn; o.thin = mcmc_thin;
  o.adapting = adapting; o.sample_beta = sample_beta; o.sample_tausq = sample_tausq; o.sample_theta = sample_theta;
  o.sample_w = sample_w; o.sample_predicts = sample_predicts;
  o.faithful_beta_index = 1;  // the reference's row indexing of the beta step
  o.rng_mode = 1;             // device-resident chain
  o.seed = (uint64_t)(R::runif(0, 1) * 9007199254740992.0);  // drawn from R's stream: set.seed() still governs the run
  st_mcmc_out out{};
  out.beta_mcmc = beta_buf.data(); out.tausq_mcmc = tausq_mcmc.memptr(); out.theta_mcmc = theta_mcmc.memptr();
  out.w_mcmc = w_all.memptr(); out.yhat_mcmc = yhat_all.memptr(); out.paramsd = paramsd.memptr();
  const int rc = st_mcmc_run(h, &o, &out);
  const std::string msg = rc ? st_last_error(h) : "";
  // block_ct_obs and parents_indexing of the reference's return list (spamtree_fit.cpp:410-412)
  arma::uvec block_ct_obs(nb);
  arma::field<arma::uvec> parents_indexing(nb);
  if (!rc) {
    std::vector<int64_t> buf(nb);
    int64_t cnt = 0;
    st_get_index(h, "block_ct_obs", 0, 0, buf.data(), nb, &cnt);
    for (int u = 0; u < nb; u++) block_ct_obs(u) = (arma::uword)buf[u];
    for (int u = 0; u < nb; u++) {
      st_get_index(h, "parents_indexing", u, 0, nullptr, 0, &cnt);
      std::vector<int64_t> pix((size_t)cnt + 1);
      st_get_index(h, "parents_indexing", u, 0, pix.data(), cnt, &cnt);
      arma::uvec v((arma::uword)cnt);
      for (int64_t k = 0; k < cnt; k++) v((arma::uword)k) = (arma::uword)pix[k];
      parents_indexing(u) = v;
    }
  }
  st_destroy(h);
  if (rc) Rcpp::stop(msg);  // ST_ERR_NOT_SPD <-> Rcpp::stop("Error at gibbs_sample_w"), spamtree_model.cpp:1216
  arma::cube beta_mcmc(p, mcmc_keep, q);
  for (int j = 0; j < q; j++)
    for (int s = 0; s < mcmc_keep; s++)
      for (int a = 0; a < p; a++) beta_mcmc(a, s, j) = beta_buf[a + (size_t)s * p + (size_t)j * p * mcmc_keep];
  // the reference returns w_mcmc / yhat_mcmc as lists of n x 1 matrices (spamtree_fit.cpp:134-135, 403-405)
  arma::field<arma::mat> w_mcmc(mcmc_keep), yhat_mcmc(mcmc_keep);
  for (int i = 0; i < mcmc_keep; i++) { w_mcmc(i) = arma::mat(w_all.col(i)); yhat_mcmc(i) = arma::mat(yhat_all.col(i)); }
  return Rcpp::List::create(
      Rcpp::Named("w_mcmc") = w_mcmc, Rcpp::Named("yhat_mcmc") = yhat_mcmc, Rcpp::Named("beta_mcmc") = beta_mcmc,
      Rcpp::Named("tausq_mcmc") = tausq_mcmc, Rcpp::Named("theta_mcmc") = theta_mcmc, Rcpp::Named("paramsd") = paramsd,
      Rcpp::Named("block_ct_obs") = block_ct_obs, Rcpp::Named("indexing") = indexing,
      Rcpp::Named("parents_indexing") = parents_indexing, Rcpp::Named("mcmc_time") = out.mcmc_time);
}
