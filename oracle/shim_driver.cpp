// oracle/shim_driver.cpp — TEST INFRASTRUCTURE ONLY.  Calls spamtree_b200/shim/spamtree_fit_b200.cpp (the Rcpp shim of
// INTEGRATION.md §1, compiled against the Armadillo/Rcpp stand-in of oracle/refshim/) the way R's .Call would, and hands its
// returned list back as plain arrays, so that tests/test_gpu_shim.py can compare the shim's results with the ctypes path.
#include <RcppArmadillo.h>

#include <cstdint>
#include <cstdio>

Rcpp::List spamtree_mv_mcmc(const arma::mat& y, const arma::mat& X, const arma::mat& Z, const arma::mat& coords, const arma::uvec& mv_id,
                            const arma::uvec& blocking, const arma::uvec& gix_block, const arma::uvec& res_is_ref,
                            const arma::field<arma::uvec>& parents, const arma::field<arma::uvec>& children, bool limited_tree,
                            const arma::vec& layer_names, const arma::vec& layer_gibbs_group, const arma::field<arma::uvec>& indexing,
                            const arma::mat& set_unif_bounds_in, const arma::mat& start_w, const arma::vec& theta, const arma::vec& beta,
                            const double& tausq, const arma::mat& mcmcsd, int mcmc_keep, int mcmc_burn, int mcmc_thin, int num_threads,
                            char use_alg, bool adapting, bool main_verbose, bool verbose, bool debug, bool printall, bool sample_beta,
                            bool sample_tausq, bool sample_theta, bool sample_w, bool sample_predicts);

namespace {
arma::field<arma::uvec> csr_field(const int64_t* ptr, const int64_t* idx, int n) {
  arma::field<arma::uvec> f(n);
  for (int i = 0; i < n; i++) {
    arma::umat u(ptr[i + 1] - ptr[i], 1);
    for (int64_t k = ptr[i]; k < ptr[i + 1]; k++) u.mem[k - ptr[i]] = (arma::uword)idx[k];
    f(i) = arma::uvec(u);
  }
  return f;
}
}  // namespace

extern "C" int shim_spamtree_mv_mcmc(int64_t n_all, int p, int q, const double* y, const double* X, const double* coords, const int64_t* mv_id,
                                     int n_blocks, const int64_t* idx_ptr, const int64_t* idx, const int64_t* par_ptr, const int64_t* par,
                                     const int64_t* chi_ptr, const int64_t* chi, const double* block_names, const double* block_groups,
                                     const int64_t* res_is_ref, int n_res, int limited_tree, const double* theta, int n_theta,
                                     const double* beta, double tausq, const double* bounds, const double* mcmcsd, int keep, int burn,
                                     int thin, int adapting, int sample_predicts, double runif_value, double* beta_mcmc, double* tausq_mcmc,
                                     double* theta_mcmc, double* w_mcmc, double* yhat_mcmc, double* paramsd_out, int64_t* block_ct_obs_out,
                                     int64_t* parents_indexing_len_out, int64_t* parents_indexing_sum_out) {
  try {
    auto& hooks = arma::rng_hooks();
    const auto old = hooks;
    hooks.unif = [runif_value]() { return runif_value; };  // R::runif(0, 1) of the shim's seed draw
    arma::mat ym(y, n_all, 1), Xm(X, n_all, p), Zm(n_all, q), cm(coords, n_all, 2), w0(n_all, q);
    arma::umat mv(n_all, 1), blocking(n_all, 1), gix(n_all, 1), rr(n_res, 1);
    for (int64_t i = 0; i < n_all; i++) mv.mem[i] = (arma::uword)mv_id[i];
    for (int i = 0; i < n_res; i++) rr.mem[i] = (arma::uword)res_is_ref[i];
    arma::mat bn(block_names, n_blocks, 1), bg(block_groups, n_blocks, 1), th(theta, n_theta, 1), be(beta, p, 1);
    const Rcpp::List l = spamtree_mv_mcmc(ym, Xm, Zm, cm, arma::uvec(mv), arma::uvec(blocking), arma::uvec(gix), arma::uvec(rr),
                                          csr_field(par_ptr, par, n_blocks), csr_field(chi_ptr, chi, n_blocks), limited_tree != 0,
                                          arma::vec(bn), arma::vec(bg), csr_field(idx_ptr, idx, n_blocks), arma::mat(bounds, n_theta, 2), w0,
                                          arma::vec(th), arma::vec(be), tausq, arma::mat(mcmcsd, n_theta, n_theta), keep, burn, thin, 1, 'S',
                                          adapting != 0, false, false, false, false, true, true, true, true, sample_predicts != 0);
    hooks = old;
    const auto& bm = l.get<arma::cube>("beta_mcmc");
    for (int j = 0; j < q; j++)
      for (int s = 0; s < keep; s++)
        for (int a = 0; a < p; a++) beta_mcmc[a + (size_t)s * p + (size_t)j * p * keep] = bm.slice(j)(a, s);
    const auto& tm = l.get<arma::mat>("tausq_mcmc");
    const auto& thm = l.get<arma::mat>("theta_mcmc");
    const auto& ps = l.get<arma::mat>("paramsd");
    std::copy(tm.mem.begin(), tm.mem.end(), tausq_mcmc);
    std::copy(thm.mem.begin(), thm.mem.end(), theta_mcmc);
    std::copy(ps.mem.begin(), ps.mem.end(), paramsd_out);
    const auto& wm = l.get<arma::field<arma::mat>>("w_mcmc");
    const auto& yh = l.get<arma::field<arma::mat>>("yhat_mcmc");
    for (int s = 0; s < keep; s++) {
      std::copy(wm(s).mem.begin(), wm(s).mem.end(), w_mcmc + (size_t)s * n_all);
      std::copy(yh(s).mem.begin(), yh(s).mem.end(), yhat_mcmc + (size_t)s * n_all);
    }
    const auto& bco = l.get<arma::uvec>("block_ct_obs");
    const auto& pix = l.get<arma::field<arma::uvec>>("parents_indexing");
    for (int u = 0; u < n_blocks; u++) {
      block_ct_obs_out[u] = (int64_t)bco(u);
      parents_indexing_len_out[u] = (int64_t)pix(u).n_elem;
      int64_t s = 0;
      for (arma::uword k = 0; k < pix(u).n_elem; k++) s += (int64_t)pix(u)(k);
      parents_indexing_sum_out[u] = s;
    }
    (void)l.get<double>("mcmc_time");
    return 0;
  } catch (const std::exception& ex) {
    fprintf(stderr, "shim_spamtree_mv_mcmc: %s\n", ex.what());
    return 1;
  } catch (...) {
    return 1;
  }
}
