// oracle/ref_driver.cpp — TEST INFRASTRUCTURE ONLY.
// Drives the reference's OWN SpamTreeMV class (compiled unmodified from /root/reference/src against oracle/refshim/)
// through a small C interface that mirrors the oracle's, so that tests can pin oracle/spamtree_oracle.cpp to it.
#include <cstdint>
#include <cstring>

#include "spamtree_model.h"  // the reference's header, from /root/reference/src
// (spamtree_model.h includes mh_adapt.h: RAMAdapt, par_huvtransf_*, calc_jacobian, do_I_accept)

namespace {
struct Rng {  // same host stream as the oracle and the product (xoshiro256++ / Box-Muller / Marsaglia-Tsang)
  uint64_t s[4];
  bool have = false;
  double spare = 0;
  static uint64_t splitmix(uint64_t& x) {
    uint64_t z = (x += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
  }
  void seed(uint64_t sd) { for (int i = 0; i < 4; i++) s[i] = splitmix(sd); have = false; }
  static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
  uint64_t next() {
    const uint64_t r = rotl(s[0] + s[3], 23) + s[0], t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
    return r;
  }
  double unif() { return ((next() >> 11) + 0.5) * (1.0 / 9007199254740992.0); }
  double norm() {
    if (have) { have = false; return spare; }
    const double u1 = unif(), u2 = unif(), rad = std::sqrt(-2.0 * std::log(u1)), ang = 6.283185307179586476925286766559 * u2;
    spare = rad * std::sin(ang); have = true;
    return rad * std::cos(ang);
  }
  double gamma(double shape, double scale) {
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
struct Ref {
  SpamTreeMV m;
  Rng rng;
};
arma::field<arma::uvec> csr_field(const int64_t* ptr, const int64_t* idx, int n) {
  arma::field<arma::uvec> f(n);
  for (int i = 0; i < n; i++) {
    arma::umat u(ptr[i + 1] - ptr[i], 1);
    for (int64_t k = ptr[i]; k < ptr[i + 1]; k++) u.mem[k - ptr[i]] = (arma::uword)idx[k];
    f(i) = arma::uvec(u);
  }
  return f;
}
template <class M> int64_t put(const M& x, double* out, int64_t cap) {
  if (out) for (int64_t i = 0; i < (int64_t)x.n_elem && i < cap; i++) out[i] = (double)x.mem[i];
  return (int64_t)x.n_elem;
}
}  // namespace

// exported functions of tree_dep.cpp that no reference header declares
arma::vec kthresholds(arma::vec x, int k);
arma::mat part_axis_parallel_lmt(const arma::mat& coords, const arma::field<arma::vec>& thresholds);
arma::umat number_revalue(const arma::umat& original_mat, const arma::uvec& from_val, const arma::uvec& to_val);
Rcpp::List make_edges(const arma::mat& parchimat, const arma::uvec& non_empty_blocks, const arma::uvec& res_is_ref);
Rcpp::List make_edges_limited(const arma::mat& parchimat, const arma::uvec& non_empty_blocks, const arma::uvec& res_is_ref);
// the reference's MCMC driver itself (spamtree_fit.cpp:5-430), compiled unmodified
Rcpp::List spamtree_mv_mcmc(const arma::mat& y, const arma::mat& X, const arma::mat& Z, const arma::mat& coords, const arma::uvec& mv_id,
                            const arma::uvec& blocking, const arma::uvec& gix_block, const arma::uvec& res_is_ref,
                            const arma::field<arma::uvec>& parents, const arma::field<arma::uvec>& children, bool limited_tree,
                            const arma::vec& layer_names, const arma::vec& layer_gibbs_group, const arma::field<arma::uvec>& indexing,
                            const arma::mat& set_unif_bounds_in, const arma::mat& start_w, const arma::vec& theta, const arma::vec& beta,
                            const double& tausq, const arma::mat& mcmcsd, int mcmc_keep, int mcmc_burn, int mcmc_thin, int num_threads,
                            char use_alg, bool adapting, bool main_verbose, bool verbose, bool debug, bool printall, bool sample_beta,
                            bool sample_tausq, bool sample_theta, bool sample_w, bool sample_predicts);

extern "C" {

void* ref_create(int64_t n_all, int p, int q, const double* y, const double* X, const double* coords, const int64_t* mv_id,
                 int n_blocks, const int64_t* idx_ptr, const int64_t* idx, const int64_t* par_ptr, const int64_t* par,
                 const int64_t* chi_ptr, const int64_t* chi, const double* block_names, const double* block_groups,
                 const int64_t* res_is_ref, int n_res, int limited_tree, const double* theta, int n_theta,
                 const double* beta, double tausq) {
  try {
    Ref* r = new Ref();
    r->rng.seed(1);
    auto& hooks = arma::rng_hooks();
    hooks.norm = [r]() { return r->rng.norm(); };
    hooks.unif = [r]() { return r->rng.unif(); };
    hooks.gamma = [r](double a, double b) { return r->rng.gamma(a, b); };
    arma::mat ym(y, n_all, 1), Xm(X, n_all, p), Zm(n_all, q), cm(coords, n_all, 2), w0(n_all, 1);
    arma::umat mv(n_all, 1), blocking(n_all, 1), gix(n_all, 1), rr(n_res, 1);
    for (int64_t i = 0; i < n_all; i++) mv.mem[i] = (arma::uword)mv_id[i];
    for (int i = 0; i < n_res; i++) rr.mem[i] = (arma::uword)res_is_ref[i];
    arma::mat bn(block_names, n_blocks, 1), bg(block_groups, n_blocks, 1), th(theta, n_theta, 1), be(beta, p, 1);
    r->m = SpamTreeMV(ym, Xm, Zm, cm, arma::uvec(mv), arma::uvec(blocking), arma::uvec(gix), arma::uvec(rr),
                      csr_field(par_ptr, par, n_blocks), csr_field(chi_ptr, chi, n_blocks), limited_tree != 0, arma::vec(bn),
                      arma::vec(bg), csr_field(idx_ptr, idx, n_blocks), w0, arma::vec(be), arma::vec(th), 1.0 / tausq, 'S', 1,
                      false, false);
    return r;
  } catch (...) {
    return nullptr;
  }
}
void ref_destroy(void* h) { delete (Ref*)h; }
void ref_seed(void* h, uint64_t s) { ((Ref*)h)->rng.seed(s); }
static SpamTreeMVData& slot(Ref* r, int s) { return s ? r->m.alter_data : r->m.param_data; }
void ref_theta_update(void* h, int s, const double* theta) {
  Ref* r = (Ref*)h;
  arma::mat th(theta, slot(r, s).theta.n_elem, 1);
  r->m.theta_update(slot(r, s), arma::vec(th));
}
int ref_build(void* h, int s, double* out3) {  // SpamTreeMV::get_loglik_comps_w
  Ref* r = (Ref*)h;
  bool ok = r->m.get_loglik_comps_w(slot(r, s));
  out3[0] = slot(r, s).loglik_w; out3[1] = slot(r, s).logdetCi; out3[2] = ok;
  return ok;
}
void ref_loglik_w(void* h, int s, double* out2) {
  Ref* r = (Ref*)h;
  r->m.get_loglik_w(slot(r, s));
  out2[0] = slot(r, s).loglik_w; out2[1] = slot(r, s).logdetCi;
}
int ref_gibbs(void* h, const double* z) {  // SpamTreeMV::deal_with_w(true); z feeds its internal arma::randn
  Ref* r = (Ref*)h;
  if (z) arma::rng_hooks().injected.assign(z, z + r->m.coords.n_rows);
  try { r->m.deal_with_w(true); } catch (...) { return 0; }
  return 1;
}
void ref_swap(void* h) { ((Ref*)h)->m.accept_make_change(); }
void ref_predict(void* h, int theta_changed) { ((Ref*)h)->m.predict(theta_changed != 0); }
void ref_sample_beta(void* h, const double* zb) {  // zb: p*q normals consumed by the arma::randn(p) calls, outcome by outcome
  Ref* r = (Ref*)h;
  if (!zb) { r->m.deal_with_beta(); return; }
  // gibbs_sample_beta draws arma::randn(p) once per outcome: feed them through the norm hook in order
  const int p = r->m.p, q = r->m.q;
  int k = 0;
  auto old = arma::rng_hooks().norm;
  arma::rng_hooks().norm = [&]() { return zb[k < p * q ? k++ : p * q - 1]; };
  r->m.deal_with_beta();
  arma::rng_hooks().norm = old;
}
void ref_sample_tausq(void* h, const double* fixed) {
  Ref* r = (Ref*)h;
  if (!fixed) { r->m.gibbs_sample_tausq(); return; }
  int k = 0;
  auto old = arma::rng_hooks().gamma;
  arma::rng_hooks().gamma = [&](double, double) { return fixed[k++]; };
  r->m.gibbs_sample_tausq();
  arma::rng_hooks().gamma = old;
}
void ref_get_w(void* h, double* out) { Ref* r = (Ref*)h; std::copy(r->m.w.mem.begin(), r->m.w.mem.end(), out); }
void ref_set_w(void* h, const double* in) { Ref* r = (Ref*)h; std::copy(in, in + r->m.w.n_elem, r->m.w.mem.begin()); }
void ref_set_tausq_inv(void* h, const double* t) {
  Ref* r = (Ref*)h;
  for (int j = 0; j < r->m.q; j++) {
    r->m.tausq_inv(j) = t[j];
    for (arma::uword i = 0; i < r->m.ix_by_q(j).n_elem; i++) r->m.tausq_inv_long(r->m.ix_by_q(j)(i)) = t[j];
  }
}
void ref_get_params(void* h, double* beta_pq, double* tausq_inv_q, double* xb_n) {
  Ref* r = (Ref*)h;
  if (beta_pq) std::copy(r->m.Bcoeff.mem.begin(), r->m.Bcoeff.mem.end(), beta_pq);
  if (tausq_inv_q) std::copy(r->m.tausq_inv.mem.begin(), r->m.tausq_inv.mem.end(), tausq_inv_q);
  if (xb_n) std::copy(r->m.XB.mem.begin(), r->m.XB.mem.end(), xb_n);
}
int64_t ref_get(void* h, const char* name, int s, int u, int c, double* out, int64_t cap) {
  Ref* r = (Ref*)h;
  SpamTreeMVData& d = slot(r, s);
  const std::string n(name);
  if (n == "H") return put(d.w_cond_mean_K(u), out, cap);
  if (n == "Ri") return put(d.Rcc_invchol(u), out, cap);
  if (n == "prec") return put(d.w_cond_prec(u), out, cap);
  if (n == "ccholprecdiag") return put(d.ccholprecdiag(u), out, cap);
  if (n == "Kxx_inv") return put(d.Kxx_inv(u), out, cap);
  if (n == "Kxc") return put(d.Kxc(u), out, cap);
  if (n == "Sigi_chol") return put(d.Sigi_chol(u), out, cap);
  if (n == "Smu_children") return put(d.Smu_children(u), out, cap);
  if (n == "logdetCi_comps") return put(d.logdetCi_comps, out, cap);
  if (n == "loglik_w_comps") return put(d.loglik_w_comps, out, cap);
  if (n == "parents_indexing") return put(r->m.parents_indexing(u), out, cap);
  if (n == "children_indexing") return put(r->m.children_indexing(u), out, cap);
  if (n == "dim_by_parent") return put(r->m.dim_by_parent(u), out, cap);
  if (n == "this_is_jth_child") return put(r->m.this_is_jth_child(u), out, cap);
  if (n == "u_by_block_groups") return put(r->m.u_by_block_groups(u), out, cap);
  if (n == "blocks_not_empty") return put(r->m.blocks_not_empty, out, cap);
  if (n == "blocks_predicting") return put(r->m.blocks_predicting, out, cap);
  if (n == "block_is_reference") return put(r->m.block_is_reference, out, cap);
  if (n == "block_ct_obs") return put(r->m.block_ct_obs, out, cap);
  if (n == "n_actual_groups") { if (out && cap > 0) out[0] = r->m.n_actual_groups; return 1; }
  if (n == "u_is_which_col") return put(r->m.u_is_which_col_f(u)(c)(s), out, cap);  // s: 0 local, 1 other
  return -1;
}

// ---- standalone exports of the reference, called directly
void ref_kthresholds(const double* x, int64_t n, int k, double* res) {
  arma::mat xm(x, n, 1);
  arma::vec r = kthresholds(arma::vec(xm), k);
  std::copy(r.mem.begin(), r.mem.end(), res);
}
void ref_cross_covariance_ag10(const double* c1, const int64_t* mv1, int64_t n1, const double* c2, const int64_t* mv2, int64_t n2,
                               const double* ai1, const double* ai2, const double* phi_i, const double* thetamv, int n_thetamv,
                               const double* Dmat, int q, double* out) {
  arma::umat m1(n1, 1), m2(n2, 1);
  for (int64_t i = 0; i < n1; i++) m1.mem[i] = (arma::uword)mv1[i];
  for (int64_t i = 0; i < n2; i++) m2.mem[i] = (arma::uword)mv2[i];
  arma::mat r = CrossCovarianceAG10(arma::mat(c1, n1, 2), arma::uvec(m1), arma::mat(c2, n2, 2), arma::uvec(m2),
                                    arma::vec(arma::mat(ai1, q, 1)), arma::vec(arma::mat(ai2, q, 1)), arma::vec(arma::mat(phi_i, q, 1)),
                                    arma::vec(arma::mat(thetamv, n_thetamv, 1)), arma::mat(Dmat, q, q));
  std::copy(r.mem.begin(), r.mem.end(), out);
}
void ref_number_revalue(const int64_t* orig, int64_t nr, int nc, const int64_t* from_val, const int64_t* to_val, int64_t nfrom, int64_t* out) {
  arma::umat om(nr, nc), fv(nfrom, 1), tv(nfrom, 1);
  for (int64_t i = 0; i < nr * nc; i++) om.mem[i] = (arma::uword)orig[i];
  for (int64_t i = 0; i < nfrom; i++) { fv.mem[i] = (arma::uword)from_val[i]; tv.mem[i] = (arma::uword)to_val[i]; }
  arma::umat r = number_revalue(om, arma::uvec(fv), arma::uvec(tv));
  for (int64_t i = 0; i < nr * nc; i++) out[i] = (int64_t)r.mem[i];
}


// make_edges / make_edges_limited (tree_dep.cpp:75-186), called directly.  parchimat: nr x L column-major, NaN = NA.
// Two-call pattern: with NULL outputs only counts = {n_blocks, total parents, total children} is filled.
int ref_make_edges(const double* parchimat, int64_t nr, int L, const int64_t* non_empty, int64_t n_ne, const int64_t* res_is_ref,
                   int limited, int64_t* par_ptr, int64_t* par_idx, int64_t* chi_ptr, int64_t* chi_idx, int64_t* counts) {
  try {
    arma::umat ne(n_ne, 1), rr(L, 1);
    for (int64_t i = 0; i < n_ne; i++) ne.mem[i] = (arma::uword)non_empty[i];
    for (int i = 0; i < L; i++) rr.mem[i] = (arma::uword)res_is_ref[i];
    const arma::mat pm(parchimat, nr, L);
    const Rcpp::List l = limited ? make_edges_limited(pm, arma::uvec(ne), arma::uvec(rr)) : make_edges(pm, arma::uvec(ne), arma::uvec(rr));
    const auto& par = l.get<arma::field<arma::uvec>>("parents");
    const auto& chi = l.get<arma::field<arma::uvec>>("children");
    int64_t np = 0, nc = 0;
    for (arma::uword i = 0; i < par.n_elem; i++) { np += par(i).n_elem; nc += chi(i).n_elem; }
    counts[0] = (int64_t)par.n_elem; counts[1] = np; counts[2] = nc;
    if (!par_ptr) return 0;
    int64_t a = 0, b = 0;
    for (arma::uword i = 0; i < par.n_elem; i++) {
      par_ptr[i] = a; chi_ptr[i] = b;
      for (arma::uword k = 0; k < par(i).n_elem; k++) par_idx[a++] = (int64_t)par(i)(k);
      for (arma::uword k = 0; k < chi(i).n_elem; k++) chi_idx[b++] = (int64_t)chi(i)(k);
    }
    par_ptr[par.n_elem] = a; chi_ptr[par.n_elem] = b;
    return 0;
  } catch (const std::exception& ex) {
    fprintf(stderr, "ref_make_edges: %s\n", ex.what());
    return 1;
  } catch (...) {
    return 1;
  }
}
// part_axis_parallel_lmt (tree_dep.cpp:58-67): coords n x d column-major; thresholds concatenated, thr_ptr d + 1
void ref_part_axis_parallel_lmt(const double* coords, int64_t n, int d, const double* thr, const int64_t* thr_ptr, double* out) {
  arma::field<arma::vec> th(d);
  for (int j = 0; j < d; j++) th(j) = arma::vec(arma::mat(thr + thr_ptr[j], thr_ptr[j + 1] - thr_ptr[j], 1));
  const arma::mat r = part_axis_parallel_lmt(arma::mat(coords, n, d), th);
  std::copy(r.mem.begin(), r.mem.end(), out);
}

// ---- the MH glue of mh_adapt.h, driven directly
// RAMAdapt over a recorded sequence: U (npar x steps, column-major), alpha (steps), iteration numbers mc = 0 .. steps-1.
// paramsd_out: npar x npar after the last step; paramsd_trace (or NULL): npar*npar per step.
int ref_ram_adapt(int npar, const double* metropolis_sd, int steps, const double* U, const double* alpha, double* paramsd_out,
                  double* paramsd_trace) {
  try {
    RAMAdapt ad(npar, arma::mat(metropolis_sd, npar, npar));
    for (int m = 0; m < steps; m++) {
      ad.count_proposal();
      ad.adapt(arma::vec(arma::mat(U + (size_t)m * npar, npar, 1)), alpha[m], m);
      ad.update_ratios();
      if (paramsd_trace) std::copy(ad.paramsd.mem.begin(), ad.paramsd.mem.end(), paramsd_trace + (size_t)m * npar * npar);
    }
    std::copy(ad.paramsd.mem.begin(), ad.paramsd.mem.end(), paramsd_out);
    return 0;
  } catch (...) {
    return 1;
  }
}
// one proposal as spamtree_fit.cpp:211-215 forms it: new = back(fwd(param) + paramsd U), clipped by unif_bounds;
// out: new_param (npar), then the Jacobian term calc_jacobian(new, param) (mh_adapt.h:230-239), then out_unif_bounds
void ref_propose(int npar, const double* param, const double* bounds, const double* paramsd, const double* U, double* out) {
  const arma::mat B(bounds, npar, 2), sd(paramsd, npar, npar);
  const arma::vec par(arma::mat(param, npar, 1)), u(arma::mat(U, npar, 1));
  arma::vec np_ = par_huvtransf_back(par_huvtransf_fwd(par, B) + sd * u, B);
  const bool oob = unif_bounds(np_, B);
  std::copy(np_.mem.begin(), np_.mem.end(), out);
  out[npar] = calc_jacobian(np_, par, B);
  out[npar + 1] = oob ? 1.0 : 0.0;
}
// do_I_accept (mh_adapt.h:20-36) with the uniform supplied by the caller
int ref_do_i_accept(double logaccept, double u) {
  auto old = arma::rng_hooks().unif;
  arma::rng_hooks().unif = [u]() { return u; };
  const bool r = do_I_accept(logaccept);
  arma::rng_hooks().unif = old;
  return r ? 1 : 0;
}

// ---- the reference's whole MCMC driver, spamtree_mv_mcmc (spamtree_fit.cpp:5-430), on the host stream seeded with `seed`
// outputs as in oracle's or_mcmc; ints_out (or NULL): block_ct_obs (n_blocks).  Returns 0, or 1 when the driver threw /
// returned its "None" list.
int ref_spamtree_mv_mcmc(int64_t n_all, int p, int q, const double* y, const double* X, const double* coords, const int64_t* mv_id,
                         int n_blocks, const int64_t* idx_ptr, const int64_t* idx, const int64_t* par_ptr, const int64_t* par,
                         const int64_t* chi_ptr, const int64_t* chi, const double* block_names, const double* block_groups,
                         const int64_t* res_is_ref, int n_res, int limited_tree, const double* theta, int n_theta, const double* beta,
                         double tausq, const double* bounds, const double* mcmcsd, int keep, int burn, int thin, int adapting,
                         int sample_beta, int sample_tausq, int sample_theta, int sample_w, int sample_predicts, uint64_t seed,
                         double* beta_mcmc, double* tausq_mcmc, double* theta_mcmc, double* w_mcmc, double* yhat_mcmc, double* paramsd_out,
                         int64_t* block_ct_obs_out, int64_t* parents_indexing_len_out) {
  try {
    Rng rng;
    rng.seed(seed);
    auto& hooks = arma::rng_hooks();
    const auto old = hooks;
    hooks.injected.clear();
    hooks.norm = [&rng]() { return rng.norm(); };
    hooks.unif = [&rng]() { return rng.unif(); };
    hooks.gamma = [&rng](double a, double b) { return rng.gamma(a, b); };
    arma::mat ym(y, n_all, 1), Xm(X, n_all, p), Zm(n_all, q), cm(coords, n_all, 2), w0(n_all, q);
    arma::umat mv(n_all, 1), blocking(n_all, 1), gix(n_all, 1), rr(n_res, 1);
    for (int64_t i = 0; i < n_all; i++) mv.mem[i] = (arma::uword)mv_id[i];
    for (int i = 0; i < n_res; i++) rr.mem[i] = (arma::uword)res_is_ref[i];
    arma::mat bn(block_names, n_blocks, 1), bg(block_groups, n_blocks, 1), th(theta, n_theta, 1), be(beta, p, 1);
    const Rcpp::List l = spamtree_mv_mcmc(ym, Xm, Zm, cm, arma::uvec(mv), arma::uvec(blocking), arma::uvec(gix), arma::uvec(rr),
                                          csr_field(par_ptr, par, n_blocks), csr_field(chi_ptr, chi, n_blocks), limited_tree != 0,
                                          arma::vec(bn), arma::vec(bg), csr_field(idx_ptr, idx, n_blocks), arma::mat(bounds, n_theta, 2), w0,
                                          arma::vec(th), arma::vec(be), tausq, arma::mat(mcmcsd, n_theta, n_theta), keep, burn, thin, 1, 'S',
                                          adapting != 0, false, false, false, false, sample_beta != 0, sample_tausq != 0, sample_theta != 0,
                                          sample_w != 0, sample_predicts != 0);
    hooks = old;
    if (!l.has("theta_mcmc")) return 1;
    const auto& bm = l.get<arma::cube>("beta_mcmc");  // p x keep x q
    for (int j = 0; j < q; j++)
      for (int s = 0; s < keep; s++)
        for (int a = 0; a < p; a++) beta_mcmc[a + (size_t)s * p + (size_t)j * p * keep] = bm.slice(j)(a, s);
    put(l.get<arma::mat>("tausq_mcmc"), tausq_mcmc, (int64_t)q * keep);
    put(l.get<arma::mat>("theta_mcmc"), theta_mcmc, (int64_t)n_theta * keep);
    put(l.get<arma::mat>("paramsd"), paramsd_out, (int64_t)n_theta * n_theta);
    const auto& wm = l.get<arma::field<arma::mat>>("w_mcmc");
    const auto& yh = l.get<arma::field<arma::mat>>("yhat_mcmc");
    for (int s = 0; s < keep; s++) {
      if (w_mcmc) put(wm(s), w_mcmc + (size_t)s * n_all, n_all);
      if (yhat_mcmc) put(yh(s), yhat_mcmc + (size_t)s * n_all, n_all);
    }
    if (block_ct_obs_out) {
      const auto& b = l.get<arma::uvec>("block_ct_obs");
      for (arma::uword i = 0; i < b.n_elem; i++) block_ct_obs_out[i] = (int64_t)b(i);
    }
    if (parents_indexing_len_out) {
      const auto& pi = l.get<arma::field<arma::uvec>>("parents_indexing");
      for (arma::uword i = 0; i < pi.n_elem; i++) parents_indexing_len_out[i] = (int64_t)pi(i).n_elem;
    }
    return 0;
  } catch (...) {
    return 1;
  }
}

}  // extern "C"
