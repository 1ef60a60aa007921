// st_capi.cpp — the extern "C" boundary declared in include/spamtree_b200.h.  No exception leaves this file.
#include <cmath>
#include <cstring>
#include <new>
#include <string>

#include "../../include/spamtree_b200.h"
#include "st_kernels.cuh"
#include <cuda_runtime.h>

#include "st_mh.hpp"
#include "st_model.hpp"
#include "st_tree.hpp"

namespace st {
int mcmc_run(Model& M, const st_mcmc_opts& o, st_mcmc_out& out);
}

struct st_handle {
  st::Model model;
};
struct st_tree {
  st::TreeResult t;
};

static thread_local std::string g_create_error;

#define ST_GUARD_BEGIN try {
#define ST_GUARD_END(h)                                                   \
  }                                                                       \
  catch (const std::bad_alloc&) {                                         \
    if (h) (h)->model.err = "out of host memory";                         \
    return ST_ERR_INVALID;                                                \
  }                                                                       \
  catch (const std::exception& ex) {                                      \
    if (h) (h)->model.err = ex.what();                                    \
    return ST_ERR_INVALID;                                                \
  }                                                                       \
  catch (...) {                                                           \
    if (h) (h)->model.err = "unknown exception";                          \
    return ST_ERR_INVALID;                                                \
  }

static void csr_assign(st::CSR& c, const int64_t* ptr, const int64_t* idx, int64_t n) {
  c.ptr.assign(ptr, ptr + n + 1);
  c.idx.assign(idx, idx + ptr[n]);
}

// every entry point that touches device state makes the handle's GPU the calling thread's current device first (a host may
// hold handles on several GPUs; streams, launches and the shared-memory opt-in are all per device)
static inline void activate(st_handle* h) {
  if (h && h->model.device >= 0) cudaSetDevice(h->model.device);
}

extern "C" {

const char* st_version(void) { return "spamtree_b200 0.1 sm_100a"; }

int st_create(const st_problem* pr, st_handle** out) {
  if (!pr || !out) return ST_ERR_INVALID;
  *out = nullptr;
  st_handle* h = nullptr;
  try {
    if (pr->n_all <= 0 || pr->p <= 0 || pr->q <= 0 || pr->n_blocks <= 0) { g_create_error = "empty problem"; return ST_ERR_INVALID; }
    h = new st_handle();
    st::Model& M = h->model;
    M.n_all = pr->n_all; M.p = pr->p; M.q = pr->q; M.n_blocks = pr->n_blocks;
    M.y.assign(pr->y, pr->y + pr->n_all);
    M.X.assign(pr->X, pr->X + (size_t)pr->n_all * pr->p);
    M.coords.assign(pr->coords, pr->coords + (size_t)pr->n_all * 2);
    M.mv_id.assign(pr->mv_id, pr->mv_id + pr->n_all);
    csr_assign(M.indexing, pr->indexing_ptr, pr->indexing_idx, pr->n_blocks);
    csr_assign(M.parents, pr->parents_ptr, pr->parents_idx, pr->n_blocks);
    csr_assign(M.children, pr->children_ptr, pr->children_idx, pr->n_blocks);
    M.block_names.assign(pr->block_names, pr->block_names + pr->n_blocks);
    M.block_groups.assign(pr->block_groups, pr->block_groups + pr->n_blocks);
    M.res_is_ref.assign(pr->res_is_ref, pr->res_is_ref + pr->n_res);
    M.keep_H = pr->keep_H != 0;
    M.limited = pr->limited_tree != 0;
    M.device = pr->device;
    if (pr->smem_panel_bytes > 0) M.smem_budget = (size_t)pr->smem_panel_bytes;
    M.theta[0].assign(pr->theta, pr->theta + pr->n_theta);
    M.theta[1] = M.theta[0];
    M.Bcoeff.assign((size_t)pr->p * pr->q, 0.0);
    for (int j = 0; j < pr->q; j++)
      for (int a = 0; a < pr->p; a++) M.Bcoeff[a + (size_t)j * pr->p] = pr->beta[a];  // spamtree_model.cpp:124-129
    M.tausq_inv.assign(pr->q, 1.0 / pr->tausq);                                         // :118
    M.rng.seed(1);
    if (pr->partition && pr->partition->nranks > 1) {
      const st_partition& pt = *pr->partition;
      if (pt.rank < 0 || pt.rank >= pt.nranks) { g_create_error = "partition: bad rank"; delete h; return ST_ERR_INVALID; }
      M.part = true;
      M.rank = pt.rank; M.nranks = pt.nranks; M.n_top_levels = pt.n_top_levels;
      M.rng_row_offset = pt.rng_row_offset; M.n_global_rows = pt.n_global_rows;
      // every random stream is keyed by the row's id in the whole problem (device) or drawn for all of its rows (host): without
      // the map the ranks would draw different numbers for the rows they share
      if (!pt.global_rows || pt.n_global_rows < pr->n_all) { g_create_error = "partition: global_rows / n_global_rows missing"; delete h; return ST_ERR_INVALID; }
      for (int64_t i = 0; i < pr->n_all; i++)
        if (pt.global_rows[i] < 0 || pt.global_rows[i] >= pt.n_global_rows) { g_create_error = "partition: global row out of range"; delete h; return ST_ERR_INVALID; }
      M.global_rows.assign(pt.global_rows, pt.global_rows + pr->n_all);
      M.allreduce_fn = pt.allreduce; M.allreduce_ctx = pt.ctx;
    }
    {
      st::CovTab tab;
      std::string e;
      if (!st::make_covtab(M.theta[0].data(), pr->n_theta, pr->q, tab, e)) { g_create_error = e; delete h; return ST_ERR_INVALID; }
    }
    std::string e;
    int rc = M.init(e);
    if (rc) { g_create_error = e.empty() ? M.err : e; delete h; return rc; }
    *out = h;
    return ST_OK;
  } catch (const std::exception& ex) {
    g_create_error = ex.what();
  } catch (...) {
    g_create_error = "unknown exception";
  }
  delete h;
  return ST_ERR_INVALID;
}

void st_destroy(st_handle* h) {
  activate(h);
  delete h;
}

const char* st_last_error(const st_handle* h) { return h ? h->model.err.c_str() : g_create_error.c_str(); }

int st_theta_update(st_handle* h, int slot, const double* theta) {
  if (!h || !theta) return ST_ERR_INVALID;
  ST_GUARD_BEGIN return h->model.theta_update(slot, theta);
  ST_GUARD_END(h)
}
int st_get_loglik_comps_w(st_handle* h, int slot, double* out3) {
  if (!h || !out3) return ST_ERR_INVALID;
  activate(h);
  ST_GUARD_BEGIN return h->model.get_loglik_comps_w(slot, out3);
  ST_GUARD_END(h)
}
int st_deal_with_w(st_handle* h, const double* z, uint64_t seed) {
  if (!h) return ST_ERR_INVALID;
  activate(h);
  ST_GUARD_BEGIN return h->model.deal_with_w(z, seed);
  ST_GUARD_END(h)
}
int st_get_loglik_w(st_handle* h, int slot, double* out2) {
  if (!h || !out2) return ST_ERR_INVALID;
  activate(h);
  ST_GUARD_BEGIN return h->model.get_loglik_w(slot, out2);
  ST_GUARD_END(h)
}
int st_accept_make_change(st_handle* h) {
  if (!h) return ST_ERR_INVALID;
  activate(h);
  h->model.accept_make_change();
  return ST_OK;
}
int st_predict(st_handle* h, int theta_changed) {
  if (!h) return ST_ERR_INVALID;
  activate(h);
  ST_GUARD_BEGIN return h->model.predict(theta_changed != 0);
  ST_GUARD_END(h)
}
int st_gibbs_sample_beta(st_handle* h, const double* zb, int faithful_index) {
  if (!h) return ST_ERR_INVALID;
  activate(h);
  ST_GUARD_BEGIN return h->model.gibbs_sample_beta(zb, faithful_index != 0);
  ST_GUARD_END(h)
}
int st_gibbs_sample_tausq(st_handle* h, const double* fixed) {
  if (!h) return ST_ERR_INVALID;
  activate(h);
  ST_GUARD_BEGIN return h->model.gibbs_sample_tausq(fixed);
  ST_GUARD_END(h)
}
int st_seed(st_handle* h, uint64_t seed) {
  if (!h) return ST_ERR_INVALID;
  h->model.rng.seed(seed);
  return ST_OK;
}
int st_get_w(st_handle* h, double* w_out) {
  if (!h || !w_out) return ST_ERR_INVALID;
  activate(h);
  ST_GUARD_BEGIN return h->model.get_w(w_out);
  ST_GUARD_END(h)
}
int st_set_w(st_handle* h, const double* w_in) {
  if (!h || !w_in) return ST_ERR_INVALID;
  activate(h);
  ST_GUARD_BEGIN return h->model.set_w(w_in);
  ST_GUARD_END(h)
}
int st_get_params(st_handle* h, double* Bcoeff, double* tausq_inv, double* XB) {
  if (!h) return ST_ERR_INVALID;
  activate(h);
  ST_GUARD_BEGIN
  st::Model& M = h->model;
  if (Bcoeff) std::copy(M.Bcoeff.begin(), M.Bcoeff.end(), Bcoeff);
  if (tausq_inv) std::copy(M.tausq_inv.begin(), M.tausq_inv.end(), tausq_inv);
  if (XB) return M.get_xb(XB);
  return ST_OK;
  ST_GUARD_END(h)
}
int st_set_tausq_inv(st_handle* h, const double* t) {
  if (!h || !t) return ST_ERR_INVALID;
  activate(h);
  ST_GUARD_BEGIN return h->model.set_tausq_inv(t);
  ST_GUARD_END(h)
}
int st_get_node_state(st_handle* h, int slot, int u, const char* which, double* out, int64_t cap, int64_t* count) {
  if (!h || !which || !count) return ST_ERR_INVALID;
  activate(h);
  ST_GUARD_BEGIN return h->model.get_node_state(slot, u, which, out, cap, count);
  ST_GUARD_END(h)
}
int st_get_index(st_handle* h, const char* which, int u, int c, int64_t* out, int64_t cap, int64_t* count) {
  if (!h || !which || !count) return ST_ERR_INVALID;
  ST_GUARD_BEGIN return h->model.get_index(which, u, c, out, cap, count);
  ST_GUARD_END(h)
}
int st_mcmc_run(st_handle* h, const st_mcmc_opts* opts, st_mcmc_out* out) {
  if (!h || !opts || !out || !opts->set_unif_bounds || !opts->mcmcsd || opts->keep < 0 || opts->thin < 1) return ST_ERR_INVALID;
  activate(h);
  ST_GUARD_BEGIN return st::mcmc_run(h->model, *opts, *out);
  ST_GUARD_END(h)
}
int st_bench_iteration(st_handle* h, const double* theta_prop, int do_swap, uint64_t seed, double* out3, float* ms_out) {
  if (!h || !theta_prop || !out3) return ST_ERR_INVALID;
  activate(h);
  ST_GUARD_BEGIN return h->model.bench_iteration(theta_prop, do_swap, seed, out3, ms_out);
  ST_GUARD_END(h)
}
int st_par_huvtransf_fwd(const double* par, int32_t npar, const double* bounds, double* out) {
  if (!par || !bounds || !out || npar < 0) return ST_ERR_INVALID;
  for (int j = 0; j < npar; j++) out[j] = st::mh_logit(par[j], bounds[j], bounds[j + npar]);
  return ST_OK;
}
int st_par_huvtransf_back(const double* par, int32_t npar, const double* bounds, double* out) {
  if (!par || !bounds || !out || npar < 0) return ST_ERR_INVALID;
  for (int j = 0; j < npar; j++) out[j] = st::mh_logistic(par[j], bounds[j], bounds[j + npar]);
  return ST_OK;
}
int st_mh_propose(int32_t npar, const double* param, const double* bounds, const double* paramsd, const double* U, double* out) {
  if (!param || !bounds || !paramsd || !U || !out || npar < 1 || npar > st::kMaxPar) return ST_ERR_INVALID;
  const bool oob = st::mh_propose(npar, param, bounds, paramsd, U, out);
  out[npar] = st::mh_jacobian(npar, out, param, bounds);
  out[npar + 1] = oob ? 1.0 : 0.0;
  return ST_OK;
}
int st_do_i_accept(double logaccept, double u) { return u < st::mh_accept_prob(logaccept) ? 1 : 0; }
int st_ram_adapt(int32_t npar, const double* metropolis_sd, int32_t steps, const double* U, const double* alpha, double* paramsd_out,
                 double* paramsd_trace) {
  if (!metropolis_sd || !paramsd_out || npar < 1 || npar > st::kMaxPar || steps < 0 || (steps > 0 && (!U || !alpha))) return ST_ERR_INVALID;
  try {
    const size_t pp = (size_t)npar * npar;
    std::vector<double> sd(pp), prod(pp), scratch(2 * pp);
    int started = 0;
    if (!st::ram_init(npar, metropolis_sd, sd.data(), prod.data())) return ST_ERR_INVALID;
    for (int m = 0; m < steps; m++) {
      st::ram_adapt(npar, sd.data(), prod.data(), &started, U + (size_t)m * npar, alpha[m], m, scratch.data());
      if (paramsd_trace) std::copy(sd.begin(), sd.end(), paramsd_trace + (size_t)m * pp);
    }
    std::copy(sd.begin(), sd.end(), paramsd_out);
    return ST_OK;
  } catch (...) {
    return ST_ERR_INVALID;
  }
}
int st_set_beta_index(st_handle* h, int faithful) {
  if (!h) return ST_ERR_INVALID;
  if (h->model.part && faithful) { h->model.err = "partitioned handles support only the corrected beta row index"; return ST_ERR_UNSUPPORTED; }
  activate(h);
  ST_GUARD_BEGIN return h->model.set_widx_mode(faithful != 0);
  ST_GUARD_END(h)
}
int st_nccl_unique_id(unsigned char* out128) {
  if (!out128) return ST_ERR_INVALID;
  std::string e;
  const int rc = st::Model::nccl_unique_id(out128, e);
  if (rc) g_create_error = e;
  return rc;
}
int st_attach_nccl(st_handle* h, const unsigned char* id128) {
  if (!h || !id128) return ST_ERR_INVALID;
  activate(h);
  return h->model.attach_nccl(id128);
}
int st_get_counters(st_handle* h, double* out8) {
  if (!h || !out8) return ST_ERR_INVALID;
  out8[0] = h->model.n_launches; out8[1] = h->model.f_alg; out8[2] = h->model.f_exec; out8[3] = h->model.n_cov;
  out8[4] = h->model.f_alg_build; out8[5] = h->model.f_exec_build; out8[6] = h->model.b_alg_build; out8[7] = (double)sizeof(st::ChainDev);
  return ST_OK;
}
int st_sync(st_handle* h) {
  if (!h) return ST_ERR_INVALID;
  activate(h);
  return h->model.sync();
}

// ---- DAG construction
int st_kthresholds(const double* x, int64_t n, int32_t k, double* res) {
  if (!x || n <= 0 || k < 1 || (k > 1 && !res)) return ST_ERR_INVALID;
  try { st::kthresholds(x, n, k, res); } catch (...) { return ST_ERR_INVALID; }
  return ST_OK;
}
int st_part_axis_parallel_lmt(const double* coords, int64_t n, int32_t d, const double* thr, const int64_t* thr_ptr, double* out) {
  if (!coords || !thr_ptr || !out) return ST_ERR_INVALID;
  try { st::part_axis_parallel_lmt(coords, n, d, thr, thr_ptr, out); } catch (...) { return ST_ERR_INVALID; }
  return ST_OK;
}
int st_number_revalue(const int64_t* orig, int64_t nr, int32_t nc, const int64_t* from_val, const int64_t* to_val, int64_t nfrom, int64_t* out) {
  if (!orig || !out) return ST_ERR_INVALID;
  try { st::number_revalue(orig, nr, nc, from_val, to_val, nfrom, out); } catch (...) { return ST_ERR_INVALID; }
  return ST_OK;
}
int st_make_edges(const double* parchimat, int64_t nr, int32_t L, const int64_t* non_empty_blocks, int64_t n_ne,
                  const int64_t* res_is_ref, int32_t limited, int64_t* par_ptr, int64_t* par_idx, int64_t* chi_ptr,
                  int64_t* chi_idx, int64_t* counts) {
  if (!parchimat || !res_is_ref || !counts || L < 1) return ST_ERR_INVALID;
  try {
    st::CSR par, chi;
    int64_t nb = 0;
    st::make_edges(parchimat, nr, L, non_empty_blocks, n_ne, res_is_ref, limited != 0, par, chi, nb);
    counts[0] = nb; counts[1] = (int64_t)par.idx.size(); counts[2] = (int64_t)chi.idx.size();
    if (par_idx) {
      std::copy(par.ptr.begin(), par.ptr.end(), par_ptr);
      std::copy(par.idx.begin(), par.idx.end(), par_idx);
      std::copy(chi.ptr.begin(), chi.ptr.end(), chi_ptr);
      std::copy(chi.idx.begin(), chi.idx.end(), chi_idx);
    }
  } catch (...) { return ST_ERR_INVALID; }
  return ST_OK;
}

int st_make_tree(const st_tree_opts* o, st_tree** out) {
  if (!o || !out || !o->coords || !o->y || !o->mv_id || o->n_all <= 0) return ST_ERR_INVALID;
  *out = nullptr;
  try {
    st_tree* t = new st_tree();
    std::string e;
    if (!st::make_tree(o->coords, o->y, o->mv_id, o->n_all, o->cell_size, o->K[0], o->K[1], o->start_level, o->tree_depth,
                       o->last_not_reference != 0, o->cherrypick_same_margin != 0, o->cherrypick_group_locations != 0,
                       o->seed, t->t, e)) {
      g_create_error = e;
      delete t;
      return ST_ERR_INVALID;
    }
    *out = t;
  } catch (const std::exception& ex) { g_create_error = ex.what(); return ST_ERR_INVALID; } catch (...) { return ST_ERR_INVALID; }
  return ST_OK;
}
void st_tree_destroy(st_tree* t) { delete t; }
int st_tree_sizes(const st_tree* t, int64_t* o) {
  if (!t || !o) return ST_ERR_INVALID;
  o[0] = t->t.n_blocks; o[1] = (int64_t)t->t.res_is_ref.size(); o[2] = t->t.parchi_rows; o[3] = t->t.parchi_cols;
  o[4] = (int64_t)t->t.parents.idx.size(); o[5] = (int64_t)t->t.children.idx.size();
  return ST_OK;
}
int st_tree_get(const st_tree* t, int64_t* blocking, int64_t* res, int64_t* res_is_ref, double* parchimat,
                int64_t* indexing_ptr, int64_t* indexing_idx, int64_t* parents_ptr, int64_t* parents_idx,
                int64_t* children_ptr, int64_t* children_idx, double* block_names, double* block_groups) {
  if (!t) return ST_ERR_INVALID;
  const st::TreeResult& T = t->t;
  auto cp = [](const auto& v, auto* dst) { if (dst) std::copy(v.begin(), v.end(), dst); };
  cp(T.blocking, blocking); cp(T.res, res); cp(T.res_is_ref, res_is_ref); cp(T.parchimat, parchimat);
  cp(T.indexing.ptr, indexing_ptr); cp(T.indexing.idx, indexing_idx);
  cp(T.parents.ptr, parents_ptr); cp(T.parents.idx, parents_idx);
  cp(T.children.ptr, children_ptr); cp(T.children.idx, children_idx);
  cp(T.block_names, block_names); cp(T.block_groups, block_groups);
  return ST_OK;
}

int st_cross_covariance_ag10(const double* coords1, const int64_t* mv1, int64_t n1, const double* coords2, const int64_t* mv2,
                             int64_t n2, const double* ai1, const double* ai2, const double* phi_i, const double* thetamv,
                             int32_t n_thetamv, const double* Dmat, int32_t q, int32_t device, double* out) {
  if (!coords1 || !coords2 || !mv1 || !mv2 || !out || q < 2 || q > st::kMaxQ) {
    g_create_error = "Invalid Dmat for multivariate data";  // covariance_functions.cpp:317-319
    return ST_ERR_INVALID;
  }
  try {
    // pack (ai1, ai2, phi_i, thetamv, lower triangle of Dmat) in the theta layout and reuse make_covtab
    const int n_cbase = q > 2 ? 3 : 1;
    if (n_thetamv < n_cbase) return ST_ERR_INVALID;
    st::dvec th;
    th.insert(th.end(), ai1, ai1 + q); th.insert(th.end(), ai2, ai2 + q); th.insert(th.end(), phi_i, phi_i + q);
    th.insert(th.end(), thetamv, thetamv + n_cbase);
    for (int j = 0; j < q; j++)
      for (int i = j + 1; i < q; i++) th.push_back(Dmat[i + (size_t)j * q]);
    st::CovTab tab;
    std::string e;
    if (!st::make_covtab(th.data(), (int)th.size(), q, tab, e)) { g_create_error = e; return ST_ERR_INVALID; }
    if (cudaSetDevice(device) != cudaSuccess) { g_create_error = "cudaSetDevice failed"; return ST_ERR_CUDA; }
    std::vector<int> q1(n1), q2(n2);
    for (int64_t i = 0; i < n1; i++) q1[i] = (int)mv1[i] - 1;
    for (int64_t i = 0; i < n2; i++) q2[i] = (int)mv2[i] - 1;
    for (int v : q1) if (v < 0 || v >= q) { g_create_error = "mv_id outside 1..q"; return ST_ERR_INVALID; }
    for (int v : q2) if (v < 0 || v >= q) { g_create_error = "mv_id outside 1..q"; return ST_ERR_INVALID; }
    double *dx1 = nullptr, *dx2 = nullptr, *dout = nullptr;
    int *dq1 = nullptr, *dq2 = nullptr;
    cudaError_t ce = cudaSuccess;
    auto chk = [&](cudaError_t c) { if (ce == cudaSuccess) ce = c; };
    chk(cudaMalloc((void**)&dx1, 2 * n1 * sizeof(double)));
    chk(cudaMalloc((void**)&dx2, 2 * n2 * sizeof(double)));
    chk(cudaMalloc((void**)&dq1, n1 * sizeof(int)));
    chk(cudaMalloc((void**)&dq2, n2 * sizeof(int)));
    chk(cudaMalloc((void**)&dout, (size_t)n1 * n2 * sizeof(double)));
    if (ce == cudaSuccess) {
      chk(cudaMemcpy(dx1, coords1, 2 * n1 * sizeof(double), cudaMemcpyHostToDevice));
      chk(cudaMemcpy(dx2, coords2, 2 * n2 * sizeof(double), cudaMemcpyHostToDevice));
      chk(cudaMemcpy(dq1, q1.data(), n1 * sizeof(int), cudaMemcpyHostToDevice));
      chk(cudaMemcpy(dq2, q2.data(), n2 * sizeof(int), cudaMemcpyHostToDevice));
      chk(st::launch_crosscov(dx1, dx1 + n1, dq1, n1, dx2, dx2 + n2, dq2, n2, tab, dout, 0));
      chk(cudaMemcpy(out, dout, (size_t)n1 * n2 * sizeof(double), cudaMemcpyDeviceToHost));
    }
    cudaFree(dx1); cudaFree(dx2); cudaFree(dq1); cudaFree(dq2); cudaFree(dout);
    if (ce != cudaSuccess) { g_create_error = std::string("CUDA error: ") + cudaGetErrorString(ce); return ST_ERR_CUDA; }
  } catch (...) { return ST_ERR_INVALID; }
  return ST_OK;
}

}  // extern "C"
