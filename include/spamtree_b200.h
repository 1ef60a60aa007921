/* spamtree_b200.h — C ABI of libspamtree_b200.so
 *
 * Drop-in boundary for the per-iteration MCMC hot path of SpamTrees (reference:
 * mkln/spamtree v0.2.1).  Every entry point below replaces one piece of the
 * reference's Rcpp/C++ model layer; the reference interface it stands in for is
 * cited as file:line (paths relative to the reference tree).  Plain pointers and
 * sizes only; no C++ exceptions cross this boundary; every function returns an
 * st_status (0 = ok) unless stated otherwise.  All matrices are FP64 COLUMN-MAJOR
 * (as arma::mat), all row ids / block ids 0-based unless stated otherwise, rows in
 * the caller's ("boundary") order — the order spamtree() passes to
 * spamtree_mv_mcmc (R/spamtree_fit.R:267-269, :327-362).
 *
 * Not re-entrant per handle.  One host thread drives one GPU per handle.
 */
#ifndef SPAMTREE_B200_H
#define SPAMTREE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct st_handle st_handle; /* opaque: the SpamTreeMV object (src/spamtree_model.h:22-212) */

typedef enum {
  ST_OK = 0,
  ST_ERR_INVALID = 1,   /* bad argument / inconsistent DAG (reference: `throw 1`, spamtree_model.cpp:201-226) */
  ST_ERR_CUDA = 2,      /* CUDA runtime error; st_last_error() has the text */
  ST_ERR_NOT_SPD = 3,   /* Cholesky failed inside the Gibbs step (reference: Rcpp::stop, spamtree_model.cpp:1215-1217) */
  ST_ERR_UNSUPPORTED = 4, /* shape or option outside what this build handles (e.g. mvbias != 0, q > 8) */
  ST_ERR_NAN = 5        /* NaN log-likelihood at the current theta (reference: `throw 1`, spamtree_fit.cpp:234-237) */
} st_status;

/* Multi-GPU: one problem cut into subtrees, one rank (process + GPU) per run of subtrees (the reference is
 * single-process).  Every rank creates its handle from ITS blocks: the blocks of the first n_top_levels tree levels are
 * present on every rank (replicated, computed redundantly), the rest are the rank's own subtrees.  The library calls
 * `allreduce` (in-place sum over ranks of `count` doubles at a DEVICE pointer, after synchronising its stream) for the
 * three scalars of the log-density, the messages of the cut-level blocks to the replicated blocks, and the beta / tausq
 * sufficient statistics.  spamtree_b200/partition.py builds these inputs; the callback is torch.distributed (NCCL). */
typedef int (*st_allreduce_fn)(void* ctx, void* device_ptr, int64_t count);
typedef struct {
  int32_t rank, nranks;
  int32_t n_top_levels;         /* replicated levels (0: the ranks hold disjoint trees) */
  int64_t rng_row_offset;       /* rows owned by lower ranks: keeps the device normal streams of the ranks disjoint */
  int64_t n_global_rows;        /* rows of the whole problem */
  const int64_t* global_rows;   /* n_all, required: row of the whole problem behind every local row (keys the device random
                                   streams; the host stream draws for all rows of the problem on every rank) */
  st_allreduce_fn allreduce;
  void* ctx;
} st_partition;

/* Inputs of the SpamTreeMV constructor (spamtree_model.cpp:8-37), i.e. the arguments
 * spamtree_mv_mcmc receives from R (spamtree_fit.cpp:5-54).  Lists of uvec are CSR.
 * Accepted-and-ignored reference arguments (Z values, blocking, gix_block, start_w,
 * use_alg) have no field here. */
typedef struct {
  int64_t n_all;               /* rows incl. rows whose y is missing */
  int32_t p;                   /* covariates */
  int32_t q;                   /* outcomes = #unique(mv_id) */
  const double* y;             /* n_all; NaN = missing (to be predicted) */
  const double* X;             /* n_all x p */
  const double* coords;        /* n_all x 2 */
  const int64_t* mv_id;        /* n_all, 1-based outcome id */
  int32_t n_blocks;
  const int64_t* indexing_ptr; /* n_blocks+1 */
  const int64_t* indexing_idx; /* rows of each block, ascending (R/spamtree_fit.R:324) */
  const int64_t* parents_ptr;  /* n_blocks+1 */
  const int64_t* parents_idx;  /* make_edges output, tree_dep.cpp:113-119 */
  const int64_t* children_ptr; /* n_blocks+1 */
  const int64_t* children_idx; /* make_edges output, tree_dep.cpp:102-110 */
  const double* block_names;   /* n_blocks, 1-based names (layer_names) */
  const double* block_groups;  /* n_blocks, tree level of block id i (layer_gibbs_group) */
  const int64_t* res_is_ref;   /* n_res flags per tree level */
  int32_t n_res;
  int32_t limited_tree;        /* 1: every block conditions on its direct parent only (make_edges_limited, tree_dep.cpp:133-186;
                                  spamtree_model.cpp:901-903, 1275-1278) */
  const double* theta;         /* n_theta start values (covariance_functions.cpp:34-52 layout) */
  int32_t n_theta;
  const double* beta;          /* p start values */
  double tausq;                /* start value (the ctor receives 1/tausq, spamtree_fit.cpp:104) */
  int32_t device;              /* CUDA device ordinal; < 0: host-only handle (bookkeeping for st_get_index, no compute) */
  int32_t keep_H;              /* 1: keep H = w_cond_mean_K of observed blocks and the Sigi_tot / Smu_tot probes of the Gibbs
                                  sweep on the device (needed by st_get_node_state "H", "Sigi_tot", "Smu_tot"); 0: production */
  int64_t smem_panel_bytes;    /* 0 = default; shared-memory budget for one BUILD work group */
  const st_partition* partition; /* NULL: the whole problem on one GPU */
} st_problem;

/* SpamTreeMV::SpamTreeMV — spamtree_model.cpp:8-192 (+ init_indexing :315, na_study :303,
 * make_gibbs_groups :194, init_finalize :355, init_model_data :422).  Copies inputs. */
int st_create(const st_problem* prob, st_handle** out);
void st_destroy(st_handle* h);
/* text of the last error on this handle (or of the last failed st_create when h == NULL) */
const char* st_last_error(const st_handle* h);

/* slot: 0 = param_data, 1 = alter_data (tree_utils.h:63-102; spamtree_model.h:112-113) */

/* SpamTreeMV::theta_update — spamtree_model.cpp:1420-1422 */
int st_theta_update(st_handle* h, int slot, const double* theta);
/* SpamTreeMV::get_loglik_comps_w — spamtree_model.cpp:829-998 (BUILD).
 * out3 = {loglik_w, logdetCi, ok}; ok == 0 (some Cholesky failed) is NOT an error:
 * the reference returns false and the proposal is rejected (spamtree_fit.cpp:223,249). */
int st_get_loglik_comps_w(st_handle* h, int slot, double* out3);
/* SpamTreeMV::deal_with_w(true) — spamtree_model.cpp:1000-1226 (GIBBS sweep over w).
 * z: n_all standard normals in boundary order (the reference's bigrnorm, :1018), or NULL to
 * draw them on the device from (seed, counter). */
int st_deal_with_w(st_handle* h, const double* z, uint64_t seed);
/* SpamTreeMV::get_loglik_w — spamtree_model.cpp:776-826 (LLW). out2 = {loglik_w, logdetCi} */
int st_get_loglik_w(st_handle* h, int slot, double* out2);
/* SpamTreeMV::accept_make_change — spamtree_model.cpp:1432-1435 */
int st_accept_make_change(st_handle* h);
/* SpamTreeMV::predict(theta_changed) — spamtree_model.cpp:1229-1358; uses the z of the last st_deal_with_w */
int st_predict(st_handle* h, int theta_changed);
/* SpamTreeMV::gibbs_sample_beta — spamtree_model.cpp:1364-1391. zb: p*q normals or NULL (host stream).
 * faithful_index != 0 reproduces the reference's row mis-indexing with missing data. */
int st_gibbs_sample_beta(st_handle* h, const double* zb, int faithful_index);
/* SpamTreeMV::gibbs_sample_tausq — spamtree_model.cpp:1393-1417. fixed: q values to set tausq_inv to, or NULL to draw */
int st_gibbs_sample_tausq(st_handle* h, const double* fixed);
/* seed of the host random stream used by the two functions above and by st_mcmc_run */
int st_seed(st_handle* h, uint64_t seed);

/* public fields the driver reads/writes (spamtree_fit.cpp:118-120, 378-384) */
int st_get_w(st_handle* h, double* w_out /* n_all, boundary order */);
int st_set_w(st_handle* h, const double* w_in);
int st_get_params(st_handle* h, double* Bcoeff /* p x q or NULL */, double* tausq_inv /* q or NULL */,
                  double* XB /* n_all or NULL */);
int st_set_tausq_inv(st_handle* h, const double* tausq_inv /* q */);

/* Per-block state for parity tests.  which:
 *  "H"  w_cond_mean_K(u)  m x P      (spamtree_model.cpp:887)      [needs keep_H]
 *  "Ri" Rcc_invchol(u)    m x m      (:896) ; for non-reference blocks ccholprecdiag(u), m (:945)
 *  "Sigi_tot" m x m (:1044-1051) and "Smu_tot" m (:1062-1077) as seen by the last st_deal_with_w
 *             (non-reference blocks: m scalars each, :1123-1127)
 *  "logdetCi_comps" / "loglik_w_comps": all blocks (u ignored), n_blocks values (:966-968)
 * Writes at most cap doubles; *count receives the element count. */
int st_get_node_state(st_handle* h, int slot, int u, const char* which, double* out, int64_t cap, int64_t* count);

/* Integer bookkeeping for bit-exact parity (spamtree_model.cpp:194-420).  which:
 *  "parents_indexing" (u) · "dim_by_parent" (u) · "this_is_jth_child" (u) · "u_by_block_groups" (u = group)
 *  "blocks_not_empty" · "blocks_predicting" · "block_is_reference" · "block_ct_obs" · "n_actual_groups"
 *  "u_is_which_col" (u, c = child position): {firstcol, lastcol} of spamtree_model.cpp:395-396 */
int st_get_index(st_handle* h, const char* which, int u, int c, int64_t* out, int64_t cap, int64_t* count);

/* ---- the MCMC driver: spamtree_mv_mcmc, spamtree_fit.cpp:5-430 ---- */
typedef struct {
  const double* set_unif_bounds; /* npar x 2 */
  const double* mcmcsd;          /* npar x npar */
  int32_t keep, burn, thin;
  int32_t adapting, sample_beta, sample_tausq, sample_theta, sample_w, sample_predicts;
  int32_t faithful_beta_index;   /* see st_gibbs_sample_beta */
  int32_t rng_mode;              /* 0: host-driven chain, every draw from one host stream in the reference's order (lock-step with
                                    the oracle and with the reference's own driver; one host round trip per step);
                                    1: device-resident chain: proposal, accept decision, slot swap, RAM adaptation, tausq / beta
                                    draws and yhat on the device from Philox streams keyed by (seed, row or parameter,
                                    iteration); the host only enqueues and synchronises once, at the end of the run */
  uint64_t seed;
} st_mcmc_opts;
typedef struct {                 /* caller-allocated; NULL pointers are skipped */
  double* beta_mcmc;             /* p x keep x q */
  double* tausq_mcmc;            /* q x keep */
  double* theta_mcmc;            /* npar x keep */
  double* w_mcmc;                /* n_all x keep */
  double* yhat_mcmc;             /* n_all x keep */
  double* paramsd;               /* npar x npar */
  double mcmc_time;              /* seconds, like the reference's return value (spamtree_fit.cpp:394,413) */
  int64_t n_accepted, n_chol_fail;
} st_mcmc_out;
int st_mcmc_run(st_handle* h, const st_mcmc_opts* opts, st_mcmc_out* out);

/* ---- the Metropolis-Hastings glue (mh_adapt.h / mh_adapt.cpp), as st_mcmc_run uses it; host-only, no handle ---- */
/* par_huvtransf_fwd / par_huvtransf_back (Rcpp exports), mh_adapt.cpp:3-15: logit / logistic on (bounds[j], bounds[j + npar]).
 * bounds: npar x 2. */
int st_par_huvtransf_fwd(const double* par, int32_t npar, const double* bounds, double* out);
int st_par_huvtransf_back(const double* par, int32_t npar, const double* bounds, double* out);
/* One proposal as spamtree_fit.cpp:211-215 forms it: new_param = back(fwd(param) + paramsd U), clipped by unif_bounds
 * (mh_adapt.h:188-202).  out: new_param (npar), calc_jacobian(new_param, param) (mh_adapt.h:230-239), out_unif_bounds (0/1). */
int st_mh_propose(int32_t npar, const double* param, const double* bounds, const double* paramsd, const double* U,
                  double* out /* npar + 2 */);
/* do_I_accept, mh_adapt.h:20-36, with the caller's uniform draw u: returns 1 (accept) or 0 */
int st_do_i_accept(double logaccept, double u);
/* class RAMAdapt, mh_adapt.h:40-135, over a recorded sequence: U npar x steps, alpha[steps], iteration numbers 0..steps-1.
 * paramsd_out npar x npar after the last step; paramsd_trace (or NULL) npar*npar per step.  ST_ERR_INVALID when
 * metropolis_sd is not positive definite (arma::chol throws in the reference, :86). */
int st_ram_adapt(int32_t npar, const double* metropolis_sd, int32_t steps, const double* U, const double* alpha,
                 double* paramsd_out, double* paramsd_trace);

/* ---- multi-GPU: native collectives (no reference counterpart) ---- */
/* Native collective path of a partitioned handle: the library owns an NCCL communicator and enqueues its all-reduces
 * on the handle's own stream (no host synchronisation, no callback).  Rank 0 calls st_nccl_unique_id, the 128 bytes
 * are handed to every rank by whatever means the host has (torch.distributed broadcast in spamtree_b200/dist.py), and
 * every rank calls st_attach_nccl on its handle — collectively, like ncclCommInitRank.  libnccl.so.2 is resolved at
 * run time (dlopen), so handles that never attach need no NCCL at all.  Without it the `allreduce` callback is used. */
#define ST_NCCL_UNIQUE_ID_BYTES 128
int st_nccl_unique_id(unsigned char* out128);
int st_attach_nccl(st_handle* h, const unsigned char* id128);

/* ---- bench / profiling hooks (no reference counterpart) ---- */
/* One hot-path iteration without host random draws: GIBBS (device normals) + LLW + BUILD(alter, theta_prop)
 * + optional swap + tausq + beta (spamtree_fit.cpp:167-330 minus predict/save).  ms_out[0..5] (when non-NULL) receive the
 * CUDA-event times of {gibbs (+ Gram refresh), llw, build on the main stream (+ deferred half), tausq + beta, the whole
 * iteration, the upper levels of BUILD on the second stream (0 when the iteration is sequential: then [0..3] add up to [4])}. */
int st_bench_iteration(st_handle* h, const double* theta_prop, int do_swap, uint64_t seed, double* out3, float* ms_out);
/* row index of the beta step used by st_bench_iteration: 1 = the reference's (the default on a single-GPU
 * handle), 0 = the corrected one (the only one a partitioned handle supports) */
int st_set_beta_index(st_handle* h, int faithful);
/* counts of kernel launches and algorithmic work: out[8] = {kernel launches since creation, F_alg flops of one iteration
 * (SURVEY §8d formula on the actual tree), executed-flop estimate of the lean formulation, covariance evaluations,
 * F_alg of BUILD alone (F_build), executed-flop estimate of BUILD alone, compulsory output bytes of one BUILD,
 * bytes of the chain state that a device-resident run sends up / brings back once} */
int st_get_counters(st_handle* h, double* out8);
int st_sync(st_handle* h);

/* ---- DAG construction: tree_dep.cpp ---- */
/* kthresholds, tree_dep.cpp:16-27.  res: k-1 values */
int st_kthresholds(const double* x, int64_t n, int32_t k, double* res);
/* part_axis_parallel_lmt, tree_dep.cpp:58-67 (thresholds of axis j at thr[thr_ptr[j]..thr_ptr[j+1])); out n x d */
int st_part_axis_parallel_lmt(const double* coords, int64_t n, int32_t d, const double* thr, const int64_t* thr_ptr,
                              double* out);
/* number_revalue, tree_dep.cpp:240-259 */
int st_number_revalue(const int64_t* orig, int64_t nr, int32_t nc, const int64_t* from_val, const int64_t* to_val,
                      int64_t nfrom, int64_t* out);
/* make_edges / make_edges_limited, tree_dep.cpp:75-186.  parchimat nr x L (NaN = NA, 1-based block names).
 * Two-call protocol: with par_idx == NULL only counts[0..2] = {n_blocks, #parent entries, #child entries} are written. */
int st_make_edges(const double* parchimat, int64_t nr, int32_t L, const int64_t* non_empty_blocks, int64_t n_ne,
                  const int64_t* res_is_ref, int32_t limited, int64_t* par_ptr, int64_t* par_idx, int64_t* chi_ptr,
                  int64_t* chi_idx, int64_t* counts);

/* Deterministic stand-in for R's make_tree() (R/make_tree.R:1-420): same structure, with the
 * per-cell `sample()` (R/make_tree.R:92) replaced by "smallest splitmix64(ix ^ seed) in the cell" and the FNN
 * kd-tree 1-NN replaced by an exact 1-NN.  Rows must already be sorted by (Var1, Var2, ix) as spamtree() does
 * (R/spamtree_fit.R:214).  tree_depth <= 0 means Inf.  Results are fetched from the returned object. */
typedef struct st_tree st_tree;
typedef struct {
  int64_t n_all;
  const double* coords;   /* n_all x 2, sorted */
  const double* y;        /* NaN = missing */
  const int64_t* mv_id;   /* 1-based */
  int32_t cell_size;      /* 25 */
  int32_t K[2];           /* 2,2 */
  int32_t start_level, tree_depth;
  int32_t last_not_reference, cherrypick_same_margin, cherrypick_group_locations;
  uint64_t seed;
} st_tree_opts;
int st_make_tree(const st_tree_opts* o, st_tree** out);
void st_tree_destroy(st_tree* t);
/* sizes: {n_blocks, n_res, parchi rows, parchi cols, #parent entries, #child entries} */
int st_tree_sizes(const st_tree* t, int64_t* out6);
/* blocking/res: n_all (1-based block name and tree level of each row); parchimat: rows x cols (NaN = NA);
 * the rest is exactly what st_problem takes.  Any pointer may be NULL. */
int st_tree_get(const st_tree* t, int64_t* blocking, int64_t* res, int64_t* res_is_ref, double* parchimat,
                int64_t* indexing_ptr, int64_t* indexing_idx, int64_t* parents_ptr, int64_t* parents_idx,
                int64_t* children_ptr, int64_t* children_idx, double* block_names, double* block_groups);

/* CrossCovarianceAG10 (R export), covariance_functions.cpp:301-355, evaluated on the GPU.
 * coords n x 2 col-major, mv 1-based, Dmat q x q; out n1 x n2. */
int st_cross_covariance_ag10(const double* coords1, const int64_t* mv1, int64_t n1, const double* coords2,
                             const int64_t* mv2, int64_t n2, const double* ai1, const double* ai2,
                             const double* phi_i, const double* thetamv, int32_t n_thetamv, const double* Dmat,
                             int32_t q, int32_t device, double* out);

/* library info: returns a static string "spamtree_b200 <version> sm_100a" */
const char* st_version(void);

#ifdef __cplusplus
}
#endif
#endif /* SPAMTREE_B200_H */
