// st_tree.hpp — host-side DAG construction (see st_tree.cpp)
#pragma once
#include <string>

#include "st_common.hpp"

namespace st {

void kthresholds(const double* x, int64_t n, int k, double* res);
void part_axis_parallel_lmt(const double* coords, int64_t n, int d, const double* thr, const int64_t* thr_ptr,
                            double* out);
void number_revalue(const int64_t* orig, int64_t nr, int nc, const int64_t* from_val, const int64_t* to_val,
                    int64_t nfrom, int64_t* out);
void make_edges(const double* parchimat, int64_t nr, int L, const int64_t* non_empty_blocks, int64_t n_ne,
                const int64_t* res_is_ref, bool limited, CSR& parents, CSR& children, int64_t& n_blocks);

struct TreeResult {
  int64_t n_all = 0, n_blocks = 0, parchi_rows = 0;
  int parchi_cols = 0;
  ivec blocking, res, res_is_ref;
  dvec parchimat;  // parchi_rows x parchi_cols, NaN = NA
  CSR indexing, parents, children;
  dvec block_names, block_groups;
};
bool make_tree(const double* coords, const double* y, const int64_t* mv_id, int64_t n_all, int cell_size, int K0,
               int K1, int start_level, int tree_depth, bool last_not_reference, bool cherrypick_same_margin,
               bool cherrypick_group_locations, uint64_t seed, TreeResult& T, std::string& err);

}  // namespace st
