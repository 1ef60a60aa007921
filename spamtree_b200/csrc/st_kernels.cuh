// st_kernels.cuh — launch wrappers of the sm_100a kernels in st_kernels.cu (host-callable; no kernel types leak out)
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

#include "st_model.hpp"

namespace st {

constexpr int kBuildThreads = 256;
constexpr int kBuildTR = 5;  // register tile rows: blocks have ~25 rows (5 x 5 knots per cell, R/spamtree_fit.R:229-233)
constexpr int kBuildTC = 2;
constexpr int kGibbsThreads = 128;
constexpr int kGramThreads = 128;
constexpr int kLlwThreads = 128;
constexpr int kMaxStats = 40;  // q * (p + 1)
constexpr double kHl2pi = -0.91893853320467274178;  // -0.5 * log(2 pi)  (spamtree_model.h:20)

// leading dimension of the shared-memory panel: columns rounded up to 4, made odd (conflict-free column walks)
__host__ __device__ inline int build_ld(int NC) { return ((NC + 3) & ~3) + 1; }
__host__ __device__ inline int build_s1n(int mode, int maxmj, int LD, int sumsq) {
  const int a = maxmj * LD;
  return (mode == 0 && sumsq > a) ? sumsq : a;
}
// dynamic shared memory of build_level_kernel for one work group
inline size_t build_smem_bytes(int mode, int P, int NC, int sumsq, int maxmj) {
  const int LD = build_ld(NC);
  const size_t dbl = (size_t)P * LD + build_s1n(mode, maxmj, LD, sumsq) + (mode == 0 ? sumsq : LD) + 3 * (size_t)P + 3 * (size_t)LD;
  const size_t ints = (size_t)P + 2 * (size_t)LD;
  return dbl * 8 + ints * 4 + 16;
}
inline size_t gibbs_smem_bytes(int is_ref, int m, int P, int k) {
  const size_t msq = is_ref ? (size_t)m * m : (size_t)m;
  return 8 * (2 * msq + (size_t)P + (size_t)(k + 1) * m + 3 * (size_t)m) + 16;
}

cudaError_t launch_build(int mode, const DevTree& T, const DevSlot& S, double* outH, double* outRi, const int* grp_slot0,
                         const int* grp_nn, int ngrp, const double* w, const CovTab& tab, int* fail, int keep_H,
                         size_t smem, cudaStream_t st);
cudaError_t launch_gibbs(int is_ref, const DevTree& T, const DevSlot& S, int slot0, int nslots, double* w,
                         const double* xb, const double* z, const double* tausq_inv, const double* SigS, double* V,
                         double* probe_sig, double* probe_smu, int* fail, size_t smem, cudaStream_t st);
cudaError_t launch_gram(const DevTree& T, const DevSlot& S, int slot0, int nslots, double* U, double* SigS,
                        cudaStream_t st);
cudaError_t launch_llw(const DevTree& T, const DevSlot& S, int nslots, const double* w, cudaStream_t st);
cudaError_t launch_loglik_reduce(const double* logdet, const double* llcomp, int n, const int* fail, double* out,
                                 cudaStream_t st);
cudaError_t launch_predict_sample(const DevTree& T, int slot0, int nslots, const double* Hpred, const double* sdpred,
                                  double* w, const double* z, cudaStream_t st);
cudaError_t launch_normals(double* z, long long n, uint64_t seed, uint64_t counter, cudaStream_t st);
cudaError_t launch_rowstats(const DevTree& T, const int* widx, long long n_all, int p, int q, const double* w,
                            const double* xb, double* partial, int nblocks, double* out, cudaStream_t st);
cudaError_t launch_xb(const DevTree& T, long long n_all, int p, const double* bcoeff, double* xb, cudaStream_t st);
cudaError_t launch_permute(const double* src, double* dst, const long long* map, long long n, cudaStream_t st);
cudaError_t launch_crosscov(const double* x1, const double* y1, const int* q1, long long n1, const double* x2,
                            const double* y2, const int* q2, long long n2, const CovTab& tab, double* out,
                            cudaStream_t st);

}  // namespace st
