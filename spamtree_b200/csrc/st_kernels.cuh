// st_kernels.cuh — launch wrappers of the sm_100a kernels in st_kernels.cu (host-callable; no kernel types leak out)
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

#include "st_model.hpp"

namespace st {

constexpr int kBuildThreads = 256;
constexpr int kBuildTR = 5;  // register tile rows: blocks have ~25 rows (5 x 5 knots per cell, R/spamtree_fit.R:229-233)
constexpr int kBuildTC = 4;
constexpr int kGibbsThreads = 128;
constexpr int kGramThreads = 128;
constexpr int kLlwThreads = 128;
constexpr int kLlwMaxP = 1024;  // parent-set rows the LLW kernel stages per warp (checked at st_create)
constexpr int kMaxStats = 40;  // q * (p + 1)
constexpr double kHl2pi = -0.91893853320467274178;  // -0.5 * log(2 pi)  (spamtree_model.h:20)

// row stride of a stored tile with `cols` columns: even (16-byte rows) and = 2 mod 4 (conflict-free 5-row gathers)
__host__ __device__ inline int tile_rs(int cols) {
  int r = (cols + 1) & ~1;
  if ((r & 3) == 0) r += 2;
  return r;
}

// ---- shared-memory plan of build_level_kernel (one work group)
constexpr int kBuildStages = 3;   // cp.async ring depth
constexpr int kMaxFam = 8;        // sibling sets ("families") per work group
constexpr int kMaxGroupNodes = 64;
constexpr int kMaxChain = 32;
struct BuildShape {
  int mode;    // 0 reference level, 1 non-reference level, 2 prediction blocks
  int share;   // 1: the deepest ancestor differs per family (cousin group)
  int kc;      // ancestors common to the whole group
  int Pc;      // rows of the common ancestors
  int mmaxs;   // max rows of the family-specific ancestor (0 when share == 0)
  int F;       // families
  int NCp;     // panel columns, every family padded to a multiple of 4
  int sumR;    // mode 0: sum of m_d * tile_rs(m_d) over the group's blocks
  int maxtile; // doubles of the largest operand tile
  int maxmd;   // largest block of the group
};
struct BuildPlan {
  int LD, Ppad, Pv, nsets, Fst;
  size_t o_panel, o_R, o_ring, o_pxs, o_pys, o_wpa, o_cxs, o_cys, o_ecol, o_vtmp, o_pq, o_cq, o_colnode, o_cgfam, o_desc, total;
};
__host__ __device__ inline int build_nsets(int kc, int share) { return kc * (kc + 1) + share * (2 * kc + 2); }
__host__ __device__ inline BuildPlan build_plan(const BuildShape& s) {
  BuildPlan p;
  p.LD = s.NCp + 2;
  p.Ppad = s.Pc + (s.share ? s.mmaxs : 0);
  p.Pv = s.Pc + (s.share ? s.F * s.mmaxs : 0);
  p.nsets = build_nsets(s.kc, s.share);
  p.Fst = s.share ? s.F : 1;
  size_t o = 0;
  auto take = [&](size_t n_doubles) { size_t r = o; o += ((n_doubles + 1) & ~(size_t)1); return r; };
  p.o_panel = take((size_t)p.Ppad * p.LD);
  p.o_R = take(s.mode == 0 ? (size_t)s.sumR : (size_t)p.LD);
  p.o_ring = take((size_t)kBuildStages * p.Fst * s.maxtile + 8);  // +8: mac_cols may read past the last row of a tile
  p.o_pxs = take(p.Pv); p.o_pys = take(p.Pv); p.o_wpa = take(p.Pv);
  p.o_cxs = take(p.LD); p.o_cys = take(p.LD); p.o_ecol = take(p.LD);
  p.o_vtmp = take((size_t)(kBuildThreads / 32) * (s.maxmd + 2));
  p.o_pq = take((p.Pv + 1) / 2); p.o_cq = take((p.LD + 1) / 2); p.o_colnode = take((p.LD + 1) / 2); p.o_cgfam = take((p.LD / 4 + 2) / 2);
  p.o_desc = take((size_t)p.nsets * 4);
  p.total = o * 8 + 16;
  return p;
}
inline size_t gibbs_smem_bytes(int is_ref, int m, int P, int k) {
  const size_t msq = is_ref ? (size_t)m * tile_rs(m) + (size_t)m * m : 2 * (size_t)m;
  return 8 * (msq + (size_t)P + (size_t)(k + 1) * m + 3 * (size_t)m) + 16;
}

cudaError_t launch_build(int mode, const DevTree& T, const DevSlot& S, double* outH, double* outRi, const int* grp_slot0,
                         const int* grp_nn, const int* grp_share, int ngrp, const double* w, const CovTab& tab, int* fail,
                         int keep_H, size_t smem, cudaStream_t st, unsigned long long* prof = nullptr,
                         int nthreads = kBuildThreads);
cudaError_t launch_gibbs(int is_ref, const DevTree& T, const DevSlot& S, int slot0, int nslots, double* w,
                         const double* xb, const double* z, const double* tausq_inv, const double* SigS, double* V,
                         double* probe_sig, double* probe_smu, int* fail, size_t smem, cudaStream_t st);
cudaError_t launch_gram(const DevTree& T, const DevSlot& S, int slot0, int nslots, double* U, double* SigS,
                        cudaStream_t st);
cudaError_t launch_llw(const DevTree& T, const DevSlot& S, int nslots, const double* w, cudaStream_t st);
cudaError_t launch_loglik_reduce(const double* logdet, const double* llcomp, int first, int n, const int* fail,
                                 int fail_as_count, double* out, cudaStream_t st);
cudaError_t launch_frontier_sum(const DevTree& T, int n, const int* pseudo, const int* c0, const int* c1, const int* vlen,
                                const int* ulen, double* V, double* U, int do_v, int do_u, cudaStream_t st);
cudaError_t launch_predict_sample(const DevTree& T, int slot0, int nslots, const double* Hpred, const double* sdpred,
                                  double* w, const double* z, cudaStream_t st);
cudaError_t launch_normals(double* z, long long n, uint64_t seed, uint64_t counter, long long n_shared, long long offset,
                           cudaStream_t st);
cudaError_t launch_rowstats(const DevTree& T, const int* widx, long long n_all, int p, int q, const double* w,
                            const double* xb, double* partial, int nblocks, double* out, cudaStream_t st);
cudaError_t launch_xb(const DevTree& T, long long n_all, int p, const double* bcoeff, double* xb, cudaStream_t st);
cudaError_t launch_permute(const double* src, double* dst, const long long* map, long long n, cudaStream_t st);
cudaError_t launch_crosscov(const double* x1, const double* y1, const int* q1, long long n1, const double* x2,
                            const double* y2, const int* q2, long long n2, const CovTab& tab, double* out,
                            cudaStream_t st);

}  // namespace st
