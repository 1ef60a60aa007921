// st_kernels.cuh — launch wrappers of the sm_100a kernels in st_kernels.cu (host-callable; no kernel types leak out)
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

#include "st_model.hpp"

namespace st {

constexpr int kBuildMaxThreads = 512;  // build_level_kernel: one warp per 8 panel columns (at most 16 warps)
constexpr int kBuildRS = 16;     // rows of the chain's inverse Cholesky factor staged per step (two DMMA m-tiles)
constexpr int kBuildST = 20;     // row stride of a backward stage: 16 columns + 4 (conflict-free fragment reads)
constexpr int kMaxGroupCols = 128;
constexpr int kMaxGroupNodes = 32;
constexpr int kMaxChain = 32;
constexpr int kGibbsThreads = 64;   // two warps per block (measured on C4: 128 -> 1.46 ms, 64 -> 1.14 ms, 32 -> 1.21 ms per sweep)
constexpr int kGramThreads = 256;
constexpr int kGramMaxRows = 256;   // rows gram_level_kernel stages per chunk
constexpr int kGramChildTab = 256;  // (child, tile) offsets it tabulates
constexpr int kGramFusedTab = 64;   // children whose row blocks it tabulates
constexpr int kLlwThreads = 128;
constexpr int kLlwMaxP = 1024;  // parent-set rows the LLW kernel stages per warp (checked at st_create)
constexpr double kHl2pi = -0.91893853320467274178;  // -0.5 * log(2 pi)  (spamtree_model.h:20)

// row stride of a stored m x m Ri tile with `cols` columns: even (16-byte rows) and = 2 mod 4
__host__ __device__ inline int tile_rs(int cols) {
  int r = (cols + 1) & ~1;
  if ((r & 3) == 0) r += 2;
  return r;
}
// Row stride of a block's row block of the chain's inverse Cholesky factor: [ G (P) | -Ri (m, reference blocks) | 0 ].
// (tree_utils.cpp:204-206 stores [-Ri H | Ri]; the sign is flipped here so that consumers read G directly.)
__host__ __device__ inline int g_stride(int P, int m, int isref) { return (P + (isref ? m : 0) + 3) & ~3; }
// shared-memory layout of a block's m x m Schur complement inside build_level_kernel: stride = 4 mod 8, rows padded to 8
__host__ __device__ inline int rb_stride(int m) {
  int r = (m + 3) & ~3;
  if ((r & 7) == 0) r += 4;
  return r;
}
__host__ __device__ inline int rb_doubles(int m) { return ((m + 7) & ~7) * rb_stride(m); }

// ---- shared-memory plan of build_level_kernel (one work group = a run of sibling blocks)
struct BuildPlan {
  int Ppad, NCp, NT, LD, SA, slot, ring;
  size_t o_panel, o_ring, o_pxs, o_pys, o_wpa, o_cxs, o_cys, o_wcol, o_tvec, o_gw, o_rdiag, o_rowsrc, o_colbase, o_vtmp, o_scr,
      o_rowlen, o_cq, o_colnode, total;
  int npair;
};
// P parent rows, ncols rows in the group's blocks, sumRb = sum of rb_doubles over its blocks (reference levels),
// maxmd = largest block, ns = ring depth (1 or 2), nchol = warps factorising at once (reference levels: min(blocks, 16)),
// nwarps = warps of the CTA
__host__ __device__ inline int build_npair(int NT, int nwarps) { const int x = nwarps - NT; return x < 0 ? 0 : (x < NT ? x : NT); }
__host__ __device__ inline BuildPlan build_plan(int P, int ncols, int sumRb, int maxmd, int ns, int nchol, int nwarps) {
  BuildPlan p;
  p.Ppad = (P + 15) & ~15;
  p.NCp = (ncols + 7) & ~7;
  p.NT = p.NCp >> 3;
  p.LD = p.NCp + 4;   // = 4 or 12 mod 16: the 4 x 8 and 8 x 4 DMMA fragment reads hit 16 distinct banks per half-warp
  p.SA = p.Ppad + 4;  // = 4 mod 16, same reason
  const int fw = kBuildRS * p.SA, bw = p.Ppad * kBuildST;
  p.slot = fw > bw ? fw : bw;
  p.ring = ns * p.slot > sumRb ? ns * p.slot : sumRb;
  size_t o = 0;
  auto take = [&](size_t n_doubles) { size_t r = o; o += ((n_doubles + 1) & ~(size_t)1); return r; };
  p.o_panel = take((size_t)p.Ppad * p.LD);
  p.o_ring = take((size_t)p.ring);
  p.npair = build_npair(p.NT, nwarps);  // column tiles whose reduction range is shared by two warps
  {
    // one scratch region, three lives: parent coordinates (covariance phase: x, y, outcome), partial sums of the paired
    // warps (sweeps: 128 doubles per pair), pivot columns and 1/diag of the factorising warps (2 x 64 + 32 doubles each)
    size_t a = 2 * (size_t)p.Ppad + (p.Ppad + 1) / 2, b = (size_t)p.npair * 128, c = (size_t)nchol * 160;
    if (b > a) a = b;
    if (c > a) a = c;
    p.o_scr = take(a);
  }
  p.o_pxs = p.o_scr; p.o_pys = p.o_scr + p.Ppad; p.o_wpa = take(p.Ppad);
  p.o_cxs = take(p.LD); p.o_cys = take(p.LD); p.o_wcol = take(p.LD); p.o_tvec = take(p.LD); p.o_gw = take(p.LD);
  p.o_rdiag = take(p.LD);
  p.o_rowsrc = take(p.Ppad); p.o_colbase = take(p.LD);
  p.o_vtmp = take(maxmd > 32 ? (size_t)(kBuildMaxThreads / 32) * (maxmd + 2) : 0);
  p.o_rowlen = take((p.Ppad + 1) / 2);
  p.o_cq = take((p.LD + 1) / 2); p.o_colnode = take((p.LD + 1) / 2);
  p.total = o * 8 + 16;
  return p;
}
inline size_t gibbs_smem_bytes(int is_ref, int m, int P, int k) {
  const size_t msq = is_ref ? (size_t)m * tile_rs(m) + (size_t)m * m : 2 * (size_t)m;
  return 8 * (msq + (size_t)P + (size_t)(k + 1) * m + 3 * (size_t)m) + 16;
}

// Opts a kernel in to `smem` bytes of dynamic shared memory on the CURRENT device if it has not been opted in to at
// least that much there yet (the attribute is per device; static + dynamic together exceed the default 48 KB long before
// the dynamic part alone does, so the opt-in is unconditional).  `table` is the caller's per-kernel cache.
struct SmemOptIn { size_t configured[64] = {}; };
template <class K>
inline cudaError_t ensure_dynamic_smem(K kern, size_t smem, SmemOptIn& table) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= 64) return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (smem <= table.configured[dev]) return cudaSuccess;
  e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess) table.configured[dev] = smem;
  return e;
}

// Slot arguments: `D` holds both theta-slots and the device chain state; `rel` = 0 selects param_data, 1 alter_data
// (the kernel reads chain->cur).  `run_flag` (or NULL): device int, the launch is a no-op when it is 0.
// mode 0 / 1: reference / non-reference level of the slot; mode 2: prediction blocks (predG = Hpred receives H, predRi = sd).
cudaError_t launch_build(int mode, const DevTree& T, const DevSlots& D, int rel, double* predG, double* predRi, int want_H,
                         const int* grp_slot0, const int* grp_nn, int ngrp, const double* w, int* fail, int ns, int phase, size_t smem,
                         cudaStream_t st, int nthreads, unsigned long long* prof = nullptr, bool pdl = false,
                         const int* run_flag = nullptr);  // pdl: programmatic dependent launch on the previous kernel of the stream
cudaError_t launch_gibbs(int is_ref, const DevTree& T, const DevSlots& D, int slot0, int nslots, double* w,
                         const double* xb, const double* z, const double* tausq_inv, const double* SigS, double* V,
                         double* probe_sig, double* probe_smu, int* fail, size_t smem, cudaStream_t st, bool pdl = false);
cudaError_t launch_gram(const DevTree& T, const DevSlots& D, int slot0, int nslots, double* U, double* SigS, int rch,
                        int ldx, int tile_doubles, int stage_off, int threads, cudaStream_t st, const int* run_flag = nullptr);
// blocks [slot0, slot0 + nslots) of the slot
cudaError_t launch_llw(const DevTree& T, const DevSlots& D, int rel, int slot0, int nslots, const double* w, int maxlen, cudaStream_t st);
// out8[0..2] = {sum logdet + sum llcomp, sum logdet, 0} over blocks [0, n_top) and out8[4..6] = the same over [n_top, n) with
// out8[6] = *fail (or 0): the two parts a partitioned run needs (replicated blocks once, the rank's own all-reduced)
// scratch: kReduceScratch doubles (partial sums + two counters that must start at zero), private to the stream
constexpr int kReduceScratch = 4 * 32 + 2;
cudaError_t launch_loglik_reduce(const DevSlots& D, int rel, int n_top, int n, const int* fail, double* out8, double* scratch,
                                 cudaStream_t st);
cudaError_t launch_frontier_sum(const DevTree& T, int n, const int* pseudo, const int* c0, const int* c1, const int* vlen,
                                const int* ulen, double* V, double* U, int do_v, int do_u, cudaStream_t st);
cudaError_t launch_predict_sample(const DevTree& T, int slot0, int nslots, const double* Hpred, const double* sdpred,
                                  double* w, const double* z, cudaStream_t st);
// z[i] = N(0,1) from Philox keyed by (seed, rowkey[i], counter + *iter_ptr): rowkey = the row's id in the whole problem, so
// that a partitioned run draws the same number for the same row whatever the number of ranks
cudaError_t launch_normals(double* z, long long n, uint64_t seed, uint64_t counter, const long long* rowkey, const int* iter_ptr,
                           cudaStream_t st);
cudaError_t launch_rowstats(const DevTree& T, const int* widx, long long n_all, int p, int q, const double* w,
                            const double* xb, double* partial, int nblocks, double* out, cudaStream_t st);
cudaError_t launch_xb(const DevTree& T, long long n_all, int p, const double* bcoeff, double* xb, cudaStream_t st);
cudaError_t launch_permute(const double* src, double* dst, const long long* map, long long n, cudaStream_t st);
cudaError_t launch_crosscov(const double* x1, const double* y1, const int* q1, long long n1, const double* x2,
                            const double* y2, const int* q2, long long n2, const CovTab& tab, double* out,
                            cudaStream_t st);
// ---- the device-resident chain (st_chain.hpp): one tiny kernel per step of spamtree_fit.cpp:203-289 / :376-389
// proposal: U ~ N(0, I) (Philox), theta' = back(fwd(theta) + paramsd U) into the alter slot's theta, its covariance table
cudaError_t launch_mh_propose(ChainDev* C, cudaStream_t st);
// accept step.  mode 0: Metropolis rule (Jacobian, uniform draw, RAM adaptation); 1: take the proposal if BUILD succeeded;
// 2: reject (1 / 2: bench hooks).  have_llw: red_llw holds the log-density of the current slot at the new w.
cudaError_t launch_mh_accept(ChainDev* C, int mode, int have_llw, cudaStream_t st);
// need_update of spamtree_fit.cpp:300 -> C->predict_build; predict_param <- theta
cudaError_t launch_predict_gate(ChainDev* C, cudaStream_t st);
// gibbs_sample_tausq (:1393-1417) then gibbs_sample_beta (:1364-1391) from the statistics of rowstats_kernel, random
// numbers from Philox; scratch: q * 3 * p * p doubles
cudaError_t launch_tausq_beta(ChainDev* C, const double* stats, const double* xtx, double* tausq_inv, double* bcoeff,
                              double* scratch, int sample_tausq, int sample_beta, cudaStream_t st);
// the saved iteration's theta / beta / tausq into the device sample arrays (layouts of st_mcmc_out), then C->msaved++
cudaError_t launch_record(ChainDev* C, const double* tausq_inv, const double* bcoeff, double* theta_mcmc, double* beta_mcmc,
                          double* tausq_mcmc, int keep, cudaStream_t st);
// yhat = XB + w + tausq^(1/2) N(0,1) in boundary order (spamtree_fit.cpp:384)
cudaError_t launch_yhat(const DevTree& T, const double* w, const double* xb, const double* tausq_inv, const long long* iperm,
                        const long long* rowkey, long long n, const ChainDev* C, double* out, cudaStream_t st);
cudaError_t launch_chain_tick(ChainDev* C, cudaStream_t st);  // iter++
// host-driven path: theta and its covariance table into the slot `rel` (0 param_data, 1 alter_data) — passed by value as
// kernel parameters, so no host buffer has to outlive the call and nothing synchronises; flip: accept_make_change
struct ThetaPack {
  int n;
  double theta[kMaxPar];
  CovTab tab;
};
cudaError_t launch_chain_set_theta(ChainDev* C, int rel, const ThetaPack& pack, cudaStream_t st);
cudaError_t launch_chain_flip(ChainDev* C, cudaStream_t st);

}  // namespace st
