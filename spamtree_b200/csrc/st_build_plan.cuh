// st_build_plan.cuh — tiling constants and the shared-memory plan of build_level_kernel (st_build.cu), shared by the kernel
// and the host code that sizes its launches.  Together with st_build.cu and st_device.cuh this file defines the BUILD
// kernels: profiles/*_build_dram.json is stamped with their hash (tools/ncu_build_dram.py, bench.py).
#pragma once
#include <cstddef>

namespace st {

constexpr int kBuildMaxThreads = 512;  // build_level_kernel: one warp per 8 panel columns (at most 16 warps)
constexpr int kBuildRS = 16;     // rows of the chain's inverse Cholesky factor staged per step (two DMMA m-tiles)
constexpr int kBuildST = 20;     // row stride of a backward stage: 16 columns + 4 (conflict-free fragment reads)
constexpr int kMaxGroupCols = 128;
constexpr int kMaxGroupNodes = 32;
constexpr int kMaxChain = 32;

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
}  // namespace st
