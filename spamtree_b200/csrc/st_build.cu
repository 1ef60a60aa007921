// st_build.cu — build_level_kernel: BUILD of one tree level (get_loglik_comps_w_std, spamtree_model.cpp:834-998) and
// of the prediction blocks (predict_std, :1234-1358).
//
// One CTA per work group = a run of blocks that share their ancestor chain (siblings), or share all but the deepest
// ancestor (cousins; the deepest ancestor is then "family specific").  Per group, entirely in shared memory:
//   1. K = K_{pa,u}: cross-covariance panel, one column per row of the group's blocks (Covariancef_inplace, :885)
//   2. Z = L^-1 K : the chain's inverse Cholesky factor is never stored; its block row j is [-G_a | Ri_a] of ancestor
//                   a = chain[j] (tree_utils.cpp:204-206), so Z is formed tile by tile, deepest ancestor first, in place
//   3. R = K_uu - Z'Z (= Kcc - H Kxc, :896-897), Ri = chol(R)^-1 in place (non-reference levels: diagonal only, :931-948)
//   4. H' = L^-T Z, shallowest ancestor first, in place  (H = Kxc' Kxx_inv, :887, without ever forming Kxx_inv)
//   5. G = Ri H, e = w - H w_pa, wcore, logdet (:888, :912-913, :966-968)
// The operand tiles (G_a, Ri_a of the ancestors) stream from global/L2 through a 3-stage cp.async ring so that the
// register-tiled FP64 FMA loops read both operands from shared memory.
#include <cuda_pipeline.h>
#include <cuda_runtime.h>

#include <cstdint>

#include "st_device.cuh"
#include "st_kernels.cuh"

namespace st {

namespace {
constexpr int TR = kBuildTR, TC = kBuildTC;
constexpr int F_RI = 1, F_TRANS = 2, F_PERFAM = 4, F_FIRST = 8, F_LAST = 16;
constexpr int kDescInts = 6;  // flags, chain index of the operand's owner (-1: family parent), tile index, brow, orow0, ochain
}  // namespace

// acc[tr][tc] (+/-)= A(tr, kk) * B(kk, tc) for kk in [0, n).  Both operands are in shared memory.
// mac_rows: A(tr, kk) = pa[tr][kk] (tile rows, stride 1 along kk); mac_cols: A(tr, kk) = pt[kk * rs + tr] (transposed use)
template <bool NEG>
__device__ __forceinline__ void mac_rows(double (&acc)[TR][TC], const double* (&pa)[TR], const double* __restrict__ Bp,
                                         int LD, int n) {
#pragma unroll 4
  for (int kk = 0; kk < n; kk++) {
    const double2 b01 = *reinterpret_cast<const double2*>(Bp);
    const double2 b23 = *reinterpret_cast<const double2*>(Bp + 2);
    Bp += LD;
#pragma unroll
    for (int tr = 0; tr < TR; tr++) {
      const double a = NEG ? -pa[tr][kk] : pa[tr][kk];
      acc[tr][0] = fma(a, b01.x, acc[tr][0]);
      acc[tr][1] = fma(a, b01.y, acc[tr][1]);
      acc[tr][2] = fma(a, b23.x, acc[tr][2]);
      acc[tr][3] = fma(a, b23.y, acc[tr][3]);
    }
  }
}
template <bool NEG>
__device__ __forceinline__ void mac_cols(double (&acc)[TR][TC], const double* __restrict__ pt, int rs,
                                         const double* __restrict__ Bp, int LD, int n) {
#pragma unroll 4
  for (int kk = 0; kk < n; kk++) {
    const double2 b01 = *reinterpret_cast<const double2*>(Bp);
    const double2 b23 = *reinterpret_cast<const double2*>(Bp + 2);
    Bp += LD;
#pragma unroll
    for (int tr = 0; tr < TR; tr++) {
      const double a = NEG ? -pt[tr] : pt[tr];  // rows past the block's end read padding; their accumulators are never stored
      acc[tr][0] = fma(a, b01.x, acc[tr][0]);
      acc[tr][1] = fma(a, b01.y, acc[tr][1]);
      acc[tr][2] = fma(a, b23.x, acc[tr][2]);
      acc[tr][3] = fma(a, b23.y, acc[tr][3]);
    }
    pt += rs;
  }
}

template <int MODE>
__global__ void __launch_bounds__(kBuildThreads)
build_level_kernel(DevTree T, DevSlot S, double* __restrict__ outH, double* __restrict__ outRi,
                   const int* __restrict__ grp_slot0, const int* __restrict__ grp_nn, const int* __restrict__ grp_share,
                   const double* __restrict__ w, CovTab tab, int* __restrict__ fail, int keep_H,
                   unsigned long long* __restrict__ prof) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ CovTabS ct;
  __shared__ int s_chain[kMaxChain], s_cm[kMaxChain], s_crow[kMaxChain + 1], s_crow0g[kMaxChain];
  __shared__ int s_nm[kMaxGroupNodes], s_nc0[kMaxGroupNodes], s_nR0[kMaxGroupNodes], s_nfam[kMaxGroupNodes];
  __shared__ int s_fpar[kMaxFam], s_fm[kMaxFam], s_fc0[kMaxFam + 1], s_frow0g[kMaxFam];
  __shared__ int s_npar[kMaxGroupNodes], s_nrow0[kMaxGroupNodes], s_rspref[kMaxChain + 1];
  __shared__ long long s_cgoff[kMaxChain], s_crioff[kMaxChain], s_fgoff[kMaxFam], s_frioff[kMaxFam];
  __shared__ long long s_ngoff[kMaxGroupNodes], s_nrioff[kMaxGroupNodes];
  __shared__ BuildShape sh;
  __shared__ BuildPlan pl;
  __shared__ int s_nfwd;

  const int tid = threadIdx.x, nth = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nth >> 5;
  const int s0 = grp_slot0[blockIdx.x], nn = grp_nn[blockIdx.x];
  const int k = T.k[s0];

  long long tprev = clock64();
  auto mark = [&](int ph) {
    if (prof && tid == 0) { const long long t = clock64(); atomicAdd(prof + ph, (unsigned long long)(t - tprev)); tprev = t; }
  };
  load_covtab(ct, tab);
  // ---- setup: metadata of the chain and of the group's blocks, loaded in parallel (no dependent global loads later)
  {
    const int coff = T.chain_off[s0];
    for (int j = tid; j < k; j += nth) {
      const int a = T.chain[coff + j];
      s_chain[j] = a; s_cm[j] = T.m[a]; s_crow[j] = T.chain_poff[coff + j]; s_crow0g[j] = T.row0[a];
      s_cgoff[j] = T.goff[a]; s_crioff[j] = T.rioff[a];
    }
    for (int d = tid; d < nn; d += nth) {
      s_nm[d] = T.m[s0 + d]; s_npar[d] = T.lastpar[s0 + d]; s_nrow0[d] = T.row0[s0 + d];
      s_ngoff[d] = T.goff[s0 + d]; s_nrioff[d] = T.rioff[s0 + d];
    }
  }
  __syncthreads();
  if (tid == 0) {
    const int share = grp_share[blockIdx.x];
    const int kc = share ? k - 1 : k;
    const int Pc = (kc < k) ? s_crow[kc] : T.P[s0];
    s_crow[kc] = Pc;
    int F = 0, c = 0, sumR = 0, maxmd = 1, prevpar = -2;
    for (int d = 0; d < nn; d++) {
      const int par = s_npar[d], md = s_nm[d];
      if (d == 0 || (share && par != prevpar)) {
        c = (c + 3) & ~3;
        s_fpar[F] = par;
        s_fc0[F] = c;
        F++;
        prevpar = par;
      }
      s_nfam[d] = F - 1;
      s_nc0[d] = c;
      s_nR0[d] = sumR;
      c += md;
      sumR += md * tile_rs(md);
      maxmd = max(maxmd, md);
    }
    const int NCp = (c + 3) & ~3;
    s_fc0[F] = NCp;
    sh.mode = MODE; sh.share = share; sh.kc = kc; sh.Pc = Pc; sh.F = F; sh.NCp = NCp; sh.sumR = sumR; sh.maxmd = maxmd;
    int acc_rs = 0;
    for (int j = 0; j <= kc && j < kMaxChain; j++) { s_rspref[j] = acc_rs; if (j < kc) acc_rs += tile_rs(s_cm[j]); }
  }
  __syncthreads();
  if (tid < sh.F) {
    const int par = s_fpar[tid];
    const bool sp = sh.share != 0;
    s_fm[tid] = sp ? T.m[par] : 0;
    s_frow0g[tid] = sp ? T.row0[par] : 0;
    s_fgoff[tid] = sp ? T.goff[par] : 0;
    s_frioff[tid] = sp ? T.rioff[par] : 0;
  }
  __syncthreads();
  if (tid == 0) {
    int maxm = 1, mmaxs = 0;
    for (int j = 0; j < sh.kc; j++) maxm = max(maxm, s_cm[j]);
    for (int f = 0; f < sh.F; f++) mmaxs = max(mmaxs, s_fm[f]);
    maxm = max(maxm, mmaxs);
    sh.mmaxs = mmaxs;
    sh.maxtile = maxm * tile_rs(maxm);
    pl = build_plan(sh);
    s_nfwd = (sh.share ? sh.kc + 1 : 0) + sh.kc * (sh.kc + 1) / 2;
  }
  __syncthreads();
  const int share = sh.share, kc = sh.kc, Pc = sh.Pc, mmaxs = sh.mmaxs, F = sh.F, NCp = sh.NCp, maxtile = sh.maxtile;
  const int LD = pl.LD, Ppad = pl.Ppad, nsets = pl.nsets, Fst = pl.Fst, nfwd = s_nfwd;
  double* base = reinterpret_cast<double*>(smem_raw);
  double* panel = base + pl.o_panel;
  double* Rb = base + pl.o_R;
  double* ring = base + pl.o_ring;
  double* pxs = base + pl.o_pxs;
  double* pys = base + pl.o_pys;
  double* wpa = base + pl.o_wpa;
  double* cxs = base + pl.o_cxs;
  double* cys = base + pl.o_cys;
  double* ecol = base + pl.o_ecol;
  double* vtmp = base + pl.o_vtmp;
  int* pq = reinterpret_cast<int*>(base + pl.o_pq);
  int* cq = reinterpret_cast<int*>(base + pl.o_cq);
  int* colnode = reinterpret_cast<int*>(base + pl.o_colnode);
  int* cgfam = reinterpret_cast<int*>(base + pl.o_cgfam);
  int* desc = reinterpret_cast<int*>(base + pl.o_desc);

  // ---- operand-tile descriptors of the whole group, in execution order (forward sweep, then backward sweep)
  if (tid == 0) {
    int n = 0;
    auto put = [&](int flags, int owner, int tidx, int brow, int orow0, int ochain) {
      int* d = desc + n * kDescInts;
      d[0] = flags; d[1] = owner; d[2] = tidx; d[3] = brow; d[4] = orow0; d[5] = ochain;
      n++;
    };
    if (share) {  // Z of the family-specific (deepest) ancestor
      for (int i2 = 0; i2 < kc; i2++) put(F_PERFAM | (i2 == 0 ? F_FIRST : 0), -1, i2, s_crow[i2], Pc, -1);
      put(F_RI | F_PERFAM | F_LAST | (kc == 0 ? F_FIRST : 0), -1, 0, Pc, Pc, -1);
    }
    for (int j = kc - 1; j >= 0; j--) {
      for (int i2 = 0; i2 < j; i2++) put(i2 == 0 ? F_FIRST : 0, j, i2, s_crow[i2], s_crow[j], j);
      put(F_RI | F_LAST | (j == 0 ? F_FIRST : 0), j, 0, s_crow[j], s_crow[j], j);
    }
    for (int i = 0; i < kc; i++) {
      const bool more = (i + 1 < kc) || share;
      put(F_RI | F_TRANS | F_FIRST | (more ? 0 : F_LAST), i, 0, s_crow[i], s_crow[i], i);
      for (int j = i + 1; j < kc; j++) put(F_TRANS | ((j + 1 == kc && !share) ? F_LAST : 0), j, i, s_crow[j], s_crow[i], i);
      if (share) put(F_TRANS | F_PERFAM | F_LAST, -1, i, Pc, s_crow[i], i);
    }
    if (share) put(F_RI | F_TRANS | F_PERFAM | F_FIRST | F_LAST, -1, 0, Pc, Pc, -1);
  }

  // ---- phase 1: coordinates (ancestor rows are contiguous in the node-major layout)
  for (int j = 0; j < kc; j++) {
    const int r0 = s_crow0g[j], po = s_crow[j];
    for (int t = tid; t < s_cm[j]; t += nth) {
      pxs[po + t] = T.cx[r0 + t]; pys[po + t] = T.cy[r0 + t]; pq[po + t] = T.mvq[r0 + t]; wpa[po + t] = w[r0 + t];
    }
  }
  if (share)
    for (int f = 0; f < F; f++) {
      const int r0 = s_frow0g[f], po = Pc + f * mmaxs;
      for (int t = tid; t < s_fm[f]; t += nth) {
        pxs[po + t] = T.cx[r0 + t]; pys[po + t] = T.cy[r0 + t]; pq[po + t] = T.mvq[r0 + t]; wpa[po + t] = w[r0 + t];
      }
    }
  for (int c = tid; c < LD; c += nth) { cxs[c] = 0; cys[c] = 0; cq[c] = 0; colnode[c] = -1; }
  for (int cg = tid; cg < NCp / 4; cg += nth) {
    int f = 0;
    while (f + 1 < F && 4 * cg >= s_fc0[f + 1]) f++;
    cgfam[cg] = f;
  }
  __syncthreads();
  for (int d = 0; d < nn; d++) {
    const int r0 = s_nrow0[d], c0 = s_nc0[d];
    for (int t = tid; t < s_nm[d]; t += nth) {
      cxs[c0 + t] = T.cx[r0 + t]; cys[c0 + t] = T.cy[r0 + t]; cq[c0 + t] = T.mvq[r0 + t]; colnode[c0 + t] = d;
    }
  }
  __syncthreads();

  mark(0);
  // ---- cp.async ring over the operand tiles
  auto tile_geom = [&](const int* d, int f, int& slot, int& rows, int& cols) {
    if (d[0] & F_PERFAM) { slot = s_fpar[f]; rows = s_fm[f]; } else { slot = s_chain[d[1]]; rows = s_cm[d[1]]; }
    cols = (d[0] & F_RI) ? rows : s_cm[d[2]];
  };
  auto issue = [&](int s) {
    const int* d = desc + s * kDescInts;
    const int nf = (d[0] & F_PERFAM) ? F : 1;
    double* dst0 = ring + (size_t)(s % kBuildStages) * Fst * maxtile;
    for (int f = 0; f < nf; f++) {
      int slot, rows, cols;
      tile_geom(d, f, slot, rows, cols);
      const bool pf = d[0] & F_PERFAM;
      const double* src = (d[0] & F_RI) ? S.Ri + (pf ? s_frioff[f] : s_crioff[d[1]])
                                        : S.G + (pf ? s_fgoff[f] : s_cgoff[d[1]]) + (long long)rows * s_rspref[d[2]];
      const int n16 = rows * tile_rs(cols) / 2;  // 16-byte chunks (tile sizes are even)
      double* dst = dst0 + (size_t)f * maxtile;
      for (int c = tid; c < n16; c += nth) __pipeline_memcpy_async(dst + 2 * c, src + 2 * c, 16);
    }
  };
  // (each sweep primes the ring itself: between the sweeps the ring doubles as scratch for the triangular inverses)

  // ---- phase 2: covariance panel K_{pa,u} and K_uu
  for (int c = lane; c < LD; c += 32) {
    const bool real = (c < NCp) && colnode[c] >= 0;
    const double xc = cxs[c], yc = cys[c];
    const int qc = cq[c];
    int i = warp;
    for (; i + 3 * nwarps < Pc; i += 4 * nwarps) {  // four independent evaluations in flight per thread
      const int i1 = i + nwarps, i2 = i + 2 * nwarps, i3 = i + 3 * nwarps;
      const double v0 = cov_eval(ct, pxs[i], pys[i], pq[i], xc, yc, qc);
      const double v1 = cov_eval(ct, pxs[i1], pys[i1], pq[i1], xc, yc, qc);
      const double v2 = cov_eval(ct, pxs[i2], pys[i2], pq[i2], xc, yc, qc);
      const double v3 = cov_eval(ct, pxs[i3], pys[i3], pq[i3], xc, yc, qc);
      panel[(size_t)i * LD + c] = real ? v0 : 0.0;
      panel[(size_t)i1 * LD + c] = real ? v1 : 0.0;
      panel[(size_t)i2 * LD + c] = real ? v2 : 0.0;
      panel[(size_t)i3 * LD + c] = real ? v3 : 0.0;
    }
    for (; i < Pc; i += nwarps)
      panel[(size_t)i * LD + c] = real ? cov_eval(ct, pxs[i], pys[i], pq[i], xc, yc, qc) : 0.0;
    if (share) {
      const int f = real ? cgfam[c >> 2] : 0;
      const int mf = real ? s_fm[f] : 0;
      for (int t = warp; t < mmaxs; t += nwarps) {
        const int pi = Pc + f * mmaxs + t;
        panel[(size_t)(Pc + t) * LD + c] = (t < mf) ? cov_eval(ct, pxs[pi], pys[pi], pq[pi], xc, yc, qc) : 0.0;
      }
    }
  }
  if (MODE == 0) {
    for (int d = 0; d < nn; d++) {
      const int md = s_nm[d], c0 = s_nc0[d], rsd = tile_rs(md);
      double* Kuu = Rb + s_nR0[d];
      for (int e = tid; e < md * rsd; e += nth) {
        const int r = e / rsd, r2 = e - r * rsd;
        Kuu[e] = (r2 < md) ? cov_eval(ct, cxs[c0 + r], cys[c0 + r], cq[c0 + r], cxs[c0 + r2], cys[c0 + r2], cq[c0 + r2]) : 0.0;
      }
    }
  } else {
    for (int c = tid; c < LD; c += nth) Rb[c] = (c < NCp && colnode[c] >= 0) ? cov_eval(ct, cxs[c], cys[c], cq[c], cxs[c], cys[c], cq[c]) : 1.0;
  }
  // (the first __syncthreads of the sweep below publishes the panel)

  // ---- phases 3 and 6: one generic sweep over descriptor range [s_begin, s_end).
  // The CTA's threads form two halves: both map to the same register tiles and each half takes one half of every
  // tile's reduction range (split-K), so that all warps issue FP64 work; the halves are summed through the panel.
  const int n_cg = NCp / TC;
  const int half = nth >> 1;
  const int htid = (tid < half) ? tid : tid - half;
  const int hgrp = (tid < half) ? 0 : 1;
  double acc[TR][TC];
  int it_r0 = 0, it_c0 = 0, it_fam = 0, it_orows = 0;
  bool it_active = false;
  auto sweep = [&](int s_begin, int s_end) {
    if (s_begin < s_end) issue(s_begin);
    __pipeline_commit();
    if (s_begin + 1 < s_end) issue(s_begin + 1);
    __pipeline_commit();
    for (int s = s_begin; s < s_end; s++) {
      const int* d = desc + s * kDescInts;
      const int flags = d[0];
      long long tq0 = 0;
      if (prof && tid == 0) tq0 = clock64();
      __pipeline_wait_prior(kBuildStages - 2);
      __syncthreads();
      if (prof && tid == 0) { const long long t = clock64(); atomicAdd(prof + 8, (unsigned long long)(t - tq0)); tq0 = t; }
      if (s + 2 < s_end) issue(s + 2);
      __pipeline_commit();
      if (prof && tid == 0) { const long long t = clock64(); atomicAdd(prof + 9, (unsigned long long)(t - tq0)); tq0 = t; }
      if (flags & F_FIRST) {  // new output block: map one register tile to this thread
        const int omax = (d[5] < 0) ? mmaxs : s_cm[d[5]];
        const int n_rg = (omax + TR - 1) / TR;
        it_active = htid < n_rg * n_cg;
        if (it_active) {
          const int rg = htid % n_rg, cg = htid / n_rg;
          it_r0 = rg * TR;
          it_c0 = cg * TC;
          it_fam = cgfam[cg];
          it_orows = (d[5] < 0) ? s_fm[it_fam] : omax;
          it_active = it_r0 < it_orows;
        }
#pragma unroll
        for (int tr = 0; tr < TR; tr++)
#pragma unroll
          for (int tc = 0; tc < TC; tc++) acc[tr][tc] = 0.0;
      }
      if (it_active) {
        const int fi = (flags & F_PERFAM) ? it_fam : 0;
        int slot, rows, cols;
        tile_geom(d, fi, slot, rows, cols);
        const int rs = tile_rs(cols);
        const double* tile = ring + ((size_t)(s % kBuildStages) * Fst + fi) * maxtile;
        const bool trans = flags & F_TRANS;
        const int K = trans ? rows : cols;
        const int kmid = (K + 1) >> 1;
        const int kk0 = hgrp ? kmid : 0, kk1 = hgrp ? K : kmid;
        const double* Bp = panel + (size_t)(d[3] + kk0) * LD + it_c0;
        // acc = sum G*B - Ri*B ; the block written out is -acc
        if (!trans) {
          const double* pa[TR];
#pragma unroll
          for (int tr = 0; tr < TR; tr++) pa[tr] = tile + min(it_r0 + tr, it_orows - 1) * rs + kk0;
          if (flags & F_RI) mac_rows<true>(acc, pa, Bp, LD, kk1 - kk0);
          else mac_rows<false>(acc, pa, Bp, LD, kk1 - kk0);
        } else {
          const double* pt = tile + kk0 * rs + it_r0;
          if (flags & F_RI) mac_cols<true>(acc, pt, rs, Bp, LD, kk1 - kk0);
          else mac_cols<false>(acc, pt, rs, Bp, LD, kk1 - kk0);
        }
      }
      if (prof && tid == 0) { const long long t = clock64(); atomicAdd(prof + 10, (unsigned long long)(t - tq0)); tq0 = t; }
      if (flags & F_LAST) {  // every thread has finished reading the rows that are about to be overwritten
        __syncthreads();
        if (it_active && hgrp == 1) {
#pragma unroll
          for (int tr = 0; tr < TR; tr++)
            if (it_r0 + tr < it_orows) {
              double* o = panel + (size_t)(d[4] + it_r0 + tr) * LD + it_c0;
              *reinterpret_cast<double2*>(o) = make_double2(-acc[tr][0], -acc[tr][1]);
              *reinterpret_cast<double2*>(o + 2) = make_double2(-acc[tr][2], -acc[tr][3]);
            }
        }
        __syncthreads();
        if (it_active && hgrp == 0) {
#pragma unroll
          for (int tr = 0; tr < TR; tr++)
            if (it_r0 + tr < it_orows) {
              double* o = panel + (size_t)(d[4] + it_r0 + tr) * LD + it_c0;
              const double2 p01 = *reinterpret_cast<const double2*>(o), p23 = *reinterpret_cast<const double2*>(o + 2);
              *reinterpret_cast<double2*>(o) = make_double2(p01.x - acc[tr][0], p01.y - acc[tr][1]);
              *reinterpret_cast<double2*>(o + 2) = make_double2(p23.x - acc[tr][2], p23.y - acc[tr][3]);
            }
        }
        if (prof && tid == 0) { const long long t = clock64(); atomicAdd(prof + 11, (unsigned long long)(t - tq0)); tq0 = t; }
      }
    }
    __syncthreads();
  };

  __syncthreads();
  mark(1);
  sweep(0, nfwd);  // Z = L^-1 K
  mark(2);

  // ---- phase 4/5: Schur complement and its inverse Cholesky factor
  if (MODE == 0) {
    int total = 0;
    for (int d = 0; d < nn; d++) total += ((s_nm[d] + TR - 1) / TR) * ((s_nm[d] + TC - 1) / TC);
    for (int item = tid; item < total; item += nth) {
      int d = 0, rem = item;
      for (;; d++) {
        const int cnt = ((s_nm[d] + TR - 1) / TR) * ((s_nm[d] + TC - 1) / TC);
        if (rem < cnt) break;
        rem -= cnt;
      }
      const int md = s_nm[d], c0d = s_nc0[d], rsd = tile_rs(md);
      const int nrg = (md + TR - 1) / TR;
      const int r0 = (rem % nrg) * TR, c0 = (rem / nrg) * TC;
      double a2[TR][TC];
#pragma unroll
      for (int tr = 0; tr < TR; tr++)
#pragma unroll
        for (int tc = 0; tc < TC; tc++) a2[tr][tc] = 0.0;
      int ro[TR], co[TC];
#pragma unroll
      for (int tr = 0; tr < TR; tr++) ro[tr] = c0d + min(r0 + tr, md - 1);
#pragma unroll
      for (int tc = 0; tc < TC; tc++) co[tc] = c0d + min(c0 + tc, md - 1);
#pragma unroll 2
      for (int pp = 0; pp < Ppad; pp++) {
        const double* row = panel + (size_t)pp * LD;
        double a[TR], b[TC];
#pragma unroll
        for (int tr = 0; tr < TR; tr++) a[tr] = row[ro[tr]];
#pragma unroll
        for (int tc = 0; tc < TC; tc++) b[tc] = row[co[tc]];
#pragma unroll
        for (int tr = 0; tr < TR; tr++)
#pragma unroll
          for (int tc = 0; tc < TC; tc++) a2[tr][tc] = fma(a[tr], b[tc], a2[tr][tc]);
      }
      double* R = Rb + s_nR0[d];
#pragma unroll
      for (int tr = 0; tr < TR; tr++)
#pragma unroll
        for (int tc = 0; tc < TC; tc++)
          if (r0 + tr < md && c0 + tc < md) R[(r0 + tr) * rsd + c0 + tc] -= a2[tr][tc];
    }
    __syncthreads();
    mark(3);
    for (int d = warp; d < nn; d += nwarps) {
      const int md = s_nm[d], rsd = tile_rs(md);
      double* R = Rb + s_nR0[d];
      long long tc0 = 0;
      if (prof && tid == 0) tc0 = clock64();
      bool okc = warp_chol(R, md, rsd, lane);
      if (prof && tid == 0) { const long long t = clock64(); atomicAdd(prof + 12, (unsigned long long)(t - tc0)); tc0 = t; }
      if (okc) {
        if ((size_t)kBuildStages * Fst * maxtile >= (size_t)sh.sumR) {  // the idle ring holds the inverse (column-parallel)
          double* X = ring + s_nR0[d];
          warp_inv_lower_cols(R, X, md, rsd, vtmp + warp * (sh.maxmd + 2), lane);
          for (int e = lane; e < md * rsd; e += 32) R[e] = X[e];
          __syncwarp();
        } else {
          warp_inv_lower_inplace(R, md, rsd, vtmp + warp * (sh.maxmd + 2), lane);
        }
      }
      if (prof && tid == 0) { const long long t = clock64(); atomicAdd(prof + 13, (unsigned long long)(t - tc0)); }
      if (!okc) {
        if (lane == 0) atomicAdd(fail, 1);
        __syncwarp();
        for (int e = lane; e < md * rsd; e += 32) R[e] = 0.0;
      }
    }
    __syncthreads();
  } else {
    for (int c = tid; c < NCp; c += nth) {
      if (colnode[c] < 0) continue;
      double s = 0;
      for (int pp = 0; pp < Ppad; pp++) { const double z = panel[(size_t)pp * LD + c]; s = fma(z, z, s); }
      const double R = Rb[c] - s;
      const bool ok = (R > 0.0) && isfinite(R);
      if (MODE == 1) {
        if (!ok) atomicAdd(fail, 1);
        Rb[c] = ok ? 1.0 / sqrt(R) : 0.0;  // ccholprecdiag (:945)
      } else {
        Rb[c] = ok ? sqrt(R) : 0.0;  // predict_std zeroes the sd when the Cholesky fails (:1316-1322)
      }
    }
    __syncthreads();
  }

  mark(4);
  sweep(nfwd, nsets);  // H' = L^-T Z
  mark(5);

  // ---- phase 7: outputs.  panel[p][c] = H(c, p)
  {
    const int parts = max(1, min(4, nth / max(NCp, 1)));  // threads per column
    for (int c = NCp + tid; c < LD; c += nth) ecol[c] = 0.0;
    for (int c = tid; c < NCp; c += nth) ecol[c] = (colnode[c] >= 0) ? w[s_nrow0[colnode[c]] + (c - s_nc0[colnode[c]])] : 0.0;
    __syncthreads();
    const int c = tid % NCp, part = tid / NCp;
    double s = 0;
    if (part < parts && colnode[c] >= 0) {
      const int lo = (int)((long long)Pc * part / parts), hi = (int)((long long)Pc * (part + 1) / parts);
      double s1 = 0;
      int pp = lo;
      for (; pp + 1 < hi; pp += 2) {
        s = fma(panel[(size_t)pp * LD + c], wpa[pp], s);
        s1 = fma(panel[(size_t)(pp + 1) * LD + c], wpa[pp + 1], s1);
      }
      if (pp < hi) s = fma(panel[(size_t)pp * LD + c], wpa[pp], s);
      s += s1;
      if (share && part == 0) {
        const int f = s_nfam[colnode[c]];
        for (int t = 0; t < s_fm[f]; t++) s = fma(panel[(size_t)(Pc + t) * LD + c], wpa[Pc + f * mmaxs + t], s);
      }
    }
    // fixed-order combination of the partial sums (deterministic): part 0 first, then 1, 2, 3
    for (int q2 = 0; q2 < parts; q2++) {
      if (part == q2 && part < parts && colnode[c] >= 0) ecol[c] -= s;  // e = w_x - H w_pa (:888)
      __syncthreads();
    }
  }
  __syncthreads();
  // G = Ri H (bottom-left block of Kxx_invchol(u), tree_utils.cpp:205) and H, one (block, row group, parent column)
  // item per thread pass, parent column fastest so that the global writes of a warp are contiguous
  {
    int total = 0;
    for (int d = 0; d < nn; d++) total += ((s_nm[d] + TR - 1) / TR) * (Pc + (share ? s_fm[s_nfam[d]] : 0));
    for (int item = tid; item < total; item += nth) {
      int d = 0, rem = item;
      for (;; d++) {
        const int cnt = ((s_nm[d] + TR - 1) / TR) * (Pc + (share ? s_fm[s_nfam[d]] : 0));
        if (rem < cnt) break;
        rem -= cnt;
      }
      const int md = s_nm[d], c0d = s_nc0[d], f = s_nfam[d], rsd = tile_rs(md);
      const int Pd = Pc + (share ? s_fm[f] : 0);
      const int pcol = rem % Pd, r0 = (rem / Pd) * TR;
      int j = 0, mj, prow;
      if (pcol >= Pc) { j = kc; mj = s_fm[f]; prow = Pc; }
      else { while (j + 1 < kc && pcol >= s_crow[j + 1]) j++; mj = s_cm[j]; prow = s_crow[j]; }
      const int pp = pcol - prow, rsj = tile_rs(mj);
      const long long bo = s_ngoff[d] + (long long)md * s_rspref[j] + pp;
      const double* hp = panel + (size_t)(prow + pp) * LD + c0d;
      const double* Rid = Rb + ((MODE == 0) ? s_nR0[d] : c0d);
      if (MODE == 0) {
        double g[TR];
        int ro_[TR];
#pragma unroll
        for (int tr = 0; tr < TR; tr++) { g[tr] = 0.0; ro_[tr] = min(r0 + tr, md - 1) * rsd; }
        const int rmax = min(r0 + TR, md);
#pragma unroll 4
        for (int r2 = 0; r2 < rmax; r2++) {
          const double bv = hp[r2];
#pragma unroll
          for (int tr = 0; tr < TR; tr++) g[tr] = fma(Rid[ro_[tr] + r2], bv, g[tr]);
        }
#pragma unroll
        for (int tr = 0; tr < TR; tr++)
          if (r0 + tr < md) {
            S.G[bo + (size_t)(r0 + tr) * rsj] = g[tr];
            if (keep_H) outH[bo + (size_t)(r0 + tr) * rsj] = hp[r0 + tr];
          }
      } else {
#pragma unroll
        for (int tr = 0; tr < TR; tr++)
          if (r0 + tr < md) {
            const double h = hp[r0 + tr];
            if (MODE == 1) {
              S.G[bo + (size_t)(r0 + tr) * rsj] = Rid[r0 + tr] * h;  // non-reference rows: G_i = H_i / sqrt(R_ii)
              if (keep_H) outH[bo + (size_t)(r0 + tr) * rsj] = h;
            } else {
              outH[bo + (size_t)(r0 + tr) * rsj] = h;
            }
          }
      }
    }
  }
  for (int d = 0; d < nn; d++) {
    const int sd = s0 + d, md = s_nm[d], c0d = s_nc0[d], rsd = tile_rs(md);
    const double* Rid = Rb + ((MODE == 0) ? s_nR0[d] : c0d);
    const long long ro = s_nrioff[d];
    for (int e = tid; e < ((MODE == 0) ? md * rsd : md); e += nth) outRi[ro + e] = Rid[e];
    if (MODE != 2 && warp == (d % nwarps)) {
      // wcore = e' prec e = |Ri e|^2 (:913 / :950) ; logdet = sum log diag(Ri) (:966)
      double wc = 0, ld = 0;
      for (int r = lane; r < md; r += 32) {
        double t;
        if (MODE == 0) {
          t = 0;
          for (int r2 = 0; r2 <= r; r2++) t = fma(Rid[r * rsd + r2], ecol[c0d + r2], t);
          ld += log(Rid[r * rsd + r]);
        } else {
          t = Rid[r] * ecol[c0d + r];
          ld += log(Rid[r]);
        }
        wc = fma(t, t, wc);
      }
      wc = warp_sum(wc);
      ld = warp_sum(ld);
      if (lane == 0) {
        S.logdet[sd] = ld;
        S.llcomp[sd] = (double)md * kHl2pi - 0.5 * wc;  // :967-968
      }
    }
  }
  __syncthreads();
  mark(6);
}

template <int MODE>
static cudaError_t launch_build_t(const DevTree& T, const DevSlot& S, double* outH, double* outRi, const int* grp_slot0,
                                  const int* grp_nn, const int* grp_share, int ngrp, const double* w, const CovTab& tab,
                                  int* fail, int keep_H, size_t smem, cudaStream_t st, unsigned long long* prof, int nthreads) {
  auto kern = build_level_kernel<MODE>;
  static size_t configured[3] = {0, 0, 0};
  if (smem > configured[MODE]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured[MODE] = smem;
  }
  kern<<<ngrp, nthreads, smem, st>>>(T, S, outH, outRi, grp_slot0, grp_nn, grp_share, w, tab, fail, keep_H, prof);
  return cudaGetLastError();
}
cudaError_t launch_build(int mode, const DevTree& T, const DevSlot& S, double* outH, double* outRi, const int* grp_slot0,
                         const int* grp_nn, const int* grp_share, int ngrp, const double* w, const CovTab& tab, int* fail,
                         int keep_H, size_t smem, cudaStream_t st, unsigned long long* prof, int nthreads) {
  if (ngrp <= 0) return cudaSuccess;
  if (mode == 0) return launch_build_t<0>(T, S, outH, outRi, grp_slot0, grp_nn, grp_share, ngrp, w, tab, fail, keep_H, smem, st, prof, nthreads);
  if (mode == 1) return launch_build_t<1>(T, S, outH, outRi, grp_slot0, grp_nn, grp_share, ngrp, w, tab, fail, keep_H, smem, st, prof, nthreads);
  return launch_build_t<2>(T, S, outH, outRi, grp_slot0, grp_nn, grp_share, ngrp, w, tab, fail, keep_H, smem, st, prof, nthreads);
}

}  // namespace st
