// st_kernels.cu — sm_100a kernels of the SpamTrees hot path.
//
//  build_level_kernel   BUILD of one tree level (get_loglik_comps_w_std, spamtree_model.cpp:834-998): per work group
//                       (a run of sibling blocks, which share their ancestor chain) the cross-covariance panel
//                       K_{pa,u} is built in shared memory, pushed through the chain's inverse Cholesky factor
//                       (Z = L^-1 K), reduced to the Schur complement R = K_uu - Z'Z, factorised, and pulled back
//                       (H' = L^-T Z) — all without leaving shared memory.  Also used for the
//                       prediction blocks (predict_std, :1234-1358).
//  gibbs_level_kernel   Gibbs update of w for one level (gibbs_sample_w_std, :1011-1226) including the messages to
//                       every ancestor, accumulated by pulling from the direct children (deterministic, no atomics).
//  gram_level_kernel    theta-only part of the messages (Sigi_children, :1190-1192), refreshed only after theta changes.
//  llw_kernel           log-density at the current theta with the new w (get_loglik_w_std, :781-826).
//  row kernels          beta / tausq sufficient statistics (:1364-1417), X*beta, device normals, predictive draws.
//
// FP64 throughout; B200 has no tcgen05 FP64 kind, the dense contractions are register-tiled DFMA.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>

#include "st_device.cuh"
#include "st_kernels.cuh"

namespace st {

// ------------------------------------------------------------------------------------------------ GIBBS
// One block per tree block (node).  Shared memory: RiS m*m | Sig m*m | wpa P | gwj (k+1)*m | smu m | wn m | rr m
// REF = 1: full m x m conditional (:1037-1089); REF = 0: row-wise scalars (:1091-1155)
template <int REF>
__global__ void __launch_bounds__(kGibbsThreads)
gibbs_level_kernel(DevTree T, DevSlots D, int slot0, double* __restrict__ w, const double* __restrict__ xb,
                   const double* __restrict__ z, const double* __restrict__ tausq_inv, const double* __restrict__ SigS,
                   double* __restrict__ V, double* probe_sig, double* probe_smu, int* __restrict__ fail) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // programmatic dependent launch: the next (shallower) level may start early; it waits below before it reads the
  // messages this level writes (everything before that point depends on ancestors only)
  asm volatile("griddepcontrol.launch_dependents;");
  const DevSlot S = pick_slot(D, D.chain->cur);  // param_data
  const int tid = threadIdx.x, nth = blockDim.x, lane = tid & 31, warp = tid >> 5;
  const int sd = slot0 + blockIdx.x;
  const int m = T.m[sd], k = T.k[sd], P = T.P[sd], coff = T.chain_off[sd], row0 = T.row0[sd];
  const int rsm = tile_rs(m), gs = T.gs[sd];
  const int msq = REF ? m * rsm : m;
  double* RiS = reinterpret_cast<double*>(smem_raw);
  double* Sig = RiS + msq;
  double* wpa = Sig + (REF ? m * m : m);
  double* gwj = wpa + P;  // k tiles of m, then the total at [k*m, (k+1)*m)
  double* smu = gwj + (size_t)(k + 1) * m;
  double* wn = smu + m;
  double* rr = wn + m;
  const double* G = S.G + T.goff[sd];  // the block's rows of the chain factor, [G | -Ri | 0], one contiguous run
  const double* Rig = S.Ri + T.rioff[sd];
  const int nch = T.child_ptr[sd + 1] - T.child_ptr[sd];
  const int* ch = T.child_idx + T.child_ptr[sd];
  // chain metadata once, in parallel (no dependent global loads inside the loops below)
  __shared__ int c_m[32], c_po[32], c_r0[32];
  __shared__ long long c_voff[16];
  __shared__ double cbuf[128];  // pivot columns of the factorisation (2 x 64, double-buffered)
  if (tid < k) {
    const int a = T.chain[coff + tid];
    c_m[tid] = T.m[a]; c_po[tid] = T.chain_poff[coff + tid]; c_r0[tid] = T.row0[a];
  }
  for (int c = tid; c < min(nch, 16); c += nth) c_voff[c] = T.voff[ch[c]];
  for (int e = tid; e < msq; e += nth) RiS[e] = Rig[e];
  // loads that depend on nothing computed here are issued now, so that their latency hides under the phases below:
  // the children's Gram sum (into Sig), and per row tausq_inv (wn), tausq_inv (y - XB) (smu) and the normal draw (rr)
  if (REF) {
    const long long so0 = T.soff[sd];
    for (int e = tid; e < m * m; e += nth) Sig[e] = (so0 >= 0) ? SigS[so0 + e] : 0.0;
  }
  for (int a = tid; a < m; a += nth) {
    const double tq = tausq_inv[T.mvq[row0 + a]];
    wn[a] = tq;
    smu[a] = tq * (T.y[row0 + a] - xb[row0 + a]);
    rr[a] = z[row0 + a];
  }
  __syncthreads();
  for (int e = tid; e < P; e += nth) {
    int j = 0;
    while (j + 1 < k && e >= c_po[j + 1]) j++;
    wpa[e] = w[c_r0[j] + (e - c_po[j])];
  }
  __syncthreads();
  // gwj[j][r] = G_j(r,:) w_{a_j}   (pieces of H w_pa scaled by Ri; :1063 / :1103); ancestor fastest: a warp reads adjacent row segments
  for (int e = tid; e < k * m; e += nth) {
    const int r = e / k, j = e - r * k;
    const int mj = c_m[j], po = c_po[j];
    const double* g = G + (size_t)r * gs + po;
    const double* wv = wpa + po;
    double s[4] = {0, 0, 0, 0};
    int pp = 0;
    for (; pp + 8 <= mj; pp += 8) {  // eight independent global loads in flight per thread (this is the first touch of G: HBM latency)
      double x[8];
#pragma unroll
      for (int u = 0; u < 8; u++) x[u] = g[pp + u];
#pragma unroll
      for (int u = 0; u < 8; u++) s[u & 3] = fma(x[u], wv[pp + u], s[u & 3]);
    }
    if (pp + 4 <= mj) {
      double x[4];
#pragma unroll
      for (int u = 0; u < 4; u++) x[u] = g[pp + u];
#pragma unroll
      for (int u = 0; u < 4; u++) s[u] = fma(x[u], wv[pp + u], s[u]);
      pp += 4;
    }
    for (; pp < mj; pp++) s[0] = fma(g[pp], wv[pp], s[0]);
    gwj[j * m + r] = (s[0] + s[1]) + (s[2] + s[3]);
  }
  __syncthreads();
  for (int r = tid; r < m; r += nth) {
    double s = 0;
    for (int j = 0; j < k; j++) s += gwj[j * m + r];
    gwj[k * m + r] = s;
  }
  __syncthreads();
  const double* gw = gwj + (size_t)k * m;
  if (REF) {
    // Sigi_tot = prec + sum_children + diag(tausq_inv)  (:1044-1051), prec = Ri'Ri (:912)
    for (int e = tid; e < m * m; e += nth) {
      const int a = e / m, b = e - a * m;
      double s = 0;
      for (int r = max(a, b); r < m; r++) s = fma(RiS[r * rsm + a], RiS[r * rsm + b], s);
      s += Sig[e];  // sum over the children, prefetched above
      if (a == b) s += wn[a];
      Sig[e] = s;
      if (probe_sig) probe_sig[T.rioff[sd] + e] = s;
    }
    // The factorisation depends on theta and tausq only — not on the children's messages — so it runs BEFORE the wait on the
    // previous (deeper) level: with programmatic dependent launch the 25-pivot latency chains of consecutive levels overlap
    // instead of queueing behind each other, and only the two triangular solves remain on the level-to-level chain.
    __syncthreads();
    bool okc = true;
    double myinv = 0.0;  // 1 / L(lane, lane)
    if (warp == 0) {
      if (m <= 32) {
        // lane i keeps row i of Sigi_tot in registers, rotated so that the pivot column is a[0]; L overwrites Sig
        double a[32];
#pragma unroll
        for (int j = 0; j < 32; j++) a[j] = (lane < m && j <= lane) ? Sig[lane * m + j] : 0.0;
        cbuf[32 + lane] = 0.0;
        cbuf[96 + lane] = 0.0;
        __syncwarp();
        for (int j = 0; j < m; j++) {
          double d = __shfl_sync(0xffffffffu, a[0], j);
          if (!(d > 0.0) || !isfinite(d)) { okc = false; d = 1.0; }
          const double inv = rsqrt(d), sd = d * inv;
          const double l = (lane == j) ? sd : ((lane > j) ? a[0] * inv : 0.0);
          if (lane == j) myinv = inv;
          if (lane >= j && lane < m) Sig[lane * m + j] = l;
          // pivot column to every lane through shared memory (tools/microbench/chol_bench.cu: 388 cycles per pivot,
          // against 1422 with a shuffle per column)
          double* cc = cbuf + (j & 1) * 64;
          cc[lane] = l;
          __syncwarp();
          const double* cj = cc + j + 1;
#pragma unroll
          for (int i = 0; i < 31; i++) a[i] = fma(-l, cj[i], a[i + 1]);
          a[31] = 0.0;
        }
        __syncwarp();
      } else {
        okc = warp_chol(Sig, m, m, lane);
      }
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");  // the children's messages (V) are read from here on
    // Smu_tot = (H'prec)' w_pa + sum_children + tausq_inv (y - XB)  (:1062-1077)
    for (int a = tid; a < m; a += nth) {
      double s = 0;
      for (int r = a; r < m; r++) s = fma(RiS[r * rsm + a], gw[r], s);
      // a child's vector holds its messages by ancestor column; this block's own columns start at P there (its parent set
      // is this block's chain plus this block) — or at 0 in a limited tree, where the parent set is this block alone
      for (int c = 0; c < nch; c++) s += V[(c < 16 ? c_voff[c] : T.voff[ch[c]]) + (T.limited ? 0 : P) + a];
      s += smu[a];  // tausq_inv (y - XB), prefetched above
      smu[a] = s;
      if (probe_smu) probe_smu[row0 + a] = s;
    }
    __syncthreads();
    if (warp == 0) {
      // w = Sc'(Sc Smu + z), Sc = chol(Sigi_tot)^-1 (:1054, :1086), done as two triangular solves
      if (okc && m <= 32) {
        double b = (lane < m) ? smu[lane] : 0.0;
        for (int j = 0; j < m; j++) {  // forward: L x = Smu
          const double xj = __shfl_sync(0xffffffffu, b * myinv, j);
          if (lane == j) b = xj; else if (lane > j && lane < m) b = fma(-Sig[lane * m + j], xj, b);
        }
        double y = (lane < m) ? b + rr[lane] : 0.0;  // z, prefetched above
        for (int j = m - 1; j >= 0; j--) {  // backward: L' w = x + z
          const double wj = __shfl_sync(0xffffffffu, y * myinv, j);
          if (lane == j) y = wj; else if (lane < j) y = fma(-Sig[j * m + lane], wj, y);
        }
        if (lane < m) { wn[lane] = y; w[row0 + lane] = y; }
      } else if (okc) {
        for (int c = 0; c < m; c++) {  // forward: L x = smu
          __syncwarp();
          const double xc = smu[c] / Sig[c * m + c];
          __syncwarp();
          if (lane == 0) smu[c] = xc;
          for (int r = c + 1 + lane; r < m; r += 32) smu[r] -= Sig[r * m + c] * xc;
        }
        __syncwarp();
        for (int r = lane; r < m; r += 32) smu[r] += rr[r];
        for (int c = m - 1; c >= 0; c--) {  // backward: L' w = x
          __syncwarp();
          const double xc = smu[c] / Sig[c * m + c];
          __syncwarp();
          if (lane == 0) smu[c] = xc;
          for (int r = lane; r < c; r += 32) smu[r] -= Sig[c * m + r] * xc;
        }
        __syncwarp();
        for (int r = lane; r < m; r += 32) { wn[r] = smu[r]; w[row0 + r] = smu[r]; }
      }
      if (!okc) {
        if (lane == 0) atomicAdd(fail, 1);
        for (int r = lane; r < m; r += 32) wn[r] = w[row0 + r];
      }
    }
    __syncthreads();
    for (int r = tid; r < m; r += nth) {
      double s = -gw[r];
      for (int r2 = 0; r2 <= r; r2++) s = fma(RiS[r * rsm + r2], wn[r2], s);
      rr[r] = s;
    }
  } else {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    for (int a = tid; a < m; a += nth) {
      const double ri = RiS[a], prec = ri * ri;
      const double sig = prec + wn[a];                                            // :1123 (tausq_inv prefetched into wn)
      const double sm = ri * gw[a] + smu[a];                                      // :1125-1127
      if (probe_sig) probe_sig[T.rioff[sd] + a] = sig;
      if (probe_smu) probe_smu[row0 + a] = sm;
      double wa;
      if (sig > 0.0 && isfinite(sig)) {
        const double sc = 1.0 / sqrt(sig);
        wa = sc * sc * sm + sc * rr[a];                                           // :1139-1140 (z prefetched into rr)
        w[row0 + a] = wa;
      } else {
        atomicAdd(fail, 1);
        wa = w[row0 + a];
      }
      rr[a] = ri * wa - gw[a];
    }
  }
  __syncthreads();
  // messages to every ancestor (:1158-1207): V_d[J_j] = G_j'(rr + G_j w_{a_j}) + sum over direct children of theirs
  for (int e = tid; e < k * m; e += nth) gwj[e] += rr[e % m];
  __syncthreads();
  double* Vd = V + T.voff[sd];
  // all (ancestor, column) pairs at once: P independent dot products over the block's rows
  for (int e = tid; e < P; e += nth) {
    int j = 0;
    while (j + 1 < k && e >= c_po[j + 1]) j++;
    const int rsj = gs;
    const double* g = G + e;
    const double* vj = gwj + (size_t)j * m;
    double sa[4] = {0, 0, 0, 0};
    int r = 0;
    for (; r + 8 <= m; r += 8) {  // eight independent loads in flight per thread
      double x[8];
#pragma unroll
      for (int u = 0; u < 8; u++) x[u] = g[(size_t)(r + u) * rsj];
#pragma unroll
      for (int u = 0; u < 8; u++) sa[u & 3] = fma(x[u], vj[r + u], sa[u & 3]);
    }
    if (r + 4 <= m) {
      double x[4];
#pragma unroll
      for (int u = 0; u < 4; u++) x[u] = g[(size_t)(r + u) * rsj];
#pragma unroll
      for (int u = 0; u < 4; u++) sa[u] = fma(x[u], vj[r + u], sa[u]);
      r += 4;
    }
    for (; r < m; r++) sa[0] = fma(g[(size_t)r * rsj], vj[r], sa[0]);
    double s = (sa[0] + sa[1]) + (sa[2] + sa[3]);
    if (!T.limited)  // (limited trees: the children's messages stop at this block)
      for (int c = 0; c < nch; c++) s += V[(c < 16 ? c_voff[c] : T.voff[ch[c]]) + e];
    Vd[e] = s;
  }
}

cudaError_t launch_gibbs(int is_ref, const DevTree& T, const DevSlots& D, int slot0, int nslots, double* w,
                         const double* xb, const double* z, const double* tausq_inv, const double* SigS, double* V,
                         double* probe_sig, double* probe_smu, int* fail, size_t smem, cudaStream_t st, bool pdl) {
  if (nslots <= 0) return cudaSuccess;
  static SmemOptIn optin[2];
  auto kern = is_ref ? gibbs_level_kernel<1> : gibbs_level_kernel<0>;
  {
    cudaError_t e = ensure_dynamic_smem(kern, smem, optin[is_ref ? 1 : 0]);
    if (e != cudaSuccess) return e;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(nslots);
  cfg.blockDim = dim3(kGibbsThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, T, D, slot0, w, xb, z, tausq_inv, SigS, V, probe_sig, probe_smu, fail);
}

// ------------------------------------------------------------------------------------------------ message Grams
// U_d[j] = G_j'G_j + sum_children U_c[j]  (= sum over the subtree of AK_uP_u_all[J_j, J_j], :1190-1192) and
// SigS_d = sum_children U_c[tile of d]  (= arma::sum(Sigi_children(d), 2), :1049).  One CTA per node.
// Childless children ("fused": the bulk of the tree) never store their own tiles: the parent forms their G_c'G_c
// from their row blocks directly.  The rows (own + fused children) are staged in shared memory in chunks and every
// thread accumulates one 5 x 5 sub-block of the lower triangle of one tile in registers.
__global__ void __launch_bounds__(kGramThreads)
gram_level_kernel(DevTree T, DevSlots D, int slot0, double* __restrict__ U, double* __restrict__ SigS, int rch, int ldx,
                  int stage_off, const int* __restrict__ run_flag) {
  extern __shared__ __align__(16) double gram_smem[];
  // programmatic dependent launch: the next (shallower) level stages its rows and forms its own G'G while this one runs;
  // it waits below, right before it adds the tiles this level stores
  asm volatile("griddepcontrol.launch_dependents;");
  if (run_flag != nullptr && *run_flag == 0) return;  // device-resident chain: only after an accepted proposal
  const DevSlot S = pick_slot(D, D.chain->cur);       // param_data
  __shared__ int t_po[kMaxChain + 1], t_m[kMaxChain + 1], t_item0[kMaxChain + 2], t_uo[kMaxChain + 1], t_to[kMaxChain + 2];
  __shared__ int s_rows;
  __shared__ long long s_rowoff[kGramMaxRows];  // per staged row: offset of the row in S.G and its valid columns
  __shared__ int s_rowncol[kGramMaxRows];
  __shared__ long long s_cu[kGramChildTab];     // per (child, tile): offset of the child's stored tile, -1 when the child is fused
  __shared__ int f_m[kGramFusedTab], f_gs[kGramFusedTab], f_r0[kGramFusedTab + 1];  // fused children: rows, row stride, first staged row
  __shared__ long long f_goff[kGramFusedTab];
  const int tid = threadIdx.x, nth = blockDim.x;
  const int sd = slot0 + blockIdx.x;
  if (T.ufused[sd]) return;
  double* tiles = gram_smem;                 // the assembled tiles (both triangles), tile j at t_to[j]
  double* stage = gram_smem + stage_off;      // rch staged rows of ldx doubles; stage_off = 0: they share the tiles' storage
  const int m = T.m[sd], k = T.k[sd], P = T.P[sd], coff = T.chain_off[sd];
  const int nch = T.child_ptr[sd + 1] - T.child_ptr[sd];
  const int* ch = T.child_idx + T.child_ptr[sd];
  const long long so = T.soff[sd];
  const int ntile = k + (so >= 0 ? 1 : 0);  // the block's own tile exists only through its children
  if (tid <= k) {
    t_po[tid] = (tid < k) ? T.chain_poff[coff + tid] : (T.limited ? 0 : P);  // where the children's rows hold this block's columns
    t_m[tid] = (tid < k) ? T.m[T.chain[coff + tid]] : m;
    t_uo[tid] = (tid < k) ? T.chain_uoff[coff + tid] : 0;
  }
  // the fused children's row blocks, tabulated once (one thread per child: independent loads)
  const bool ftab = nch <= kGramFusedTab;
  if (ftab && tid < nch) {
    const int cc = ch[tid];
    f_m[tid] = T.ufused[cc] ? T.m[cc] : 0;
    f_goff[tid] = T.goff[cc];
    f_gs[tid] = T.gs[cc];
  }
  __syncthreads();
  if (tid == 0) {
    int it = 0, to = 0;
    for (int j = 0; j < ntile; j++) {
      t_item0[j] = it; t_to[j] = to;
      const int nb = (t_m[j] + 4) / 5;
      it += nb * (nb + 1) / 2;
      to += t_m[j] * t_m[j];
    }
    t_item0[ntile] = it; t_to[ntile] = to;
    int rows = m;
    for (int c = 0; c < nch; c++) {
      if (ftab) { f_r0[c] = rows; rows += f_m[c]; }
      else if (T.ufused[ch[c]]) rows += T.m[ch[c]];
    }
    if (ftab) f_r0[nch] = rows;
    s_rows = rows;
  }
  const bool ctab = nch * ntile <= kGramChildTab;
  if (ctab)
    for (int e = tid; e < nch * ntile; e += nth) {
      const int c = e / ntile, jj = e - c * ntile, cc = ch[c];
      // (limited trees: a child carries one tile, this block's; nothing is passed through to the ancestors)
      s_cu[e] = (T.ufused[cc] || (T.limited && jj < k)) ? -1 : T.uoff[cc] + T.chain_uoff[T.chain_off[cc] + (T.limited ? 0 : jj)];
    }
  __syncthreads();
  const int nitems = t_item0[ntile], nrows = s_rows;
  double* Ud = U + T.uoff[sd];
  for (int item0 = 0; item0 < nitems; item0 += nth) {  // (one round unless the chain is very long)
    const int item = item0 + tid;
    const bool act = item < nitems;
    int j = 0, bi = 0, bj = 0;
    if (act) {
      while (item >= t_item0[j + 1]) j++;
      const int rem = item - t_item0[j];
      while ((bi + 1) * (bi + 2) / 2 <= rem) bi++;
      bj = rem - bi * (bi + 1) / 2;
    }
    const int mj = act ? t_m[j] : 1, po = act ? t_po[j] : 0;
    int ca[5], cb[5];
#pragma unroll
    for (int t = 0; t < 5; t++) { ca[t] = po + min(5 * bi + t, mj - 1); cb[t] = po + min(5 * bj + t, mj - 1); }
    double acc[5][5];
#pragma unroll
    for (int x = 0; x < 5; x++)
#pragma unroll
      for (int y = 0; y < 5; y++) acc[x][y] = 0.0;
    // stream the rows through shared memory: row ids 0..m-1 are the block's own (tiles j < k only), then the fused children's
    for (int rbase = 0; rbase < nrows; rbase += rch) {
      const int nr = min(rch, nrows - rbase);
      __syncthreads();
      for (int rr = tid; rr < nr; rr += nth) {  // where does staged row rr live?
        int r = rbase + rr;
        if (r < m) { s_rowoff[rr] = T.goff[sd] + (long long)r * T.gs[sd]; s_rowncol[rr] = P; }
        else if (ftab) {
          int c = 0;
          while (r >= f_r0[c + 1]) c++;  // (children that keep their own tiles have no rows here)
          s_rowoff[rr] = f_goff[c] + (long long)(r - f_r0[c]) * f_gs[c];
          s_rowncol[rr] = T.limited ? m : P + m;
        } else {
          r -= m;
          for (int c = 0; c < nch; c++) {
            const int cc = ch[c];
            if (!T.ufused[cc]) continue;
            const int mc = T.m[cc];
            if (r < mc) { s_rowoff[rr] = T.goff[cc] + (long long)r * T.gs[cc]; s_rowncol[rr] = T.limited ? m : P + m; break; }
            r -= mc;
          }
        }
      }
      __syncthreads();
      {
        const int hl = ldx >> 1;  // 16-byte chunks per staged row; columns past the row's end are zero-filled
        for (int idx = tid; idx < nr * hl; idx += nth) {
          const int rr = idx / hl, x = (idx - rr * hl) * 2;
          const int valid = min(max(s_rowncol[rr] - x, 0), 2);
          asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"((unsigned)__cvta_generic_to_shared(stage + (size_t)rr * ldx + x)),
                       "l"(S.G + s_rowoff[rr] + (valid ? x : 0)), "r"(8 * valid) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
      }
      __syncthreads();
      if (act) {
        const int rb = (j == k) ? max(0, m - rbase) : 0;  // the block's own rows carry no entries of its own tile
        const int re = (T.limited && j < k) ? min(nr, max(0, m - rbase)) : nr;  // limited: the children's rows carry no ancestor tiles
        for (int rr = rb; rr < re; rr++) {
          const double* x = stage + (size_t)rr * ldx;
          double av[5], bv[5];
#pragma unroll
          for (int t = 0; t < 5; t++) { av[t] = x[ca[t]]; bv[t] = x[cb[t]]; }
#pragma unroll
          for (int xx = 0; xx < 5; xx++)
#pragma unroll
            for (int yy = 0; yy < 5; yy++) acc[xx][yy] = fma(av[xx], bv[yy], acc[xx][yy]);
        }
      }
    }
    __syncthreads();  // (the tiles may take the place of the staged rows)
    if (act) {  // both triangles of the tile
      double* tj = tiles + t_to[j];
#pragma unroll
      for (int xx = 0; xx < 5; xx++)
#pragma unroll
        for (int yy = 0; yy < 5; yy++) {
          const int a = 5 * bi + xx, b = 5 * bj + yy;
          if (a < mj && b < mj) { tj[a * mj + b] = acc[xx][yy]; tj[b * mj + a] = acc[xx][yy]; }
        }
    }
  }
  __syncthreads();
  asm volatile("griddepcontrol.wait;" ::: "memory");
  // add the stored tiles of the children that keep theirs, write out (coalesced)
  for (int eg = tid; eg < t_to[ntile]; eg += nth) {
    int j = 0;
    while (eg >= t_to[j + 1]) j++;
    const int e = eg - t_to[j];
    double v = tiles[eg];
    if (ctab) {
      for (int c = 0; c < nch; c++) {
        const long long cu = s_cu[c * ntile + j];
        if (cu >= 0) v += U[cu + e];
      }
    } else {
      for (int c = 0; c < nch; c++) {
        const int cc = ch[c];
        if (!T.ufused[cc] && !(T.limited && j < k)) v += U[T.uoff[cc] + T.chain_uoff[T.chain_off[cc] + (T.limited ? 0 : j)] + e];
      }
    }
    if (j < k) Ud[t_uo[j] + e] = v; else SigS[so + e] = v;
  }
}
cudaError_t launch_gram(const DevTree& T, const DevSlots& D, int slot0, int nslots, double* U, double* SigS, int rch,
                        int ldx, int tile_doubles, int stage_off, int threads, cudaStream_t st, const int* run_flag, bool pdl) {
  if (nslots <= 0) return cudaSuccess;
  const size_t smem = std::max((size_t)tile_doubles, (size_t)stage_off + (size_t)rch * ldx) * sizeof(double);
  static SmemOptIn optin;
  {
    cudaError_t e = ensure_dynamic_smem(gram_level_kernel, smem, optin);
    if (e != cudaSuccess) return e;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(nslots);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, gram_level_kernel, T, D, slot0, U, SigS, rch, ldx, stage_off, run_flag);
}

// ------------------------------------------------------------------------------------------------ LLW
// llcomp[s] = m*hl2pi - 0.5*|Ri w_u - G w_pa|^2  (get_loglik_w_std, :791-813); one warp per node.
// A reference block's rows of the chain factor are [G | -Ri | 0], so its contribution is one streaming product
// |[G | -Ri] [w_pa ; w_u]|^2 over contiguous rows; the warp stages [w_pa ; w_u] in shared memory once.
__global__ void __launch_bounds__(kLlwThreads)
llw_kernel(DevTree T, DevSlots D, int rel, int slot0, int nslots, const double* __restrict__ w, int maxlen,
           double* __restrict__ vrow, int parked) {
  extern __shared__ __align__(16) double llw_smem[];
  const DevSlot S = pick_slot(D, D.chain->cur ^ rel);
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int sd = slot0 + blockIdx.x * (blockDim.x >> 5) + wib;
  if (sd >= slot0 + nslots) return;
  double* wx = llw_smem + (size_t)wib * maxlen;
  const int m = T.m[sd], k = T.k[sd], P = T.P[sd], coff = T.chain_off[sd], row0 = T.row0[sd], gs = T.gs[sd];
  const bool ref = T.isref[sd] != 0;
  const double* G = S.G + T.goff[sd];
  // lane j keeps the metadata of ancestor j (k <= 32)
  int a_m = 0, a_po = 0, a_r0 = 0;
  if (lane < k) {
    const int a = T.chain[coff + lane];
    a_m = T.m[a];
    a_po = T.chain_poff[coff + lane];
    a_r0 = T.row0[a];
  }
  for (int j = 0; j < k; j++) {
    const int mj = __shfl_sync(0xffffffffu, a_m, j), po = __shfl_sync(0xffffffffu, a_po, j), ar0 = __shfl_sync(0xffffffffu, a_r0, j);
    // (parked blocks: the vector is v = L^-1 w_pa, read off the rows of the ancestors: their s = -v, see below)
    for (int t = lane; t < mj; t += 32) wx[po + t] = parked ? -vrow[ar0 + t] : w[ar0 + t];
  }
  const int len = ref ? P + m : P;
  if (ref)
    for (int t = lane; t < m; t += 32) wx[P + t] = w[row0 + t];
  __syncwarp();
  double wc = 0;
  int r = 0;
  for (; r + 1 < m; r += 2) {  // two rows at a time: independent load streams
    const double* g0 = G + (size_t)r * gs;
    const double* g1 = g0 + gs;
    double s0 = 0, s1 = 0;
    for (int c = lane; c < len; c += 32) {
      const double x = wx[c];
      s0 = fma(__ldg(g0 + c), x, s0);
      s1 = fma(__ldg(g1 + c), x, s1);
    }
    s0 = warp_sum(s0);
    s1 = warp_sum(s1);
    if (!ref) {
      const double* Ri = S.Ri + T.rioff[sd];
      if (parked) {  // the row holds Z_r (unscaled): H_r w_pa = Z_r'v, e_r = w_r - H_r w_pa, prec_r = Ri_r^2
        s0 = Ri[r] * (s0 - w[row0 + r]);
        s1 = Ri[r + 1] * (s1 - w[row0 + r + 1]);
      } else {
        s0 -= Ri[r] * w[row0 + r];
        s1 -= Ri[r + 1] * w[row0 + r + 1];
      }
    } else if (vrow != nullptr && lane == 0) {
      vrow[row0 + r] = s0;  // = -(L^-1 [w_pa ; w_u])_r: what the blocks below read as their v
      vrow[row0 + r + 1] = s1;
    }
    wc = fma(s0, s0, wc);
    wc = fma(s1, s1, wc);
  }
  if (r < m) {
    const double* g0 = G + (size_t)r * gs;
    double s0 = 0;
    for (int c = lane; c < len; c += 32) s0 = fma(__ldg(g0 + c), wx[c], s0);
    s0 = warp_sum(s0);
    if (!ref) {
      const double ri = S.Ri[T.rioff[sd] + r];
      s0 = parked ? ri * (s0 - w[row0 + r]) : s0 - ri * w[row0 + r];
    } else if (vrow != nullptr && lane == 0) {
      vrow[row0 + r] = s0;
    }
    wc = fma(s0, s0, wc);
  }
  if (lane == 0) S.llcomp[sd] = (double)m * kHl2pi - 0.5 * wc;
}
cudaError_t launch_llw(const DevTree& T, const DevSlots& D, int rel, int slot0, int nslots, const double* w, int maxlen, cudaStream_t st,
                       double* vrow, int parked) {
  if (nslots <= 0) return cudaSuccess;
  const int wpb = kLlwThreads / 32;
  maxlen = (maxlen + 1) & ~1;
  const size_t smem = (size_t)wpb * maxlen * sizeof(double);
  static SmemOptIn optin;
  {
    cudaError_t e = ensure_dynamic_smem(llw_kernel, smem, optin);
    if (e != cudaSuccess) return e;
  }
  llw_kernel<<<(nslots + wpb - 1) / wpb, kLlwThreads, smem, st>>>(T, D, rel, slot0, nslots, w, maxlen, vrow, parked);
  return cudaGetLastError();
}

// Sum of the per-block log-density pieces (:987-988 / :815-816), in two parts so that a partitioned run can count the
// replicated blocks once and all-reduce the rest: part 0 = blocks [0, n_top) -> out[0..2], part 1 = [n_top, n) -> out[4..6];
// out[.] = {sum(logdet) + sum(llcomp), sum(logdet), 0 / *fail}.  kRedBlocks CTAs per part sum fixed chunks, the last CTA to
// finish adds the partial sums in chunk order: the summation order is fixed, the result deterministic run to run.
constexpr int kRedBlocks = 32;
__global__ void __launch_bounds__(256) loglik_reduce_kernel(DevSlots D, int rel, int n_top, int n, int* __restrict__ fail,
                                                            double* __restrict__ out, double* __restrict__ partial,
                                                            unsigned int* __restrict__ counter) {
  const int nblk = gridDim.x;  // CTAs per part: kRedBlocks, or 1 for a small tree (one CTA sums a part: no second stage to wait for)
  __shared__ double sa[256], sb[256];
  __shared__ bool last;
  const DevSlot S = pick_slot(D, D.chain->cur ^ rel);
  const int part = blockIdx.y, first = part ? n_top : 0, cnt = (part ? n : n_top) - first;
  const int chunk = (cnt + nblk - 1) / nblk;
  const int lo = first + blockIdx.x * chunk, hi = min(lo + chunk, first + cnt);
  double a = 0, b = 0;
  for (int i = lo + threadIdx.x; i < hi; i += 256) { a += S.logdet[i]; b += S.llcomp[i]; }
  sa[threadIdx.x] = a;
  sb[threadIdx.x] = b;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) { sa[threadIdx.x] += sa[threadIdx.x + s]; sb[threadIdx.x] += sb[threadIdx.x + s]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    partial[(part * kRedBlocks + blockIdx.x) * 2] = sa[0];
    partial[(part * kRedBlocks + blockIdx.x) * 2 + 1] = sb[0];
    if (nblk > 1) {
      __threadfence();
      last = atomicAdd(counter + part, 1u) == (unsigned)nblk - 1;
    } else {
      last = true;
    }
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    double ta = 0, tb = 0;
    for (int k = 0; k < nblk; k++) { ta += partial[(part * kRedBlocks + k) * 2]; tb += partial[(part * kRedBlocks + k) * 2 + 1]; }
    double* o = out + 4 * part;
    o[0] = ta + tb;
    o[1] = ta;
    o[2] = 0.0;
    if (part && fail) { o[2] = (double)*fail; *fail = 0; }  // consumed and cleared: the counter is zero between BUILDs
    counter[part] = 0;  // ready for the next launch
  }
}
cudaError_t launch_loglik_reduce(const DevSlots& D, int rel, int n_top, int n, int* fail, double* out8, double* scratch,
                                 cudaStream_t st) {
  // scratch: 4 * kRedBlocks doubles of partial sums followed by two zero-initialised 32-bit counters
  // (the choice depends on the tree only: a handle always sums in the same order)
  loglik_reduce_kernel<<<dim3(n <= 8192 ? 1 : kRedBlocks, 2), 256, 0, st>>>(D, rel, n_top, n, fail, out8, scratch,
                                                            reinterpret_cast<unsigned int*>(scratch + 4 * kRedBlocks));
  return cudaGetLastError();
}

// partition only: sums of the messages of a replicated block's LOCAL children into its pseudo child, which is then
// all-reduced over the ranks (the replicated block pulls from the pseudo child like from any child)
__global__ void frontier_sum_kernel(DevTree T, const int* __restrict__ pseudo, const int* __restrict__ c0,
                                    const int* __restrict__ c1, const int* __restrict__ vlen, const int* __restrict__ ulen,
                                    double* __restrict__ V, double* __restrict__ U, int do_v, int do_u) {
  const int i = blockIdx.x, ps = pseudo[i];
  if (do_v)
    for (int e = threadIdx.x; e < vlen[i]; e += blockDim.x) {
      double s = 0;
      for (int c = c0[i]; c < c1[i]; c++) s += V[T.voff[c] + e];
      V[T.voff[ps] + e] = s;
    }
  if (do_u)
    for (int e = threadIdx.x; e < ulen[i]; e += blockDim.x) {
      double s = 0;
      for (int c = c0[i]; c < c1[i]; c++) s += U[T.uoff[c] + e];
      U[T.uoff[ps] + e] = s;
    }
}
cudaError_t launch_frontier_sum(const DevTree& T, int n, const int* pseudo, const int* c0, const int* c1, const int* vlen,
                                const int* ulen, double* V, double* U, int do_v, int do_u, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  frontier_sum_kernel<<<n, 256, 0, st>>>(T, pseudo, c0, c1, vlen, ulen, V, U, do_v, do_u);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ prediction draws
// w_i = H_i w_pa + sqrt(K_ii - H_i Kxc_i) z_i  (:1306-1326); one thread per row of a prediction block
__global__ void predict_sample_kernel(DevTree T, int slot0, int nslots, const double* __restrict__ Hpred,
                                      const double* __restrict__ sdpred, double* __restrict__ w,
                                      const double* __restrict__ z) {
  const int sd = slot0 + blockIdx.x;
  const int m = T.m[sd], k = T.k[sd], coff = T.chain_off[sd], row0 = T.row0[sd];
  const double* H = Hpred + T.goff[sd];
  const int gs = T.gs[sd];
  const double* sdv = sdpred + T.rioff[sd];
  for (int r = threadIdx.x; r < m; r += blockDim.x) {
    double s = 0;
    for (int j = 0; j < k; j++) {
      const int a = T.chain[coff + j], mj = T.m[a], ar0 = T.row0[a];
      const double* h = H + (size_t)r * gs + T.chain_poff[coff + j];
      for (int pp = 0; pp < mj; pp++) s = fma(h[pp], w[ar0 + pp], s);
    }
    w[row0 + r] = s + sdv[r] * z[row0 + r];
  }
}
cudaError_t launch_predict_sample(const DevTree& T, int slot0, int nslots, const double* Hpred, const double* sdpred,
                                  double* w, const double* z, cudaStream_t st) {
  if (nslots <= 0) return cudaSuccess;
  predict_sample_kernel<<<nslots, 64, 0, st>>>(T, slot0, nslots, Hpred, sdpred, w, z);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ row kernels
// Philox4x32-10 -> two normals per call by Box-Muller
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
  c[0] = hi1 ^ c[1] ^ k0; c[1] = lo1; c[2] = hi0 ^ c[3] ^ k1; c[3] = lo0;
}
__device__ __forceinline__ void philox4x32(uint32_t (&c)[4], uint64_t seed) {
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; r++) { philox_round(c, k0, k1); k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
}
__device__ __forceinline__ double u01(uint32_t hi, uint32_t lo) { return ((((uint64_t)hi << 32 | lo) >> 11) + 0.5) * (1.0 / 9007199254740992.0); }
// one standard normal / one uniform for (seed, key, counter): Philox4x32-10 + Box-Muller
__device__ __forceinline__ double philox_normal(uint64_t seed, unsigned long long key, uint64_t counter) {
  uint32_t c[4] = {(uint32_t)key, (uint32_t)(key >> 32), (uint32_t)counter, (uint32_t)(counter >> 32)};
  philox4x32(c, seed);
  double sn, cs;
  sincospi(2.0 * u01(c[2], c[3]), &sn, &cs);
  return sqrt(-2.0 * log(u01(c[0], c[1]))) * cs;
}
__device__ __forceinline__ double philox_uniform(uint64_t seed, unsigned long long key, uint64_t counter) {
  uint32_t c[4] = {(uint32_t)key, (uint32_t)(key >> 32), (uint32_t)counter, (uint32_t)(counter >> 32)};
  philox4x32(c, seed);
  return u01(c[0], c[1]);
}
// streams of the chain's scalar draws: key = kStream* + index, counter = iteration (row streams use key = row id < 2^40)
constexpr unsigned long long kStreamU = 1ULL << 48, kStreamAccept = 2ULL << 48, kStreamGamma = 3ULL << 48, kStreamBeta = 4ULL << 48;
constexpr uint64_t kCounterYhat = 1ULL << 62;  // yhat noise: same row keys as the Gibbs normals, disjoint counters

__global__ void normals_kernel(double* __restrict__ z, long long n, uint64_t seed, uint64_t counter,
                               const long long* __restrict__ rowkey, const int* __restrict__ iter_ptr) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // keyed by the row's id in the whole problem: the same draw whatever the partition (and on every rank that holds the row)
  z[i] = philox_normal(seed, (unsigned long long)rowkey[i], counter + (iter_ptr ? (uint64_t)*iter_ptr : 0));
}
cudaError_t launch_normals(double* z, long long n, uint64_t seed, uint64_t counter, const long long* rowkey, const int* iter_ptr,
                           cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  normals_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(z, n, seed, counter, rowkey, iter_ptr);
  return cudaGetLastError();
}

// sufficient statistics of the beta and tausq steps over the observed rows, per outcome j:
//   stat[j*(p+1) + a] = sum X(i,a) * (y_i - w[widx_i])   a < p     (X_available' (y_available - w), :1374-1375)
//   stat[j*(p+1) + p] = sum (y_i - XB_i - w_i)^2                   (bcore, :1399-1400)
// Stage 1 writes one partial vector per block, stage 2 adds the partials in block order (deterministic).
__global__ void __launch_bounds__(256) rowstats_kernel(DevTree T, const int* __restrict__ widx, long long n_all, int p,
                                                       int q, const double* __restrict__ w,
                                                       const double* __restrict__ xb, double* __restrict__ partial) {
  __shared__ double red[8];
  const int nv = q * (p + 1);
  double acc[kMaxStats];
  for (int v = 0; v < nv; v++) acc[v] = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_all; i += (long long)gridDim.x * blockDim.x) {
    const int wi = widx[i];
    if (wi < 0) continue;
    const int j = T.mvq[i];
    const double yi = T.y[i];
    const double res = yi - w[wi];
    for (int a = 0; a < p; a++) acc[j * (p + 1) + a] += T.X[i + (size_t)a * n_all] * res;
    const double e = yi - xb[i] - w[i];
    acc[j * (p + 1) + p] += e * e;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int v = 0; v < nv; v++) {
    const double s = warp_sum(acc[v]);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0;
      for (int k2 = 0; k2 < 8; k2++) t += red[k2];
      partial[(size_t)blockIdx.x * nv + v] = t;
    }
    __syncthreads();
  }
}
// one CTA per statistic: the partial sums of the stage-1 blocks in a fixed order (deterministic)
__global__ void __launch_bounds__(128) rowstats_final_kernel(const double* __restrict__ partial, int nblocks, int nv, double* __restrict__ out) {
  __shared__ double red[128];
  const int v = blockIdx.x;
  double s = 0;
  for (int b = threadIdx.x; b < nblocks; b += 128) s += partial[(size_t)b * nv + v];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 64; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[v] = red[0];
}
cudaError_t launch_rowstats(const DevTree& T, const int* widx, long long n_all, int p, int q, const double* w,
                            const double* xb, double* partial, int nblocks, double* out, cudaStream_t st) {
  rowstats_kernel<<<nblocks, 256, 0, st>>>(T, widx, n_all, p, q, w, xb, partial);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  rowstats_final_kernel<<<q * (p + 1), 128, 0, st>>>(partial, nblocks, q * (p + 1), out);
  return cudaGetLastError();
}

// XB(i) = X(i,:) Bcoeff(:, mv_i)  (:1382)
__global__ void xb_kernel(DevTree T, long long n_all, int p, const double* __restrict__ bcoeff, double* __restrict__ xb) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_all) return;
  const int j = T.mvq[i];
  double s = 0;
  for (int a = 0; a < p; a++) s += T.X[i + (size_t)a * n_all] * bcoeff[a + j * p];
  xb[i] = s;
}
cudaError_t launch_xb(const DevTree& T, long long n_all, int p, const double* bcoeff, double* xb, cudaStream_t st) {
  xb_kernel<<<(unsigned)((n_all + 255) / 256), 256, 0, st>>>(T, n_all, p, bcoeff, xb);
  return cudaGetLastError();
}

// gather / scatter between boundary order and node-major order
__global__ void permute_kernel(const double* __restrict__ src, double* __restrict__ dst, const long long* __restrict__ map,
                               long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[map[i]];
}
cudaError_t launch_permute(const double* src, double* dst, const long long* map, long long n, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  permute_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(src, dst, map, n);
  return cudaGetLastError();
}

// CrossCovarianceAG10 (covariance_functions.cpp:301-355): out(i,j), column-major n1 x n2
__global__ void crosscov_kernel(const double* __restrict__ x1, const double* __restrict__ y1, const int* __restrict__ q1,
                                long long n1, const double* __restrict__ x2, const double* __restrict__ y2,
                                const int* __restrict__ q2, long long n2, CovTab tab, double* __restrict__ out) {
  __shared__ CovTabS ct;
  load_covtab(ct, tab);
  __syncthreads();
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n1 * n2) return;
  const long long i = e % n1, j = e / n1;
  out[e] = cov_eval(ct, x1[i], y1[i], q1[i], x2[j], y2[j], q2[j]);
}
cudaError_t launch_crosscov(const double* x1, const double* y1, const int* q1, long long n1, const double* x2,
                            const double* y2, const int* q2, long long n2, const CovTab& tab, double* out,
                            cudaStream_t st) {
  const long long n = n1 * n2;
  if (n <= 0) return cudaSuccess;
  crosscov_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x1, y1, q1, n1, x2, y2, q2, n2, tab, out);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ device-resident chain
// The steps of spamtree_fit.cpp:203-289 and :376-389 that the reference runs on the host between the model-layer calls,
// as one-CTA kernels over the chain state in device memory (st_chain.hpp).  The arithmetic is st_mh.hpp's, the same
// functions the host path (and, through it, the pin against the reference's mh_adapt.h) uses.

// U ~ N(0, I), theta' = back(fwd(theta) + paramsd U) clipped (spamtree_fit.cpp:211-215) -> theta and covariance table of the
// alter slot
// iter_offset = 1: the proposal of the NEXT iteration, drawn right after this iteration's accept step (before the tick)
__global__ void mh_propose_kernel(ChainDev* C, int iter_offset) {
  const int npar = C->npar, cur = C->cur, alt = cur ^ 1;
  for (int j = threadIdx.x; j < npar; j += blockDim.x) C->U[j] = philox_normal(C->seed, kStreamU + j, (uint64_t)(C->iter + iter_offset));
  __syncthreads();
  if (threadIdx.x == 0) {
    mh_propose(npar, C->theta[cur], C->bounds, C->paramsd, C->U, C->theta[alt]);
    make_covtab_hd(C->theta[alt], npar, C->q, C->tab[alt]);
  }
}
cudaError_t launch_mh_propose(ChainDev* C, cudaStream_t st, int iter_offset) {
  mh_propose_kernel<<<1, 64, 0, st>>>(C, iter_offset);
  return cudaGetLastError();
}

// spamtree_fit.cpp:223-285: log-densities from the two reductions, Jacobian, accept decision, slot swap, RAM adaptation
// stage = 1: the adaptation's matrices (paramsd, prodparam, 2 npar^2 of scratch, U) are staged in shared memory — one thread
// runs a chain of dependent read-modify-writes over them, which in global memory is a chain of L2 round trips
__global__ void mh_accept_kernel(ChainDev* C, int mode, int have_llw, int stage) {
  extern __shared__ double acc_smem[];
  const int npar = C->npar, n2 = npar * npar;
  const bool adapt = mode == 0 && C->adapting;
  double *sp = C->paramsd, *pp = C->prodparam, *sc = C->scratch;
  const double* su = C->U;
  if (stage && adapt) {
    sp = acc_smem; pp = sp + n2; sc = pp + n2;
    double* u = sc + 2 * n2;
    for (int e = threadIdx.x; e < n2; e += blockDim.x) { sp[e] = C->paramsd[e]; pp[e] = C->prodparam[e]; }
    for (int e = threadIdx.x; e < npar; e += blockDim.x) u[e] = C->U[e];
    su = u;
    __syncthreads();
  }
  // the Jacobian's terms (four logarithms each) and the uniform draw by separate threads; thread 0 adds the terms in
  // mh_jacobian's order (st_mh.hpp), so the sum is the same
  __shared__ double s_term[kMaxPar];
  __shared__ double s_u;
  if (mode == 0) {
    const int cur0 = C->cur, alt0 = cur0 ^ 1;
    for (int j = threadIdx.x; j < npar; j += blockDim.x) {
      const double lo = C->bounds[j], hi = C->bounds[j + npar], pj = C->theta[cur0][j], nj = C->theta[alt0][j];
      s_term[j] = (-log(hi - pj) - log(pj - lo)) - (-log(hi - nj) - log(nj - lo));
    }
    if (threadIdx.x == blockDim.x - 1) s_u = philox_uniform(C->seed, kStreamAccept, (uint64_t)C->iter);
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const int cur = C->cur, alt = cur ^ 1, m = C->iter;
    if (have_llw) { C->loglik[cur] = C->red_llw[0] + C->red_llw[4]; C->logdet[cur] = C->red_llw[1] + C->red_llw[5]; }
    const bool acceptable = C->red_build[6] == 0.0;
    if (acceptable) {  // on failure the reference leaves the alter slot's log-density untouched (:971-982)
      C->loglik[alt] = C->red_build[0] + C->red_build[4];
      C->logdet[alt] = C->red_build[1] + C->red_build[5];
    } else {
      C->n_chol_fail++;
    }
    const double new_loglik = C->loglik[alt], current_loglik = C->loglik[cur];
    if (isnan(current_loglik)) C->nan_loglik = 1;  // spamtree_fit.cpp:234-237
    bool accepted;
    double logaccept = 0.0;
    if (mode == 0) {
      double jac = 0;
      for (int j = 0; j < npar; j++) jac += s_term[j];
      logaccept = new_loglik - current_loglik + jac;
      const double u = s_u;
      C->last_logaccept = logaccept; C->last_u = u;
      accepted = (u < mh_accept_prob(logaccept)) && acceptable;
    } else {
      accepted = (mode == 1) && acceptable;
    }
    if (accepted) {
      C->n_accepted++;
      C->cur = alt;  // accept_make_change (:1432-1435): the kernels that follow read the new slot
      C->pred_valid = 0;
    }
    C->accepted_now = accepted ? 1 : 0;
    if (adapt)  // :285 (a failed chol(S) keeps the previous factor and is counted)
      if (!ram_adapt(npar, sp, pp, &C->ram_started, su, (acceptable ? 1.0 : 0.0) * exp(logaccept), m, sc))
        C->n_ram_fail++;
  }
  if (stage && adapt) {
    __syncthreads();
    for (int e = threadIdx.x; e < n2; e += blockDim.x) { C->paramsd[e] = sp[e]; C->prodparam[e] = pp[e]; }
  }
}
cudaError_t launch_mh_accept(ChainDev* C, int mode, int have_llw, cudaStream_t st, int npar) {
  const size_t smem = ((size_t)4 * npar * npar + npar) * sizeof(double);
  const int stage = npar > 0 && smem <= 40 * 1024;
  mh_accept_kernel<<<1, 64, stage ? smem : 0, st>>>(C, mode, have_llw, stage);
  return cudaGetLastError();
}

// need_update = any |param - predict_param| > 1e-05 (spamtree_fit.cpp:300); predict_param <- param (:305)
__global__ void predict_gate_kernel(ChainDev* C) {
  if (threadIdx.x != 0) return;
  const double* th = C->theta[C->cur];
  bool need = !C->pred_valid;
  for (int j = 0; j < C->npar; j++) {
    if (fabs(th[j] - C->predict_param[j]) > 1e-05) need = true;
    C->predict_param[j] = th[j];
  }
  C->predict_build = need ? 1 : 0;
  C->pred_valid = 1;
}
cudaError_t launch_predict_gate(ChainDev* C, cudaStream_t st) {
  predict_gate_kernel<<<1, 32, 0, st>>>(C);
  return cudaGetLastError();
}

// Marsaglia-Tsang gamma(shape >= 1, scale) on a Philox stream (the host path's HostRng::gamma, st_common.hpp)
// x0, u0: the first attempt's normal and uniform, drawn by other threads (same streams: same value)
__device__ inline double philox_gamma(uint64_t seed, unsigned long long key, uint64_t counter, double shape, double scale, double x0, double u0) {
  const double d = shape - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
  for (unsigned long long att = 0;; att++) {
    const double x = att ? philox_normal(seed, key + (att << 8), counter) : x0;
    double v = 1.0 + c * x;
    if (v <= 0) continue;
    v = v * v * v;
    const double u = att ? philox_uniform(seed, key + (att << 8) + 128, counter) : u0;
    if (u < 1.0 - 0.0331 * x * x * x * x) return d * v * scale;
    if (log(u) < 0.5 * x * x + d * (1.0 - v + log(v))) return d * v * scale;
  }
}
// gibbs_sample_tausq (spamtree_model.cpp:1393-1417) then gibbs_sample_beta (:1364-1391), one thread per outcome.
// stats[j*(p+1) + a] = X_j'(y - w)_a, stats[j*(p+1) + p] = |y - XB - w|^2 over the outcome's observed rows (rowstats_kernel)
__global__ void tausq_beta_kernel(ChainDev* C, const double* __restrict__ stats, const double* __restrict__ xtx,
                                  double* __restrict__ tausq_inv, double* __restrict__ bcoeff, double* __restrict__ scratch,
                                  int sample_tausq, int sample_beta, int stage) {
  extern __shared__ double tb_smem[];  // stage = 1: the p x p work arrays of every outcome (else `scratch` in global memory)
  if (stage) scratch = tb_smem;
  const int j = threadIdx.x, p = C->p, q = C->q;
  const uint64_t it = (uint64_t)C->iter;
  // every draw by a thread of its own (a normal is a Philox block, a logarithm, a square root and a sincospi: the one thread
  // per outcome that used to draw p + 2 of them in sequence spent most of its time there)
  __shared__ double s_zb[kMaxStats], s_gx[kMaxQ], s_gu[kMaxQ];
  const bool par_draws = q * p <= kMaxStats;
  if (par_draws) {
    const int nz = sample_beta ? q * p : 0, ng = sample_tausq ? q : 0;
    for (int t = threadIdx.x; t < nz + 2 * ng; t += blockDim.x) {
      if (t < nz) s_zb[t] = philox_normal(C->seed, kStreamBeta + ((unsigned long long)(t / p) << 32) + (t % p), it);
      else if (t < nz + ng) s_gx[t - nz] = philox_normal(C->seed, kStreamGamma + ((unsigned long long)(t - nz) << 32), it);
      else s_gu[t - nz - ng] = philox_uniform(C->seed, kStreamGamma + ((unsigned long long)(t - nz - ng) << 32) + 128, it);
    }
    __syncthreads();
  }
  if (j >= q) return;
  if (sample_tausq) {
    const double bcore = stats[j * (p + 1) + p];
    const unsigned long long gk = kStreamGamma + ((unsigned long long)j << 32);
    const double x0 = par_draws ? s_gx[j] : philox_normal(C->seed, gk, it), u0 = par_draws ? s_gu[j] : philox_uniform(C->seed, gk + 128, it);
    tausq_inv[j] = philox_gamma(C->seed, gk, it, 2.01 + C->nobs[j] / 2.0, 1.0 / (1.0 + .5 * bcore), x0, u0);
  }
  if (sample_beta) {
    const double tq = tausq_inv[j];
    double* Si = scratch + (size_t)j * (2 * p * p + 4 * p);  // precision -> its Cholesky factor (column-major)
    double* Sc = Si + p * p;                                 // inverse of the factor
    double* v = Sc + p * p;                                  // xp | t | bmu | sz: 4 p doubles of the outcome's own
    for (int a = 0; a < p; a++)
      for (int b = 0; b < p; b++) Si[a + b * p] = tq * xtx[(size_t)j * p * p + (b <= a ? b + a * p : a + b * p)] + (a == b ? .01 : 0.0);  // symmatu; Vi = .01 I (:157)
    if (!mh_chol_lower(Si, p)) { C->gibbs_fail = 1; return; }
    for (int c = 0; c < p; c++) {  // Sc = Si^-1 (lower)
      for (int r = 0; r < c; r++) Sc[r + c * p] = 0.0;
      Sc[c + c * p] = 1.0 / Si[c + c * p];
      for (int r = c + 1; r < p; r++) {
        double s = 0;
        for (int k = c; k < r; k++) s += Si[r + k * p] * Sc[k + c * p];
        Sc[r + c * p] = -s / Si[r + r * p];
      }
    }
    double *xp = v, *t = v + p, *bmu = v + 2 * p;
    for (int a = 0; a < p; a++) xp[a] = tq * stats[j * (p + 1) + a];  // Vim = 0 (:158-159)
    for (int a = 0; a < p; a++) { double s = 0; for (int b = 0; b <= a; b++) s += Sc[a + b * p] * xp[b]; t[a] = s; }
    for (int a = 0; a < p; a++) { double s = 0; for (int b = a; b < p; b++) s += Sc[b + a * p] * t[b]; bmu[a] = s; }
    for (int a = 0; a < p; a++) xp[a] = par_draws ? s_zb[j * p + a] : philox_normal(C->seed, kStreamBeta + ((unsigned long long)j << 32) + a, it);
    for (int a = 0; a < p; a++) {
      double s = 0;
      for (int b = a; b < p; b++) s += Sc[b + a * p] * xp[b];
      bcoeff[a + j * p] = bmu[a] + s;
    }
  }
}
cudaError_t launch_tausq_beta(ChainDev* C, const double* stats, const double* xtx, double* tausq_inv, double* bcoeff,
                              double* scratch, int sample_tausq, int sample_beta, cudaStream_t st, int p, int q) {
  const size_t smem = (size_t)q * (2 * p * p + 4 * p) * sizeof(double);  // the layout of `scratch`
  const int stage = p > 0 && smem <= 40 * 1024;
  tausq_beta_kernel<<<1, 64, stage ? smem : 0, st>>>(C, stats, xtx, tausq_inv, bcoeff, scratch, sample_tausq, sample_beta, stage);
  return cudaGetLastError();
}

// save (spamtree_fit.cpp:376-382): column msaved of theta_mcmc (npar x keep), tausq_mcmc (q x keep), beta_mcmc (p x keep x q)
__global__ void record_kernel(ChainDev* C, const double* __restrict__ tausq_inv, const double* __restrict__ bcoeff,
                              double* __restrict__ theta_mcmc, double* __restrict__ beta_mcmc, double* __restrict__ tausq_mcmc, int keep,
                              int tick) {
  const int ms = C->msaved, npar = C->npar, p = C->p, q = C->q;
  if (ms < keep) {
    for (int j = threadIdx.x; j < npar; j += blockDim.x) theta_mcmc[j + (size_t)ms * npar] = C->theta[C->cur][j];
    for (int j = threadIdx.x; j < q; j += blockDim.x) tausq_mcmc[j + (size_t)ms * q] = 1.0 / tausq_inv[j];
    for (int e = threadIdx.x; e < p * q; e += blockDim.x) { const int a = e % p, j = e / p; beta_mcmc[a + (size_t)ms * p + (size_t)j * p * keep] = bcoeff[e]; }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    C->msaved = ms + 1;
    if (tick) C->iter++;  // (the end of a saved iteration: chain_tick_kernel's work rides along)
  }
}
cudaError_t launch_record(ChainDev* C, const double* tausq_inv, const double* bcoeff, double* theta_mcmc, double* beta_mcmc,
                          double* tausq_mcmc, int keep, cudaStream_t st, int tick) {
  record_kernel<<<1, 64, 0, st>>>(C, tausq_inv, bcoeff, theta_mcmc, beta_mcmc, tausq_mcmc, keep, tick);
  return cudaGetLastError();
}

// yhat = XB + w + tausq_inv^(-1/2) N(0,1), boundary order (spamtree_fit.cpp:384); iperm: boundary row -> node-major row
__global__ void yhat_kernel(DevTree T, const double* __restrict__ w, const double* __restrict__ xb, const double* __restrict__ tausq_inv,
                            const long long* __restrict__ iperm, const long long* __restrict__ rowkey, long long n, const ChainDev* C,
                            double* __restrict__ out) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n) return;
  const long long i = iperm[b];
  const double e = philox_normal(C->seed, (unsigned long long)rowkey[i], kCounterYhat + (uint64_t)C->iter);
  out[b] = xb[i] + w[i] + rsqrt(tausq_inv[T.mvq[i]]) * e;
}
cudaError_t launch_yhat(const DevTree& T, const double* w, const double* xb, const double* tausq_inv, const long long* iperm,
                        const long long* rowkey, long long n, const ChainDev* C, double* out, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  yhat_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(T, w, xb, tausq_inv, iperm, rowkey, n, C, out);
  return cudaGetLastError();
}

__global__ void chain_set_theta_kernel(ChainDev* C, int rel, const __grid_constant__ ThetaPack P) {
  const int phys = C->cur ^ rel;
  for (int j = threadIdx.x; j < P.n; j += blockDim.x) C->theta[phys][j] = P.theta[j];
  // (the table word by word, by all threads: CovTab is an int followed by doubles, 8-byte aligned)
  const unsigned long long* src = reinterpret_cast<const unsigned long long*>(&P.tab);
  unsigned long long* dst = reinterpret_cast<unsigned long long*>(&C->tab[phys]);
  for (int j = threadIdx.x; j < (int)(sizeof(CovTab) / 8); j += blockDim.x) dst[j] = src[j];
}
cudaError_t launch_chain_set_theta(ChainDev* C, int rel, const ThetaPack& pack, cudaStream_t st) {
  chain_set_theta_kernel<<<1, 64, 0, st>>>(C, rel, pack);
  return cudaGetLastError();
}
__global__ void chain_flip_kernel(ChainDev* C) { C->cur ^= 1; C->pred_valid = 0; }
cudaError_t launch_chain_flip(ChainDev* C, cudaStream_t st) {
  chain_flip_kernel<<<1, 1, 0, st>>>(C);
  return cudaGetLastError();
}
__global__ void chain_tick_kernel(ChainDev* C) { C->iter++; }
cudaError_t launch_chain_tick(ChainDev* C, cudaStream_t st) {
  chain_tick_kernel<<<1, 1, 0, st>>>(C);
  return cudaGetLastError();
}

}  // namespace st
