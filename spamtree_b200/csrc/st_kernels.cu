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

#include <cstdint>

#include "st_kernels.cuh"

namespace st {

// ------------------------------------------------------------------------------------------------ small device helpers
struct CovTabS {  // shared-memory copy of CovTab
  int q;
  double c1[kMaxQ * kMaxQ], r1[kMaxQ * kMaxQ], c2[kMaxQ * kMaxQ], r2[kMaxQ * kMaxQ];
};
__device__ __forceinline__ void load_covtab(CovTabS& s, const CovTab& t) {
  if (threadIdx.x == 0) s.q = t.q;
  for (int i = threadIdx.x; i < t.q * t.q; i += blockDim.x) {
    s.c1[i] = t.c1[i]; s.r1[i] = t.r1[i]; s.c2[i] = t.c2[i]; s.r2[i] = t.r2[i];
  }
}
// mvCovAG20107_inplace / cexpcov: see make_covtab() for how (c1, r1, c2, r2) follow from theta
__device__ __forceinline__ double cov_eval(const CovTabS& t, double x1, double y1, int q1, double x2, double y2, int q2) {
  const double dx = x1 - x2, dy = y1 - y2;
  const double h = sqrt(dx * dx + dy * dy);
  const int ix = q1 * t.q + q2;
  double v = t.c1[ix] * exp(-t.r1[ix] * h);
  const double c2 = t.c2[ix];
  if (c2 != 0.0) v += c2 * exp(-t.r2[ix] * h);
  return v;
}

// in-place lower Cholesky of the symmetric m x m matrix A (row-major, ld = m) by one warp; false on a bad pivot
__device__ bool warp_chol(double* A, int m, int lane) {
  for (int c = 0; c < m; c++) {
    __syncwarp();
    double d = A[c * m + c];
    if (!(d > 0.0) || !isfinite(d)) return false;
    d = sqrt(d);
    const double inv = 1.0 / d;
    __syncwarp();
    for (int r = c + lane; r < m; r += 32) A[r * m + c] = (r == c) ? d : A[r * m + c] * inv;
    __syncwarp();
    for (int r = c + 1 + lane; r < m; r += 32) {
      const double lrc = A[r * m + c];
      for (int c2 = c + 1; c2 <= r; c2++) A[r * m + c2] -= lrc * A[c2 * m + c];
    }
  }
  __syncwarp();
  return true;
}
// X = L^-1 for lower-triangular L (both row-major m x m); the strict upper part of X is zero-filled; one warp
__device__ void warp_inv_lower(const double* L, double* X, int m, int lane) {
  for (int c = lane; c < m; c += 32) {
    for (int r = 0; r < c; r++) X[r * m + c] = 0.0;
    X[c * m + c] = 1.0 / L[c * m + c];
    for (int r = c + 1; r < m; r++) {
      double s = 0;
      for (int kk = c; kk < r; kk++) s += L[r * m + kk] * X[kk * m + c];
      X[r * m + c] = -s / L[r * m + r];
    }
  }
  __syncwarp();
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// acc[tr][tc] += sign * A(tr, kk) * B[kk*LD + tc], kk in [0, K).  A(tr, kk) = Ap[tr] [kk * a_ks] read through the
// read-only global path; B is the shared-memory panel.
template <int TR, int TC>
__device__ __forceinline__ void tile_mac(double (&acc)[TR][TC], const double* const (&Ap)[TR], int a_ks,
                                         const double* __restrict__ Bp, int LD, int K, double sign) {
#pragma unroll 2
  for (int kk = 0; kk < K; kk++) {
    double a[TR], b[TC];
#pragma unroll
    for (int tr = 0; tr < TR; tr++) a[tr] = sign * __ldg(Ap[tr] + (size_t)kk * a_ks);
#pragma unroll
    for (int tc = 0; tc < TC; tc++) b[tc] = Bp[kk * LD + tc];
#pragma unroll
    for (int tr = 0; tr < TR; tr++)
#pragma unroll
      for (int tc = 0; tc < TC; tc++) acc[tr][tc] = fma(a[tr], b[tc], acc[tr][tc]);
  }
}

// ------------------------------------------------------------------------------------------------ BUILD
constexpr int kMaxChain = 32;
constexpr int kMaxGroupNodes = 64;

// MODE 0: reference level, 1: non-reference level (rows conditionally independent, :923-962), 2: prediction blocks
template <int TR, int TC, int MODE>
__global__ void __launch_bounds__(kBuildThreads)
build_level_kernel(DevTree T, DevSlot S, double* __restrict__ outH, double* __restrict__ outRi,
                   const int* __restrict__ grp_slot0, const int* __restrict__ grp_nn, const double* __restrict__ w,
                   CovTab tab, int* __restrict__ fail, int keep_H) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ CovTabS ct;
  __shared__ int s_chain[kMaxChain], s_cm[kMaxChain], s_cpoff[kMaxChain + 1], s_crow0[kMaxChain];
  __shared__ int s_nm[kMaxGroupNodes], s_nc0[kMaxGroupNodes + 1], s_nsq0[kMaxGroupNodes + 1];
  __shared__ int s_maxmj;

  const int tid = threadIdx.x, nth = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nth >> 5;
  const int s0 = grp_slot0[blockIdx.x], nn = grp_nn[blockIdx.x];
  const int k = T.k[s0], P = T.P[s0], coff = T.chain_off[s0];

  load_covtab(ct, tab);
  if (tid == 0) {
    int mx = 1;
    for (int j = 0; j < k; j++) {
      const int a = T.chain[coff + j];
      s_chain[j] = a;
      s_cm[j] = T.m[a];
      s_cpoff[j] = T.chain_poff[coff + j];
      s_crow0[j] = T.row0[a];
      mx = max(mx, s_cm[j]);
    }
    s_cpoff[k] = P;
    int c = 0, sq = 0;
    for (int d = 0; d < nn; d++) {
      const int md = T.m[s0 + d];
      s_nm[d] = md;
      s_nc0[d] = c;
      s_nsq0[d] = sq;
      c += md;
      sq += (MODE == 0) ? md * md : md;
    }
    s_nc0[nn] = c;
    s_nsq0[nn] = sq;
    s_maxmj = mx;
  }
  __syncthreads();
  const int NC = s_nc0[nn], sumsq = s_nsq0[nn], maxmj = s_maxmj;
  const int LD = build_ld(NC);
  const int s1n = build_s1n(MODE, maxmj, LD, sumsq);

  double* panel = reinterpret_cast<double*>(smem_raw);
  double* S1 = panel + (size_t)P * LD;
  double* S2 = S1 + s1n;
  double* pxs = S2 + ((MODE == 0) ? sumsq : LD);
  double* pys = pxs + P;
  double* wpa = pys + P;
  double* cxs = wpa + P;
  double* cys = cxs + LD;
  double* ecol = cys + LD;
  int* pq = reinterpret_cast<int*>(ecol + LD);
  int* cq = pq + P;
  int* colnode = cq + LD;

  // ---- phase 1: coordinates of the parent rows (ancestor panels are contiguous) and of this group's columns
  for (int j = 0; j < k; j++) {
    const int r0 = s_crow0[j], po = s_cpoff[j];
    for (int t = tid; t < s_cm[j]; t += nth) {
      pxs[po + t] = T.cx[r0 + t];
      pys[po + t] = T.cy[r0 + t];
      pq[po + t] = T.mvq[r0 + t];
      wpa[po + t] = w[r0 + t];
    }
  }
  for (int d = 0; d < nn; d++) {
    const int r0 = T.row0[s0 + d], c0 = s_nc0[d];
    for (int t = tid; t < s_nm[d]; t += nth) {
      cxs[c0 + t] = T.cx[r0 + t];
      cys[c0 + t] = T.cy[r0 + t];
      cq[c0 + t] = T.mvq[r0 + t];
      colnode[c0 + t] = d;
    }
  }
  for (int c = NC + tid; c < LD; c += nth) { cxs[c] = 0; cys[c] = 0; cq[c] = 0; colnode[c] = 0; }
  __syncthreads();

  // ---- phase 2: covariance panel K_{pa,u} (Covariancef_inplace, :885) and K_uu (:892 / :934)
  for (int idx = tid; idx < P * LD; idx += nth) {
    const int i = idx / LD, c = idx - i * LD;
    panel[idx] = (c < NC) ? cov_eval(ct, pxs[i], pys[i], pq[i], cxs[c], cys[c], cq[c]) : 0.0;
  }
  if (MODE == 0) {
    for (int d = 0; d < nn; d++) {
      const int md = s_nm[d], c0 = s_nc0[d];
      double* Kuu = S2 + s_nsq0[d];
      for (int e = tid; e < md * md; e += nth) {
        const int r = e / md, r2 = e - r * md;
        Kuu[e] = cov_eval(ct, cxs[c0 + r], cys[c0 + r], cq[c0 + r], cxs[c0 + r2], cys[c0 + r2], cq[c0 + r2]);
      }
    }
  } else {
    for (int c = tid; c < NC; c += nth) S2[c] = cov_eval(ct, cxs[c], cys[c], cq[c], cxs[c], cys[c], cq[c]);
  }
  __syncthreads();

  const int n_cg = (NC + TC - 1) / TC;
  // ---- phase 3: Z = L^-1 K, ancestor tile by ancestor tile, deepest first so that it can be done in place.
  //      Block row j of L^-1 is [-G_a | Ri_a] with a = chain[j] (tree_utils.cpp:204-206).
  for (int j = k - 1; j >= 0; j--) {
    const int a = s_chain[j], mj = s_cm[j], poj = s_cpoff[j];
    const double* Ga = S.G + T.goff[a];
    const double* Ria = S.Ri + T.rioff[a];
    const int acoff = T.chain_off[a];
    const int n_rg = (mj + TR - 1) / TR;
    for (int item = tid; item < n_rg * n_cg; item += nth) {
      const int rg = item % n_rg, cg = item / n_rg;
      const int r0 = rg * TR, c0 = cg * TC;
      double acc[TR][TC];
#pragma unroll
      for (int tr = 0; tr < TR; tr++)
#pragma unroll
        for (int tc = 0; tc < TC; tc++) acc[tr][tc] = 0.0;
      const double* Ap[TR];
      for (int i2 = 0; i2 < j; i2++) {
        const int mi = s_cm[i2];
        const double* Gt = Ga + T.chain_boff[acoff + i2];
#pragma unroll
        for (int tr = 0; tr < TR; tr++) Ap[tr] = Gt + (size_t)min(r0 + tr, mj - 1) * mi;
        tile_mac<TR, TC>(acc, Ap, 1, panel + (size_t)s_cpoff[i2] * LD + c0, LD, mi, -1.0);
      }
#pragma unroll
      for (int tr = 0; tr < TR; tr++) Ap[tr] = Ria + (size_t)min(r0 + tr, mj - 1) * mj;
      tile_mac<TR, TC>(acc, Ap, 1, panel + (size_t)poj * LD + c0, LD, mj, 1.0);
#pragma unroll
      for (int tr = 0; tr < TR; tr++)
        if (r0 + tr < mj)
#pragma unroll
          for (int tc = 0; tc < TC; tc++) S1[(r0 + tr) * LD + c0 + tc] = acc[tr][tc];
    }
    __syncthreads();
    for (int e = tid; e < mj * LD; e += nth) panel[(size_t)poj * LD + e] = S1[e];
    __syncthreads();
  }

  // ---- phase 4: Schur complement R = K_uu - Z'Z (equals Kcc - H*Kxc, :896-897)
  if (MODE == 0) {
    int item0 = 0;
    for (int d = 0; d < nn; d++) {
      const int md = s_nm[d], c0d = s_nc0[d];
      double* R = S2 + s_nsq0[d];
      const int nrg = (md + TR - 1) / TR, ncg = (md + TC - 1) / TC;
      for (int item = tid - item0; item < nrg * ncg; item += nth) {
        if (item < 0) continue;
        const int rg = item % nrg, cg = item / nrg;
        const int r0 = rg * TR, c0 = cg * TC;
        double acc[TR][TC];
#pragma unroll
        for (int tr = 0; tr < TR; tr++)
#pragma unroll
          for (int tc = 0; tc < TC; tc++) acc[tr][tc] = 0.0;
        for (int pp = 0; pp < P; pp++) {
          double a[TR], b[TC];
#pragma unroll
          for (int tr = 0; tr < TR; tr++) a[tr] = panel[(size_t)pp * LD + c0d + min(r0 + tr, md - 1)];
#pragma unroll
          for (int tc = 0; tc < TC; tc++) b[tc] = panel[(size_t)pp * LD + c0d + min(c0 + tc, md - 1)];
#pragma unroll
          for (int tr = 0; tr < TR; tr++)
#pragma unroll
            for (int tc = 0; tc < TC; tc++) acc[tr][tc] = fma(a[tr], b[tc], acc[tr][tc]);
        }
#pragma unroll
        for (int tr = 0; tr < TR; tr++)
#pragma unroll
          for (int tc = 0; tc < TC; tc++)
            if (r0 + tr < md && c0 + tc < md) R[(r0 + tr) * md + c0 + tc] -= acc[tr][tc];
      }
      item0 = (item0 + nrg * ncg) % nth;
    }
    __syncthreads();
    // ---- phase 5: Ri = chol(R)^-1 (:896), one warp per block
    for (int d = warp; d < nn; d += nwarps) {
      const int md = s_nm[d];
      double* R = S2 + s_nsq0[d];
      double* X = S1 + s_nsq0[d];
      if (warp_chol(R, md, lane)) {
        warp_inv_lower(R, X, md, lane);
      } else {
        if (lane == 0) atomicAdd(fail, 1);
        for (int e = lane; e < md * md; e += 32) X[e] = 0.0;
      }
    }
    __syncthreads();
    for (int e = tid; e < sumsq; e += nth) S2[e] = S1[e];
    __syncthreads();
  } else {
    for (int c = tid; c < NC; c += nth) {
      double s = 0;
      for (int pp = 0; pp < P; pp++) { const double z = panel[(size_t)pp * LD + c]; s = fma(z, z, s); }
      const double R = S2[c] - s;
      const bool ok = (R > 0.0) && isfinite(R);
      if (MODE == 1) {
        if (!ok) atomicAdd(fail, 1);
        S2[c] = ok ? 1.0 / sqrt(R) : 0.0;  // ccholprecdiag (:945)
      } else {
        S2[c] = ok ? sqrt(R) : 0.0;  // predict_std zeroes the sd on failure (:1316-1322)
      }
    }
    __syncthreads();
  }

  // ---- phase 6: H' = L^-T Z, shallowest tile first (in place again)
  for (int i = 0; i < k; i++) {
    const int ai = s_chain[i], mi = s_cm[i], poi = s_cpoff[i];
    const double* Rii = S.Ri + T.rioff[ai];
    const int n_rg = (mi + TR - 1) / TR;
    for (int item = tid; item < n_rg * n_cg; item += nth) {
      const int rg = item % n_rg, cg = item / n_rg;
      const int r0 = rg * TR, c0 = cg * TC;
      double acc[TR][TC];
#pragma unroll
      for (int tr = 0; tr < TR; tr++)
#pragma unroll
        for (int tc = 0; tc < TC; tc++) acc[tr][tc] = 0.0;
      const double* Ap[TR];
#pragma unroll
      for (int tr = 0; tr < TR; tr++) Ap[tr] = Rii + min(r0 + tr, mi - 1);
      tile_mac<TR, TC>(acc, Ap, mi, panel + (size_t)poi * LD + c0, LD, mi, 1.0);
      for (int j = i + 1; j < k; j++) {
        const int aj = s_chain[j], mj = s_cm[j];
        const double* Gt = S.G + T.goff[aj] + T.chain_boff[T.chain_off[aj] + i];
#pragma unroll
        for (int tr = 0; tr < TR; tr++) Ap[tr] = Gt + min(r0 + tr, mi - 1);
        tile_mac<TR, TC>(acc, Ap, mi, panel + (size_t)s_cpoff[j] * LD + c0, LD, mj, -1.0);
      }
#pragma unroll
      for (int tr = 0; tr < TR; tr++)
        if (r0 + tr < mi)
#pragma unroll
          for (int tc = 0; tc < TC; tc++) S1[(r0 + tr) * LD + c0 + tc] = acc[tr][tc];
    }
    __syncthreads();
    for (int e = tid; e < mi * LD; e += nth) panel[(size_t)poi * LD + e] = S1[e];
    __syncthreads();
  }

  // ---- phase 7: outputs.  panel[p][c] = H(c, p).
  // e = w_x - H w_pa (:888)
  for (int c = tid; c < NC; c += nth) {
    const int d = colnode[c];
    double s = w[T.row0[s0 + d] + (c - s_nc0[d])];
    for (int pp = 0; pp < P; pp++) s = fma(-panel[(size_t)pp * LD + c], wpa[pp], s);
    ecol[c] = s;
  }
  __syncthreads();
  for (int d = 0; d < nn; d++) {
    const int sd = s0 + d, md = s_nm[d], c0d = s_nc0[d];
    const long long go = T.goff[sd];
    const double* Rid = S2 + s_nsq0[d];  // MODE 0: Ri (m x m); MODE 1: 1/sqrt(R_ii); MODE 2: sqrt(R_ii)
    // G = Ri H (bottom-left block of Kxx_invchol(u), tree_utils.cpp:205) and H, stored as one tile per ancestor
    for (int j = 0; j < k; j++) {
      const int mj = s_cm[j], poj = s_cpoff[j];
      const long long bo = go + T.chain_boff[T.chain_off[sd] + j];
      for (int e = tid; e < md * mj; e += nth) {
        const int r = e / mj, pp = e - r * mj;
        const double* hp = panel + (size_t)(poj + pp) * LD + c0d;
        if (MODE == 2) {
          outH[bo + e] = hp[r];
        } else {
          if (keep_H) outH[bo + e] = hp[r];
          double g;
          if (MODE == 0) {
            g = 0;
            for (int r2 = 0; r2 <= r; r2++) g = fma(Rid[r * md + r2], hp[r2], g);
          } else {
            g = Rid[r] * hp[r];
          }
          S.G[bo + e] = g;
        }
      }
    }
    const long long ro = T.rioff[sd];
    for (int e = tid; e < ((MODE == 0) ? md * md : md); e += nth) outRi[ro + e] = Rid[e];
    if (MODE != 2 && warp == (d % nwarps)) {
      // wcore = e' prec e = |Ri e|^2 (:913 / :950) ; logdet = sum log diag(Ri) (:966)
      double wc = 0, ld = 0;
      for (int r = lane; r < md; r += 32) {
        double t;
        if (MODE == 0) {
          t = 0;
          for (int r2 = 0; r2 <= r; r2++) t = fma(Rid[r * md + r2], ecol[c0d + r2], t);
          ld += log(Rid[r * md + r]);
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
}

template <int MODE>
static cudaError_t launch_build_t(const DevTree& T, const DevSlot& S, double* outH, double* outRi, const int* grp_slot0,
                                  const int* grp_nn, int ngrp, const double* w, const CovTab& tab, int* fail, int keep_H,
                                  size_t smem, cudaStream_t st) {
  auto kern = build_level_kernel<kBuildTR, kBuildTC, MODE>;
  static size_t configured[3] = {0, 0, 0};
  if (smem > configured[MODE]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured[MODE] = smem;
  }
  kern<<<ngrp, kBuildThreads, smem, st>>>(T, S, outH, outRi, grp_slot0, grp_nn, w, tab, fail, keep_H);
  return cudaGetLastError();
}
cudaError_t launch_build(int mode, const DevTree& T, const DevSlot& S, double* outH, double* outRi, const int* grp_slot0,
                         const int* grp_nn, int ngrp, const double* w, const CovTab& tab, int* fail, int keep_H,
                         size_t smem, cudaStream_t st) {
  if (ngrp <= 0) return cudaSuccess;
  if (mode == 0) return launch_build_t<0>(T, S, outH, outRi, grp_slot0, grp_nn, ngrp, w, tab, fail, keep_H, smem, st);
  if (mode == 1) return launch_build_t<1>(T, S, outH, outRi, grp_slot0, grp_nn, ngrp, w, tab, fail, keep_H, smem, st);
  return launch_build_t<2>(T, S, outH, outRi, grp_slot0, grp_nn, ngrp, w, tab, fail, keep_H, smem, st);
}

// ------------------------------------------------------------------------------------------------ GIBBS
// One block per tree block (node).  Shared memory: RiS m*m | Sig m*m | wpa P | gwj (k+1)*m | smu m | wn m | rr m
// REF = 1: full m x m conditional (:1037-1089); REF = 0: row-wise scalars (:1091-1155)
template <int REF>
__global__ void __launch_bounds__(kGibbsThreads)
gibbs_level_kernel(DevTree T, DevSlot S, int slot0, double* __restrict__ w, const double* __restrict__ xb,
                   const double* __restrict__ z, const double* __restrict__ tausq_inv, const double* __restrict__ SigS,
                   double* __restrict__ V, double* probe_sig, double* probe_smu, int* __restrict__ fail) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, nth = blockDim.x, lane = tid & 31, warp = tid >> 5;
  const int sd = slot0 + blockIdx.x;
  const int m = T.m[sd], k = T.k[sd], P = T.P[sd], coff = T.chain_off[sd], row0 = T.row0[sd];
  const int msq = REF ? m * m : m;
  double* RiS = reinterpret_cast<double*>(smem_raw);
  double* Sig = RiS + msq;
  double* wpa = Sig + msq;
  double* gwj = wpa + P;  // k tiles of m, then the total at [k*m, (k+1)*m)
  double* smu = gwj + (size_t)(k + 1) * m;
  double* wn = smu + m;
  double* rr = wn + m;
  const double* G = S.G + T.goff[sd];
  const double* Rig = S.Ri + T.rioff[sd];
  const int nch = T.child_ptr[sd + 1] - T.child_ptr[sd];
  const int* ch = T.child_idx + T.child_ptr[sd];

  for (int e = tid; e < msq; e += nth) RiS[e] = Rig[e];
  for (int j = 0; j < k; j++) {
    const int a = T.chain[coff + j], po = T.chain_poff[coff + j], r0 = T.row0[a], ma = T.m[a];
    for (int t = tid; t < ma; t += nth) wpa[po + t] = w[r0 + t];
  }
  __syncthreads();
  // gwj[j][r] = G_j(r,:) w_{a_j}   (pieces of H w_pa scaled by Ri; :1063 / :1103)
  for (int e = tid; e < k * m; e += nth) {
    const int j = e / m, r = e - j * m;
    const int mj = T.m[T.chain[coff + j]], po = T.chain_poff[coff + j];
    const double* g = G + T.chain_boff[coff + j] + (size_t)r * mj;
    double s = 0;
    for (int pp = 0; pp < mj; pp++) s = fma(g[pp], wpa[po + pp], s);
    gwj[e] = s;
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
    const long long so = T.soff[sd];
    for (int e = tid; e < m * m; e += nth) {
      const int a = e / m, b = e - a * m;
      double s = 0;
      for (int r = max(a, b); r < m; r++) s = fma(RiS[r * m + a], RiS[r * m + b], s);
      if (so >= 0) s += SigS[so + e];
      if (a == b) s += tausq_inv[T.mvq[row0 + a]];
      Sig[e] = s;
      if (probe_sig) probe_sig[T.rioff[sd] + e] = s;
    }
    // Smu_tot = (H'prec)' w_pa + sum_children + tausq_inv (y - XB)  (:1062-1077)
    for (int a = tid; a < m; a += nth) {
      double s = 0;
      for (int r = a; r < m; r++) s = fma(RiS[r * m + a], gw[r], s);
      for (int c = 0; c < nch; c++) s += V[T.voff[ch[c]] + P + a];
      s += tausq_inv[T.mvq[row0 + a]] * (T.y[row0 + a] - xb[row0 + a]);
      smu[a] = s;
      if (probe_smu) probe_smu[row0 + a] = s;
    }
    __syncthreads();
    if (warp == 0) {
      // w = Sc'(Sc Smu + z), Sc = chol(Sigi_tot)^-1 (:1054, :1086), done as two triangular solves
      if (warp_chol(Sig, m, lane)) {
        for (int c = 0; c < m; c++) {  // forward: L x = smu
          __syncwarp();
          const double xc = smu[c] / Sig[c * m + c];
          __syncwarp();
          if (lane == 0) smu[c] = xc;
          for (int r = c + 1 + lane; r < m; r += 32) smu[r] -= Sig[r * m + c] * xc;
        }
        __syncwarp();
        for (int r = lane; r < m; r += 32) smu[r] += z[row0 + r];
        for (int c = m - 1; c >= 0; c--) {  // backward: L' w = x
          __syncwarp();
          const double xc = smu[c] / Sig[c * m + c];
          __syncwarp();
          if (lane == 0) smu[c] = xc;
          for (int r = lane; r < c; r += 32) smu[r] -= Sig[c * m + r] * xc;
        }
        __syncwarp();
        for (int r = lane; r < m; r += 32) { wn[r] = smu[r]; w[row0 + r] = smu[r]; }
      } else {
        if (lane == 0) atomicAdd(fail, 1);
        for (int r = lane; r < m; r += 32) wn[r] = w[row0 + r];
      }
    }
    __syncthreads();
    for (int r = tid; r < m; r += nth) {
      double s = -gw[r];
      for (int r2 = 0; r2 <= r; r2++) s = fma(RiS[r * m + r2], wn[r2], s);
      rr[r] = s;
    }
  } else {
    for (int a = tid; a < m; a += nth) {
      const double ri = RiS[a], prec = ri * ri;
      const double tsqi = tausq_inv[T.mvq[row0 + a]];
      const double sig = prec + tsqi;                                             // :1123
      const double sm = ri * gw[a] + tsqi * (T.y[row0 + a] - xb[row0 + a]);       // :1125-1127
      if (probe_sig) probe_sig[T.rioff[sd] + a] = sig;
      if (probe_smu) probe_smu[row0 + a] = sm;
      double wa;
      if (sig > 0.0 && isfinite(sig)) {
        const double sc = 1.0 / sqrt(sig);
        wa = sc * sc * sm + sc * z[row0 + a];                                     // :1139-1140
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
  for (int j = 0; j < k; j++) {
    const int mj = T.m[T.chain[coff + j]], po = T.chain_poff[coff + j];
    const double* g = G + T.chain_boff[coff + j];
    const double* vj = gwj + (size_t)j * m;
    for (int pp = tid; pp < mj; pp += nth) {
      double s = 0;
      for (int r = 0; r < m; r++) s = fma(g[(size_t)r * mj + pp], vj[r], s);
      for (int c = 0; c < nch; c++) s += V[T.voff[ch[c]] + po + pp];
      Vd[po + pp] = s;
    }
  }
}

cudaError_t launch_gibbs(int is_ref, const DevTree& T, const DevSlot& S, int slot0, int nslots, double* w,
                         const double* xb, const double* z, const double* tausq_inv, const double* SigS, double* V,
                         double* probe_sig, double* probe_smu, int* fail, size_t smem, cudaStream_t st) {
  if (nslots <= 0) return cudaSuccess;
  static size_t configured[2] = {0, 0};
  if (is_ref) {
    auto kern = gibbs_level_kernel<1>;
    if (smem > configured[1]) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
      configured[1] = smem;
    }
    kern<<<nslots, kGibbsThreads, smem, st>>>(T, S, slot0, w, xb, z, tausq_inv, SigS, V, probe_sig, probe_smu, fail);
  } else {
    auto kern = gibbs_level_kernel<0>;
    if (smem > configured[0]) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
      configured[0] = smem;
    }
    kern<<<nslots, kGibbsThreads, smem, st>>>(T, S, slot0, w, xb, z, tausq_inv, SigS, V, probe_sig, probe_smu, fail);
  }
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ message Grams
// U_d[j] = G_j'G_j + sum_children U_c[j]  (= sum over the subtree of AK_uP_u_all[J_j, J_j], :1190-1192) and
// SigS_d = sum_children U_c[tile of d]  (= arma::sum(Sigi_children(d), 2), :1049).  One block per node.
__global__ void __launch_bounds__(kGramThreads)
gram_level_kernel(DevTree T, DevSlot S, int slot0, double* __restrict__ U, double* __restrict__ SigS) {
  const int tid = threadIdx.x, nth = blockDim.x;
  const int sd = slot0 + blockIdx.x;
  const int m = T.m[sd], k = T.k[sd], coff = T.chain_off[sd];
  const int nch = T.child_ptr[sd + 1] - T.child_ptr[sd];
  const int* ch = T.child_idx + T.child_ptr[sd];
  const double* G = S.G + T.goff[sd];
  double* Ud = U + T.uoff[sd];
  for (int j = 0; j < k; j++) {
    const int mj = T.m[T.chain[coff + j]];
    const double* g = G + T.chain_boff[coff + j];
    const int uo = T.chain_uoff[coff + j];
    for (int e = tid; e < mj * mj; e += nth) {
      const int a = e / mj, b = e - a * mj;
      double s = 0;
      for (int r = 0; r < m; r++) s = fma(__ldg(g + (size_t)r * mj + a), __ldg(g + (size_t)r * mj + b), s);
      for (int c = 0; c < nch; c++) s += U[T.uoff[ch[c]] + T.chain_uoff[T.chain_off[ch[c]] + j] + e];
      Ud[uo + e] = s;
    }
  }
  const long long so = T.soff[sd];
  if (so >= 0) {
    for (int e = tid; e < m * m; e += nth) {
      double s = 0;
      for (int c = 0; c < nch; c++) s += U[T.uoff[ch[c]] + T.chain_uoff[T.chain_off[ch[c]] + k] + e];
      SigS[so + e] = s;
    }
  }
}
cudaError_t launch_gram(const DevTree& T, const DevSlot& S, int slot0, int nslots, double* U, double* SigS,
                        cudaStream_t st) {
  if (nslots <= 0) return cudaSuccess;
  gram_level_kernel<<<nslots, kGramThreads, 0, st>>>(T, S, slot0, U, SigS);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ LLW
// llcomp[s] = m*hl2pi - 0.5*|Ri w_u - G w_pa|^2  (get_loglik_w_std, :791-813); one warp per node
__global__ void __launch_bounds__(kLlwThreads)
llw_kernel(DevTree T, DevSlot S, int nslots, const double* __restrict__ w) {
  const int lane = threadIdx.x & 31;
  const int sd = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (sd >= nslots) return;
  const int m = T.m[sd], k = T.k[sd], coff = T.chain_off[sd], row0 = T.row0[sd];
  const bool ref = T.isref[sd] != 0;
  const double* G = S.G + T.goff[sd];
  const double* Ri = S.Ri + T.rioff[sd];
  double wc = 0;
  for (int r = 0; r < m; r++) {
    double s = 0;
    if (ref) {
      for (int r2 = lane; r2 <= r; r2 += 32) s = fma(Ri[r * m + r2], w[row0 + r2], s);
    } else if (lane == 0) {
      s = Ri[r] * w[row0 + r];
    }
    for (int j = 0; j < k; j++) {
      const int a = T.chain[coff + j], mj = T.m[a], ar0 = T.row0[a];
      const double* g = G + T.chain_boff[coff + j] + (size_t)r * mj;
      for (int pp = lane; pp < mj; pp += 32) s = fma(-g[pp], w[ar0 + pp], s);
    }
    s = warp_sum(s);
    wc = fma(s, s, wc);
  }
  if (lane == 0) S.llcomp[sd] = (double)m * kHl2pi - 0.5 * wc;
}
cudaError_t launch_llw(const DevTree& T, const DevSlot& S, int nslots, const double* w, cudaStream_t st) {
  if (nslots <= 0) return cudaSuccess;
  const int wpb = kLlwThreads / 32;
  llw_kernel<<<(nslots + wpb - 1) / wpb, kLlwThreads, 0, st>>>(T, S, nslots, w);
  return cudaGetLastError();
}

// out[0] = sum(logdet) + sum(llcomp), out[1] = sum(logdet), out[2] = (fail == 0)  (:987-988 / :815-816).
// Fixed summation order: deterministic run to run.
__global__ void __launch_bounds__(1024) loglik_reduce_kernel(const double* __restrict__ logdet,
                                                             const double* __restrict__ llcomp, int n,
                                                             const int* __restrict__ fail, double* __restrict__ out) {
  __shared__ double sa[1024], sb[1024];
  double a = 0, b = 0;
  for (int i = threadIdx.x; i < n; i += 1024) { a += logdet[i]; b += llcomp[i]; }
  sa[threadIdx.x] = a;
  sb[threadIdx.x] = b;
  __syncthreads();
  for (int s = 512; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) { sa[threadIdx.x] += sa[threadIdx.x + s]; sb[threadIdx.x] += sb[threadIdx.x + s]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    out[0] = sa[0] + sb[0];
    out[1] = sa[0];
    out[2] = (fail == nullptr || *fail == 0) ? 1.0 : 0.0;
  }
}
cudaError_t launch_loglik_reduce(const double* logdet, const double* llcomp, int n, const int* fail, double* out,
                                 cudaStream_t st) {
  loglik_reduce_kernel<<<1, 1024, 0, st>>>(logdet, llcomp, n, fail, out);
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
  const double* sdv = sdpred + T.rioff[sd];
  for (int r = threadIdx.x; r < m; r += blockDim.x) {
    double s = 0;
    for (int j = 0; j < k; j++) {
      const int a = T.chain[coff + j], mj = T.m[a], ar0 = T.row0[a];
      const double* h = H + T.chain_boff[coff + j] + (size_t)r * mj;
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
__global__ void normals_kernel(double* __restrict__ z, long long n, uint64_t seed, uint64_t counter) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long pair = i;  // each thread produces z[2i], z[2i+1]
  if (2 * pair >= n) return;
  uint32_t c[4] = {(uint32_t)pair, (uint32_t)(pair >> 32), (uint32_t)counter, (uint32_t)(counter >> 32)};
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; r++) { philox_round(c, k0, k1); k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
  const double u1 = ((((uint64_t)c[0] << 32 | c[1]) >> 11) + 0.5) * (1.0 / 9007199254740992.0);
  const double u2 = ((((uint64_t)c[2] << 32 | c[3]) >> 11) + 0.5) * (1.0 / 9007199254740992.0);
  const double rad = sqrt(-2.0 * log(u1));
  double sn, cs;
  sincospi(2.0 * u2, &sn, &cs);
  z[2 * pair] = rad * cs;
  if (2 * pair + 1 < n) z[2 * pair + 1] = rad * sn;
}
cudaError_t launch_normals(double* z, long long n, uint64_t seed, uint64_t counter, cudaStream_t st) {
  const long long pairs = (n + 1) / 2;
  if (pairs <= 0) return cudaSuccess;
  normals_kernel<<<(unsigned)((pairs + 255) / 256), 256, 0, st>>>(z, n, seed, counter);
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
__global__ void rowstats_final_kernel(const double* __restrict__ partial, int nblocks, int nv, double* __restrict__ out) {
  const int v = threadIdx.x;
  if (v >= nv) return;
  double s = 0;
  for (int b = 0; b < nblocks; b++) s += partial[(size_t)b * nv + v];
  out[v] = s;
}
cudaError_t launch_rowstats(const DevTree& T, const int* widx, long long n_all, int p, int q, const double* w,
                            const double* xb, double* partial, int nblocks, double* out, cudaStream_t st) {
  rowstats_kernel<<<nblocks, 256, 0, st>>>(T, widx, n_all, p, q, w, xb, partial);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  rowstats_final_kernel<<<1, 256, 0, st>>>(partial, nblocks, q * (p + 1), out);
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

}  // namespace st
