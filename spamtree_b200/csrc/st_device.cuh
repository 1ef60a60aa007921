// st_device.cuh — device helpers shared by the kernels (covariance evaluation, warp-level dense routines)
#pragma once
#include <cuda_runtime.h>

#include "st_kernels.cuh"

namespace st {

// the theta-slot `phys` of D, by selects (dynamic indexing of a kernel parameter would force a local-memory copy)
__device__ __forceinline__ DevSlot pick_slot(const DevSlots& D, int phys) {
  DevSlot S;
  S.G = phys ? D.s[1].G : D.s[0].G;
  S.H = phys ? D.s[1].H : D.s[0].H;
  S.Ri = phys ? D.s[1].Ri : D.s[0].Ri;
  S.logdet = phys ? D.s[1].logdet : D.s[0].logdet;
  S.llcomp = phys ? D.s[1].llcomp : D.s[0].llcomp;
  return S;
}

struct CovTabS {  // shared-memory copy of CovTab, plus the table of exp_neg
  int q;
  double c1[kMaxQ * kMaxQ], r1[kMaxQ * kMaxQ], c2[kMaxQ * kMaxQ], r2[kMaxQ * kMaxQ];
  double t64[64];  // 2^(-j/64)
};
__device__ __forceinline__ void load_covtab(CovTabS& s, const CovTab& t) {
  if (threadIdx.x == 0) s.q = t.q;
  for (int i = threadIdx.x; i < t.q * t.q; i += blockDim.x) {
    s.c1[i] = t.c1[i]; s.r1[i] = t.r1[i]; s.c2[i] = t.c2[i]; s.r2[i] = t.r2[i];
  }
  for (int i = threadIdx.x; i < 64; i += blockDim.x) s.t64[i] = exp2(-(double)i / 64.0);
}
// exp(-a) for a >= 0 (NaN propagates), about 1 ulp: a = (64 e + j) ln2/64 + r with |r| <= ln2/128, so that
// exp(-a) = 2^-e * 2^(-j/64) * exp(-r); table for the middle factor, degree-6 polynomial for the last, exponent
// arithmetic for the first.  11 FP64 operations and no slow path, against ~20 plus a branch in the general exp().
__device__ __forceinline__ double exp_neg(const double* __restrict__ t64, double a) {
  const double kd = fma(a, 92.332482616893656754, 6755399441055744.0);  // 64 / ln 2; magic constant: round to nearest
  const int ki = __double2loint(kd);
  const double kf = kd - 6755399441055744.0;
  double r = fma(kf, -0.01083042469326756, a);      // ln2/64, high part (21 trailing zero bits: kf * hi is exact)
  r = fma(kf, -2.9815858269852933e-12, r);         // low part
  const double s = -r;
  double p = fma(s, 1.3888888888888889e-03, 8.3333333333333332e-03);
  p = fma(p, s, 4.1666666666666664e-02);
  p = fma(p, s, 1.6666666666666666e-01);
  p = fma(p, s, 0.5);
  p = fma(p, s, 1.0);
  p = fma(p, s, 1.0);
  const double v = t64[ki & 63] * p;  // in (0.49, 1.01)
  const int e = ki >> 6;
  return (e > 1000) ? 0.0 : __hiloint2double(__double2hiint(v) - (e << 20), __double2loint(v));
}
// mvCovAG20107_inplace / cexpcov (covariance_functions.cpp:95-111, :213-286): see make_covtab() for (c1, r1, c2, r2)
__device__ __forceinline__ double cov_eval(const CovTabS& t, double x1, double y1, int q1, double x2, double y2, int q2) {
  const double dx = x1 - x2, dy = y1 - y2;
  const double h = sqrt(dx * dx + dy * dy);
  const int ix = q1 * t.q + q2;
  // branch-free (c2 = r2 = 0 for cross-outcome pairs): lets several evaluations per thread overlap their latencies
  return fma(t.c2[ix], exp_neg(t.t64, t.r2[ix] * h), t.c1[ix] * exp_neg(t.t64, t.r1[ix] * h));
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// in-place lower Cholesky of the symmetric m x m matrix A (row-major, leading dimension ld) by one warp;
// reads the lower triangle only; false on a non-positive or non-finite pivot (dpotrf's info > 0).
// Lane r owns row r; the trailing update is batched 8 columns at a time (loads, FMAs, stores) so that the
// shared-memory latencies overlap instead of chaining.
__device__ inline bool warp_chol(double* A, int m, int ld, int lane) {
  for (int c = 0; c < m; c++) {
    __syncwarp();
    double d = A[c * ld + c];
    if (!(d > 0.0) || !isfinite(d)) return false;
    d = sqrt(d);
    const double inv = 1.0 / d;
    __syncwarp();
    for (int r = c + lane; r < m; r += 32) A[r * ld + c] = (r == c) ? d : A[r * ld + c] * inv;
    __syncwarp();
    for (int r = c + 1 + lane; r < m; r += 32) {
      const double lrc = A[r * ld + c];
      double* row = A + r * ld;
      int c2 = c + 1;
      for (; c2 + 7 <= r; c2 += 8) {
        double x[8], l[8];
#pragma unroll
        for (int u = 0; u < 8; u++) { x[u] = row[c2 + u]; l[u] = A[(c2 + u) * ld + c]; }
#pragma unroll
        for (int u = 0; u < 8; u++) row[c2 + u] = fma(-lrc, l[u], x[u]);
      }
      for (; c2 <= r; c2++) row[c2] = fma(-lrc, A[c2 * ld + c], row[c2]);
    }
  }
  __syncwarp();
  return true;
}
// L <- L^-1 in place for a lower-triangular L (dtrtri); v: m doubles of scratch owned by the warp.
// The strict upper triangle is zero-filled afterwards so that the result can be used as a dense tile.
__device__ inline void warp_inv_lower_inplace(double* L, int m, int ld, double* v, int lane) {
  for (int j = m - 1; j >= 0; j--) {
    __syncwarp();
    const double ajj = 1.0 / L[j * ld + j];
    for (int r = j + 1 + lane; r < m; r += 32) v[r] = L[r * ld + j];
    __syncwarp();
    for (int r = j + 1 + lane; r < m; r += 32) {
      const double* row = L + r * ld;
      double s[4] = {0, 0, 0, 0};
      int kk = j + 1;
      for (; kk + 7 <= r; kk += 8) {
        double x[8], y[8];
#pragma unroll
        for (int u = 0; u < 8; u++) { x[u] = row[kk + u]; y[u] = v[kk + u]; }
#pragma unroll
        for (int u = 0; u < 8; u++) s[u & 3] = fma(x[u], y[u], s[u & 3]);
      }
      for (; kk <= r; kk++) s[0] = fma(row[kk], v[kk], s[0]);
      L[r * ld + j] = -((s[0] + s[1]) + (s[2] + s[3])) * ajj;
    }
    if (lane == 0) L[j * ld + j] = ajj;
  }
  __syncwarp();
  for (int e = lane; e < m * m; e += 32) {
    const int r = e / m, c = e - r * m;
    if (c > r) L[r * ld + c] = 0.0;
  }
  __syncwarp();
}

// X = L^-1 for a lower-triangular L, column c by lane c with no inter-lane dependency (forward substitution on e_c);
// X has the layout of L and must not alias it; dinv: m doubles of scratch (1 / diag(L)); strict upper part zeroed.
__device__ inline void warp_inv_lower_cols(const double* L, double* X, int m, int ld, double* dinv, int lane) {
  for (int r = lane; r < m; r += 32) dinv[r] = 1.0 / L[r * ld + r];
  __syncwarp();
  for (int c = lane; c < m; c += 32) {
    for (int r = 0; r < c; r++) X[r * ld + c] = 0.0;
    X[c * ld + c] = dinv[c];
    for (int r = c + 1; r < m; r++) {
      const double* row = L + r * ld;
      double s[4] = {0, 0, 0, 0};
      int kk = c;
      for (; kk + 7 < r; kk += 8) {
        double x[8], y[8];
#pragma unroll
        for (int u = 0; u < 8; u++) { x[u] = row[kk + u]; y[u] = X[(kk + u) * ld + c]; }
#pragma unroll
        for (int u = 0; u < 8; u++) s[u & 3] = fma(x[u], y[u], s[u & 3]);
      }
      for (; kk < r; kk++) s[0] = fma(row[kk], X[kk * ld + c], s[0]);
      X[r * ld + c] = -((s[0] + s[1]) + (s[2] + s[3])) * dinv[r];
    }
  }
  __syncwarp();
}

}  // namespace st
