// Microbenchmark of the warp-level 25 x 25 Cholesky factorisation used by BUILD / GIBBS: cycles per pivot step for
// several formulations (one warp per CTA, many CTAs, clock64 around the factorisation only).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/microbench/chol_bench.cu -o tools/microbench/chol_bench
#include <cstdio>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ double rsqrt_fast(double d) {  // d > 0, finite, normal
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  // two Newton steps in the residual form: e = 1 - d y^2 ; y += y e (1/2 + 3/8 e)
  double e = fma(-d * y, y, 1.0);
  y = fma(y * e, fma(e, 0.375, 0.5), y);
  e = fma(-d * y, y, 1.0);
  y = fma(y * e, fma(e, 0.375, 0.5), y);
  return y;
}

// V0: as in st_build.cu (shared-memory broadcast of the pivot column, library rsqrt)
template <int VAR>
__device__ __forceinline__ bool chol_var(double* R, int m, int rs, double* cb, double* dv, int lane) {
  bool ok = true;
  double a[32];
#pragma unroll
  for (int j = 0; j < 32; j++) a[j] = (lane < m && j <= lane) ? R[lane * rs + j] : 0.0;
  cb[32 + lane] = 0.0;
  cb[96 + lane] = 0.0;
  if (VAR == 0 || VAR == 1 || VAR == 3) {
    for (int j = 0; j < m; j++) {
      double d = __shfl_sync(0xffffffffu, a[0], j);
      if (!(d > 0.0) || !isfinite(d)) { ok = false; d = 1.0; }
      const double inv = (VAR == 3) ? rsqrt_fast(d) : rsqrt(d), sd = d * inv;
      const double l = (lane == j) ? sd : ((lane > j) ? a[0] * inv : 0.0);
      double* c = cb + (j & 1) * 64;
      c[lane] = l;
      if (lane >= j && lane < m) R[lane * rs + j] = l;
      if (lane == 0) dv[j] = inv;
      __syncwarp();
      if (VAR == 1) {  // pivot chain only: rotate without the update
#pragma unroll
        for (int i = 0; i < 31; i++) a[i] = a[i + 1];
      } else {
        const double* cj = c + j + 1;
#pragma unroll
        for (int i = 0; i < 31; i++) a[i] = fma(-l, cj[i], a[i + 1]);
      }
      a[31] = 0.0;
    }
  } else if (VAR == 6) {  // smem broadcast, trailing update in uniform groups of 8 columns that still exist
    for (int j = 0; j < m; j++) {
      double d = __shfl_sync(0xffffffffu, a[0], j);
      if (!(d > 0.0) || !isfinite(d)) { ok = false; d = 1.0; }
      const double inv = rsqrt(d), sd = d * inv;
      const double l = (lane == j) ? sd : ((lane > j) ? a[0] * inv : 0.0);
      double* c = cb + (j & 1) * 64;
      c[lane] = l;
      if (lane >= j && lane < m) R[lane * rs + j] = l;
      if (lane == 0) dv[j] = inv;
      __syncwarp();
      const double* cj = c + j + 1;
      const int rem = m - 1 - j;
#pragma unroll
      for (int g = 0; g < 4; g++) {
        if (8 * g < rem) {
#pragma unroll
          for (int i = 8 * g; i < 8 * g + 8 && i < 31; i++) a[i] = fma(-l, cj[i], a[i + 1]);
        } else {
#pragma unroll
          for (int i = 8 * g; i < 8 * g + 8 && i < 31; i++) a[i] = a[i + 1];
        }
      }
      a[31] = 0.0;
    }
  } else if (VAR == 2) {  // shuffle broadcast
    for (int j = 0; j < m; j++) {
      double d = __shfl_sync(0xffffffffu, a[0], j);
      if (!(d > 0.0) || !isfinite(d)) { ok = false; d = 1.0; }
      const double inv = rsqrt(d), sd = d * inv;
      const double l = (lane == j) ? sd : ((lane > j) ? a[0] * inv : 0.0);
      if (lane >= j && lane < m) R[lane * rs + j] = l;
      if (lane == 0) dv[j] = inv;
#pragma unroll
      for (int i = 0; i < 31; i++) {
        const double lc = __shfl_sync(0xffffffffu, l, min(j + 1 + i, 31));
        a[i] = fma(-l, lc, a[i + 1]);
      }
      a[31] = 0.0;
    }
  } else if (VAR == 4 || VAR == 5) {
    // look-ahead: the next pivot's rsqrt chain is started before the trailing update of the current step
    double d = __shfl_sync(0xffffffffu, a[0], 0);
    if (!(d > 0.0) || !isfinite(d)) { ok = false; d = 1.0; }
    double inv = rsqrt_fast(d);
    for (int j = 0; j < m; j++) {
      const double sd = d * inv;
      const double l = (lane == j) ? sd : ((lane > j) ? a[0] * inv : 0.0);
      // next pivot: row j+1's diagonal after this step's update, known to lane j+1 from its own l
      double dn = __shfl_sync(0xffffffffu, fma(-l, l, a[1]), min(j + 1, 31));
      if (j + 1 < m && (!(dn > 0.0) || !isfinite(dn))) { ok = false; dn = 1.0; }
      const double invn = rsqrt_fast(dn);
      if (lane >= j && lane < m) R[lane * rs + j] = l;
      if (lane == 0) dv[j] = inv;
      if (VAR == 4) {
        double* c = cb + (j & 1) * 64;
        c[lane] = l;
        __syncwarp();
        const double* cj = c + j + 1;
#pragma unroll
        for (int i = 0; i < 31; i++) a[i] = fma(-l, cj[i], a[i + 1]);
      } else {
#pragma unroll
        for (int i = 0; i < 31; i++) {
          const double lc = __shfl_sync(0xffffffffu, l, min(j + 1 + i, 31));
          a[i] = fma(-l, lc, a[i + 1]);
        }
      }
      a[31] = 0.0;
      d = dn;
      inv = invn;
    }
  }
  return ok;
}

// inverse of the lower-triangular L (in R, stride rs, rows padded to 32 with zeros) in place; dv = 1 / diag
// I0: as in st_build.cu (row r of L^-1 by forward substitution, lane = column)
// I1: right-looking with rotating accumulators: lane c keeps the not-yet-final entries of column c of L^-1 in registers,
//     acc[0] is row k; once x_k = acc[0] / L_kk is final, rows k+1.. receive -L[r][k] x_k (column k of L by broadcast loads)
template <int IV>
__device__ __forceinline__ void inv_var(double* R, int m, int rs, const double* dv, int lane) {
  if (IV == 0) {
    for (int r = 0; r < m; r++) {
      const double* Lr = R + r * rs;
      double s0 = (r == lane) ? 1.0 : 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
      int kx = 0;
      for (; kx + 3 < r; kx += 4) {
        s0 = fma(-Lr[kx], R[kx * rs + lane], s0);
        s1 = fma(-Lr[kx + 1], R[(kx + 1) * rs + lane], s1);
        s2 = fma(-Lr[kx + 2], R[(kx + 2) * rs + lane], s2);
        s3 = fma(-Lr[kx + 3], R[(kx + 3) * rs + lane], s3);
      }
      for (; kx < r; kx++) s0 = fma(-Lr[kx], R[kx * rs + lane], s0);
      const double xr = ((s0 + s1) + (s2 + s3)) * dv[r];
      __syncwarp();
      if (lane < m) R[r * rs + lane] = (lane <= r) ? xr : 0.0;
      __syncwarp();
    }
  } else {
    double acc[32];
#pragma unroll
    for (int i = 0; i < 32; i++) acc[i] = (i == lane) ? 1.0 : 0.0;
    const int rmax = ((m + 7) & ~7) - 1;  // last row that exists in R
    for (int k = 0; k < m; k++) {
      const double x = acc[0] * dv[k];
      const double* col = R + k;  // column k of L: L[r][k] at col[r * rs]
#pragma unroll
      for (int i = 0; i < 31; i++) acc[i] = fma(-col[min(k + 1 + i, rmax) * rs], x, acc[i + 1]);
      acc[31] = 0.0;
      __syncwarp();  // every lane has read column k of L (row k of L is dead: its later columns are never read again)
      if (lane < m) R[k * rs + lane] = x;  // zero above the diagonal (lanes > k never received a contribution)
    }
    __syncwarp();
  }
}

template <int IV>
__global__ void bench_inv(const double* A, double* out, long long* cyc, int m, int rs, int reps) {
  __shared__ double R[32 * 36], cb[128], dv[32];
  const int lane = threadIdx.x;
  long long tot = 0;
  for (int r = 0; r < reps; r++) {
    for (int e = lane; e < 32 * rs; e += 32) R[e] = 0.0;
    __syncwarp();
    for (int e = lane; e < m * m; e += 32) { const int i = e / m, j = e % m; R[i * rs + j] = A[e]; }
    __syncwarp();
    chol_var<0>(R, m, rs, cb, dv, lane);
    __syncwarp();
    const long long t0 = clock64();
    inv_var<IV>(R, m, rs, dv, lane);
    __syncwarp();
    tot += clock64() - t0;
  }
  if (lane == 0) cyc[blockIdx.x] = tot;
  if (blockIdx.x == 0) for (int e = lane; e < m * m; e += 32) { const int i = e / m, j = e % m; out[e] = (j <= i) ? R[i * rs + j] : 0.0; }
}

template <int VAR>
__global__ void bench(const double* A, double* out, long long* cyc, int m, int rs, int reps) {
  __shared__ double R[32 * 36], cb[128], dv[32];
  const int lane = threadIdx.x;
  long long tot = 0;
  bool ok = true;
  for (int r = 0; r < reps; r++) {
    for (int e = lane; e < 32 * rs; e += 32) R[e] = 0.0;
    __syncwarp();
    for (int e = lane; e < m * m; e += 32) { const int i = e / m, j = e % m; R[i * rs + j] = A[e]; }
    __syncwarp();
    const long long t0 = clock64();
    ok &= chol_var<VAR>(R, m, rs, cb, dv, lane);
    __syncwarp();
    tot += clock64() - t0;
  }
  if (lane == 0) cyc[blockIdx.x] = tot;
  if (blockIdx.x == 0) for (int e = lane; e < m * m; e += 32) { const int i = e / m, j = e % m; out[e] = (j <= i && ok) ? R[i * rs + j] : 0.0; }
}

int main() {
  const int m = 25, rs = 28, reps = 20, nblk = 148 * 4;
  std::vector<double> A(m * m), L(m * m, 0.0);
  for (int i = 0; i < m; i++)
    for (int j = 0; j < m; j++) A[i * m + j] = exp(-0.3 * fabs(i - j)) + (i == j ? 0.5 : 0.0);
  // host reference
  std::vector<double> W = A;
  for (int j = 0; j < m; j++) {
    double d = W[j * m + j];
    for (int k = 0; k < j; k++) d -= L[j * m + k] * L[j * m + k];
    L[j * m + j] = sqrt(d);
    for (int i = j + 1; i < m; i++) {
      double s = W[i * m + j];
      for (int k = 0; k < j; k++) s -= L[i * m + k] * L[j * m + k];
      L[i * m + j] = s / L[j * m + j];
    }
  }
  double *dA, *dO;
  long long* dC;
  cudaMalloc(&dA, m * m * 8); cudaMalloc(&dO, m * m * 8); cudaMalloc(&dC, nblk * 8);
  cudaMemcpy(dA, A.data(), m * m * 8, cudaMemcpyHostToDevice);
  std::vector<double> O(m * m);
  std::vector<long long> C(nblk);
  auto run = [&](int var, const char* name) {
    for (int it = 0; it < 2; it++) {
      switch (var) {
        case 0: bench<0><<<nblk, 32>>>(dA, dO, dC, m, rs, reps); break;
        case 1: bench<1><<<nblk, 32>>>(dA, dO, dC, m, rs, reps); break;
        case 2: bench<2><<<nblk, 32>>>(dA, dO, dC, m, rs, reps); break;
        case 3: bench<3><<<nblk, 32>>>(dA, dO, dC, m, rs, reps); break;
        case 4: bench<4><<<nblk, 32>>>(dA, dO, dC, m, rs, reps); break;
        case 5: bench<5><<<nblk, 32>>>(dA, dO, dC, m, rs, reps); break;
        case 6: bench<6><<<nblk, 32>>>(dA, dO, dC, m, rs, reps); break;
      }
      cudaDeviceSynchronize();
    }
    cudaMemcpy(O.data(), dO, m * m * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(C.data(), dC, nblk * 8, cudaMemcpyDeviceToHost);
    double err = 0, cyc = 0;
    for (int e = 0; e < m * m; e++) err = fmax(err, fabs(O[e] - L[e]));
    for (auto c : C) cyc += (double)c;
    printf("%-46s %8.0f cycles / factorisation  %6.1f / pivot   max |L - L_ref| = %.2e  (%s)\n", name, cyc / nblk / reps, cyc / nblk / reps / m, err,
           cudaGetErrorString(cudaGetLastError()));
  };
  run(0, "V0 smem broadcast, library rsqrt (current)");
  run(1, "V1 pivot chain only (no trailing update)");
  run(2, "V2 shuffle broadcast, library rsqrt");
  run(3, "V3 smem broadcast, branch-free rsqrt");
  run(4, "V4 look-ahead pivot, smem broadcast");
  run(5, "V5 look-ahead pivot, shuffle broadcast");
  run(6, "V6 smem broadcast, remaining columns by 8");
  // inverse variants: compare with the host inverse of L
  std::vector<double> X(m * m, 0.0);
  for (int c = 0; c < m; c++)
    for (int r = c; r < m; r++) {
      double sacc = (r == c) ? 1.0 : 0.0;
      for (int k = c; k < r; k++) sacc -= L[r * m + k] * X[k * m + c];
      X[r * m + c] = sacc / L[r * m + r];
    }
  auto run_inv = [&](int iv, const char* name) {
    for (int it = 0; it < 2; it++) {
      if (iv == 0) bench_inv<0><<<nblk, 32>>>(dA, dO, dC, m, rs, reps); else bench_inv<1><<<nblk, 32>>>(dA, dO, dC, m, rs, reps);
      cudaDeviceSynchronize();
    }
    cudaMemcpy(O.data(), dO, m * m * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(C.data(), dC, nblk * 8, cudaMemcpyDeviceToHost);
    double err = 0, cyc = 0;
    for (int e = 0; e < m * m; e++) err = fmax(err, fabs(O[e] - X[e]));
    for (auto c : C) cyc += (double)c;
    printf("%-46s %8.0f cycles / inverse        %6.1f / row     max |X - X_ref| = %.2e  (%s)\n", name, cyc / nblk / reps, cyc / nblk / reps / m, err,
           cudaGetErrorString(cudaGetLastError()));
  };
  run_inv(0, "I0 row-wise forward substitution (current)");
  run_inv(1, "I1 right-looking, rotating accumulators");
  return 0;
}
