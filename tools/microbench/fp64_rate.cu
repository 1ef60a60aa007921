// FP64 issue-rate microbenchmark for B200: DFMA (register-only), DFMA fed from shared memory like the BUILD inner loop,
// and DMMA m8n8k4 — per SM, as a function of resident warps.  Build: nvcc -arch=sm_100a -O3 -o fp64_rate fp64_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_reg(double* out, int iters) {
  double a[20];
  for (int i = 0; i < 20; i++) a[i] = threadIdx.x * 1e-3 + i;
  const double b = 1.0000001, c = 1e-9;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 20; i++) a[i] = fma(a[i], b, c);
  }
  double s = 0;
  for (int i = 0; i < 20; i++) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// 5x4 register tile, A from 5 rows of a shared tile (LDS.64), B as two LDS.128 — the BUILD sweep's inner loop
__global__ void dfma_smem(double* out, int iters) {
  __shared__ __align__(16) double As[32 * 26];
  __shared__ __align__(16) double Bs[32 * 106];
  for (int i = threadIdx.x; i < 32 * 26; i += blockDim.x) As[i] = 1e-3 * i;
  for (int i = threadIdx.x; i < 32 * 106; i += blockDim.x) Bs[i] = 1e-4 * i;
  __syncthreads();
  const int lane = threadIdx.x & 127;
  const int rg = lane % 5, cg = lane / 5;
  double acc[5][4] = {};
  for (int it = 0; it < iters; it++) {
    const double* pa = As + rg * 5 * 26;
    const double* pb = Bs + cg * 4;
#pragma unroll 4
    for (int kk = 0; kk < 24; kk++) {
      const double2 b01 = *reinterpret_cast<const double2*>(pb);
      const double2 b23 = *reinterpret_cast<const double2*>(pb + 2);
      pb += 106;
#pragma unroll
      for (int tr = 0; tr < 5; tr++) {
        const double a = pa[tr * 26 + kk];
        acc[tr][0] = fma(a, b01.x, acc[tr][0]);
        acc[tr][1] = fma(a, b01.y, acc[tr][1]);
        acc[tr][2] = fma(a, b23.x, acc[tr][2]);
        acc[tr][3] = fma(a, b23.y, acc[tr][3]);
      }
    }
  }
  double s = 0;
  for (int i = 0; i < 5; i++) for (int j = 0; j < 4; j++) s += acc[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void dmma_reg(double* out, int iters) {
  double c[8][2];
  for (int i = 0; i < 8; i++) { c[i][0] = i; c[i][1] = -i; }
  const double a = 1.0 + threadIdx.x * 1e-6, b = 1.0 - threadIdx.x * 1e-6;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
  for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
static float time_ms(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f();
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  f();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  double* out;
  cudaMalloc(&out, sizeof(double) * sms * 1024 * 4);
  printf("%s, %d SMs, clock %.0f MHz\n", p.name, sms, p.clockRate / 1e3);
  const int iters = 20000;
  for (int warps : {1, 2, 4, 8, 16, 32}) {
    const int threads = (warps * 32 > 1024) ? 1024 : warps * 32;
    const int blocks = sms * ((warps * 32 + 1023) / 1024);
    float ms = time_ms([&] { dfma_reg<<<blocks, threads>>>(out, iters); });
    double fma = (double)blocks * threads * 20.0 * iters;
    printf("dfma_reg   %2d warps/SM: %8.3f ms  %7.2f TFLOP/s  %6.1f FMA/clk/SM\n", warps, ms, 2 * fma / ms / 1e9, fma / (ms * 1e-3) / sms / (p.clockRate * 1e3));
  }
  for (int warps : {4, 8, 16}) {
    const int threads = warps * 32;
    float ms = time_ms([&] { dfma_smem<<<sms, threads>>>(out, iters / 20); });
    double fma = (double)sms * threads * 20.0 * 24 * (iters / 20);
    printf("dfma_smem  %2d warps/SM: %8.3f ms  %7.2f TFLOP/s  %6.1f FMA/clk/SM\n", warps, ms, 2 * fma / ms / 1e9, fma / (ms * 1e-3) / sms / (p.clockRate * 1e3));
  }
  for (int warps : {1, 2, 4, 8, 16, 32}) {
    const int threads = warps * 32;
    float ms = time_ms([&] { dmma_reg<<<sms, threads>>>(out, iters); });
    double fma = (double)sms * warps * 8.0 * 256 * iters;
    printf("dmma_reg   %2d warps/SM: %8.3f ms  %7.2f TFLOP/s  %6.1f FMA/clk/SM\n", warps, ms, 2 * fma / ms / 1e9, fma / (ms * 1e-3) / sms / (p.clockRate * 1e3));
  }
  return 0;
}
