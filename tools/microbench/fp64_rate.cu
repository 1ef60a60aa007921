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


// the sm_90+ DMMA shapes (fewer issue slots and fragment loads per FMA): do they run at rate on sm_100a?
__global__ void dmma_m16n8k4(double* out, int iters) {
  double c[4][4];
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) c[i][j] = i - j;
  const double a0 = 1.0 + threadIdx.x * 1e-6, a1 = 1.0 - threadIdx.x * 2e-6, b = 1.0 - threadIdx.x * 1e-6;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 4; i++)
      asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a0), "d"(a1), "d"(b));
  }
  double s = 0;
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void dmma_m16n8k8(double* out, int iters) {
  double c[4][4];
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) c[i][j] = i - j;
  const double a0 = 1.0 + threadIdx.x * 1e-6, a1 = 1.0 - threadIdx.x * 2e-6, b = 1.0 - threadIdx.x * 1e-6;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 4; i++)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a0), "d"(a1), "d"(a1), "d"(a0), "d"(b), "d"(a0));
  }
  double s = 0;
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void dmma_m16n8k16(double* out, int iters) {
  double c[4][4];
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) c[i][j] = i - j;
  const double a0 = 1.0 + threadIdx.x * 1e-6, a1 = 1.0 - threadIdx.x * 2e-6, b = 1.0 - threadIdx.x * 1e-6;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 4; i++)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                   : "d"(a0), "d"(a1), "d"(a1), "d"(a0), "d"(a0), "d"(a1), "d"(a1), "d"(a0), "d"(b), "d"(a0), "d"(a1), "d"(b));
  }
  double s = 0;
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// DMMA fed from shared memory the way the BUILD sweeps feed it: per k-step of 4, one A fragment (8x4, LDS.64 each lane) and
// one B fragment feeding NM m-tiles.  SHAPE 0: m8n8k4 with 2 m-tiles per B (the shipped loop); SHAPE 1: m16n8k8
template <int SHAPE>
__global__ void dmma_smem(double* out, int iters) {
  extern __shared__ __align__(16) double dyn_smem[];
  double* As = dyn_smem;             // 16 staged rows, SA = 212 (= 4 mod 16)
  double* Bs = dyn_smem + 16 * 212;  // panel 208 x (104 + 4)
  for (int i = threadIdx.x; i < 16 * 212; i += blockDim.x) As[i] = 1e-3 * (i % 97);
  for (int i = threadIdx.x; i < 208 * 108; i += blockDim.x) Bs[i] = 1e-4 * (i % 89);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nt = warp % 13;
  double acc[4] = {0, 0, 0, 0}, acc2[4] = {0, 0, 0, 0};
  for (int it = 0; it < iters; it++) {
    if (SHAPE == 0) {
      const double* ap = As + (lane >> 2) * 212 + (lane & 3);
      const double* bp = Bs + (lane & 3) * 108 + 8 * nt + (lane >> 2);
#pragma unroll 2
      for (int kk = 0; kk < 208; kk += 8) {
        const double b0 = bp[kk * 108], b1 = bp[(kk + 4) * 108];
        const double a1 = ap[8 * 212 + kk], a3 = ap[8 * 212 + kk + 4], a0 = ap[kk], a2 = ap[kk + 4];
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(acc[0]), "+d"(acc[1]) : "d"(a0), "d"(b0));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(acc[2]), "+d"(acc[3]) : "d"(a1), "d"(b0));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(acc2[0]), "+d"(acc2[1]) : "d"(a2), "d"(b1));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(acc2[2]), "+d"(acc2[3]) : "d"(a3), "d"(b1));
      }
    } else {
      // m16n8k8: A 16x8 row: a0=(g, t) a1=(g+8, t) a2=(g, t+4) a3=(g+8, t+4); B 8x8 col: b0=(t, g) b1=(t+4, g); g = lane>>2, t = lane&3
      const double* ap = As + (lane >> 2) * 212 + (lane & 3);
      const double* bp = Bs + (lane & 3) * 108 + 8 * nt + (lane >> 2);
#pragma unroll 2
      for (int kk = 0; kk < 208; kk += 16) {
        const double b0 = bp[kk * 108], b1 = bp[(kk + 4) * 108], b2 = bp[(kk + 8) * 108], b3 = bp[(kk + 12) * 108];
        const double a0 = ap[kk], a1 = ap[8 * 212 + kk], a2 = ap[kk + 4], a3 = ap[8 * 212 + kk + 4];
        const double e0 = ap[kk + 8], e1 = ap[8 * 212 + kk + 8], e2 = ap[kk + 12], e3 = ap[8 * 212 + kk + 12];
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+d"(acc[0]), "+d"(acc[1]), "+d"(acc[2]), "+d"(acc[3]) : "d"(a0), "d"(a1), "d"(a2), "d"(a3), "d"(b0), "d"(b1));
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+d"(acc2[0]), "+d"(acc2[1]), "+d"(acc2[2]), "+d"(acc2[3]) : "d"(e0), "d"(e1), "d"(e2), "d"(e3), "d"(b2), "d"(b3));
      }
    }
  }
  double s = 0;
  for (int i = 0; i < 4; i++) s += acc[i] + acc2[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// Are the DMMA datapath and the FP64 FMA datapath separate units?  Even warps issue DMMA, odd warps DFMA (registers only);
// if the SM total exceeds 64 FMA/clk the two can be overlapped.
__global__ void mixed_reg(double* out, int iters, int dmma_warps_of_4) {
  const int warp = threadIdx.x >> 5;
  double s = 0;
  if ((warp & 3) < dmma_warps_of_4) {
    double c[8][2];
    for (int i = 0; i < 8; i++) { c[i][0] = i; c[i][1] = -i; }
    const double a = 1.0 + threadIdx.x * 1e-6, b = 1.0 - threadIdx.x * 1e-6;
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int i = 0; i < 8; i++)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
  } else {
    double a[16];
    for (int i = 0; i < 16; i++) a[i] = threadIdx.x * 1e-3 + i;
    const double b = 1.0000001, c = 1e-9;
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int i = 0; i < 16; i++) a[i] = fma(a[i], b, c);
    }
    for (int i = 0; i < 16; i++) s += a[i];
  }
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
  for (int warps : {1, 4, 8, 16}) {
    const int threads = warps * 32;
    float ms = time_ms([&] { dmma_m16n8k4<<<sms, threads>>>(out, iters); });
    double fma = (double)sms * warps * 4.0 * 512 * iters;
    printf("dmma_m16n8k4  %2d warps/SM: %8.3f ms  %7.2f TFLOP/s  %6.1f FMA/clk/SM\n", warps, ms, 2 * fma / ms / 1e9, fma / (ms * 1e-3) / sms / (p.clockRate * 1e3));
    ms = time_ms([&] { dmma_m16n8k8<<<sms, threads>>>(out, iters); });
    fma = (double)sms * warps * 4.0 * 1024 * iters;
    printf("dmma_m16n8k8  %2d warps/SM: %8.3f ms  %7.2f TFLOP/s  %6.1f FMA/clk/SM\n", warps, ms, 2 * fma / ms / 1e9, fma / (ms * 1e-3) / sms / (p.clockRate * 1e3));
    ms = time_ms([&] { dmma_m16n8k16<<<sms, threads>>>(out, iters / 2); });
    fma = (double)sms * warps * 4.0 * 2048 * (iters / 2);
    printf("dmma_m16n8k16 %2d warps/SM: %8.3f ms  %7.2f TFLOP/s  %6.1f FMA/clk/SM\n", warps, ms, 2 * fma / ms / 1e9, fma / (ms * 1e-3) / sms / (p.clockRate * 1e3));
  }
  const size_t SM = (16 * 212 + 208 * 108) * sizeof(double);
  cudaFuncSetAttribute(dmma_smem<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM);
  cudaFuncSetAttribute(dmma_smem<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM);
  for (int warps : {4, 8, 13, 16}) {
    const int threads = warps * 32;
    float ms = time_ms([&] { dmma_smem<0><<<sms, threads, SM>>>(out, iters / 20); });
    double fma = (double)sms * warps * 26.0 * 4 * 256 * (iters / 20);
    printf("dmma_smem m8n8k4  %2d warps/SM: %8.3f ms  %7.2f TFLOP/s  %6.1f FMA/clk/SM\n", warps, ms, 2 * fma / ms / 1e9, fma / (ms * 1e-3) / sms / (p.clockRate * 1e3));
    ms = time_ms([&] { dmma_smem<1><<<sms, threads, SM>>>(out, iters / 20); });
    fma = (double)sms * warps * 13.0 * 2 * 1024 * (iters / 20);
    printf("dmma_smem m16n8k8 %2d warps/SM: %8.3f ms  %7.2f TFLOP/s  %6.1f FMA/clk/SM\n", warps, ms, 2 * fma / ms / 1e9, fma / (ms * 1e-3) / sms / (p.clockRate * 1e3));
  }
  for (int warps : {8, 16})
    for (int nd : {0, 1, 2, 3, 4}) {
      const int threads = warps * 32;
      float ms = time_ms([&] { mixed_reg<<<sms, threads>>>(out, iters, nd); });
      // per iteration: a DMMA warp does 8 x 256 FMA, a DFMA warp 16 x 32 FMA
      const double wd = warps * nd / 4.0, wf = warps - wd;
      double fma = (double)sms * (wd * 8.0 * 256 + wf * 16.0 * 32) * iters;
      printf("mixed_reg %2d warps/SM, %d of 4 DMMA: %8.3f ms  %7.2f TFLOP/s  %6.1f FMA/clk/SM (dmma part %.1f, dfma part %.1f)\n", warps, nd, ms, 2 * fma / ms / 1e9,
             fma / (ms * 1e-3) / sms / (p.clockRate * 1e3), sms * wd * 8.0 * 256 * iters / (ms * 1e-3) / sms / (p.clockRate * 1e3),
             sms * wf * 16.0 * 32 * iters / (ms * 1e-3) / sms / (p.clockRate * 1e3));
    }
  return 0;
}
