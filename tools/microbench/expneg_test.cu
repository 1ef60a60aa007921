#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
#include "../../spamtree_b200/csrc/st_device.cuh"
using namespace st;
__global__ void k(const double* a, double* o, int n) {
  __shared__ double t64[64];
  for (int i = threadIdx.x; i < 64; i += blockDim.x) t64[i] = exp2(-(double)i / 64.0);
  __syncthreads();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) o[i] = exp_neg(t64, a[i]);
}
int main() {
  const int n = 1 << 20;
  double *ha = new double[n], *ho = new double[n], *da, *dout;
  for (int i = 0; i < n; i++) { double u = (i + 0.5) / n; ha[i] = (i % 3 == 0) ? u * 1e-3 : (i % 3 == 1 ? u * 50 : u * 800); }
  ha[0] = 0; ha[1] = 745.2; ha[2] = 1e6; ha[3] = 708.0;
  cudaMalloc(&da, n * 8); cudaMalloc(&dout, n * 8);
  cudaMemcpy(da, ha, n * 8, cudaMemcpyHostToDevice);
  k<<<n / 256, 256>>>(da, dout, n);
  cudaMemcpy(ho, dout, n * 8, cudaMemcpyDeviceToHost);
  double worst = 0; int wi = 0;
  for (int i = 0; i < n; i++) {
    long double ref = expl(-(long double)ha[i]);
    if (ref < 1e-290L) continue;
    double e = fabs((double)((ho[i] - ref) / ref));
    if (e > worst) { worst = e; wi = i; }
  }
  printf("max rel err %.3e at a=%.17g (got %.17g)  exp_neg(0)=%.17g exp_neg(745.2)=%g exp_neg(1e6)=%g\n", worst, ha[wi], ho[wi], ho[0], ho[1], ho[2]);
  return worst < 4e-16 ? 0 : 1;
}
