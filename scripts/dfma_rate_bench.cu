// Micro-benchmark: FP64 pipe of the B200 outside the tensor core -- DFMA issue rate (16 independent chains, 1 / 2 / 4
// warps per scheduler), the latency of a dependent DFMA chain, of an FP64 division and of an FP64 square root (one
// warp).  Ground truth for the K3 (Cholesky diagonal block) and FP64-mode K1 estimates in DESIGN.md.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(double *out, long long *cyc, int iters, double seed) {
  double a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = seed + i + threadIdx.x * 1e-3;
  double x0 = seed * 0.5, x1 = seed * 0.25;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {          // 64 DFMA, 16 independent chains
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = fma(a[i], x0, x1);
    } else if (MODE == 1) {   // 64 DFMA, one dependent chain
#pragma unroll
      for (int r = 0; r < 64; ++r) a[0] = fma(a[0], x0, x1);
    } else if (MODE == 2) {   // 8 dependent divisions
#pragma unroll
      for (int r = 0; r < 8; ++r) a[0] = x1 / (a[0] + 1.5);
    } else if (MODE == 3) {   // 8 dependent square roots
#pragma unroll
      for (int r = 0; r < 8; ++r) a[0] = sqrt(a[0] + 1.5);
    }
  }
  long long t1 = clock64();
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE>
void run(const char *name, double *out, long long *cyc, int per_iter, bool sweep) {
  for (int warps : {1, 4, 8, 16}) {
    if (!sweep && warps != 1) continue;
    const int iters = 500;
    k<MODE><<<148, warps * 32>>>(out, cyc, iters, 1.0);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double mx = 0;
    for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
    const int per_smsp = warps >= 4 ? warps / 4 : 1;
    printf("%-40s warps/CTA %2d  cycles per op per warp = %.1f   (per SMSP-instruction: %.2f)\n", name, warps,
           mx / ((double)iters * per_iter), mx / ((double)iters * per_iter * per_smsp));
  }
}
int main() {
  double *out; long long *cyc;
  cudaMalloc(&out, 148 * 512 * 8); cudaMalloc(&cyc, 148 * 8);
  run<0>("DFMA, 16 independent chains", out, cyc, 64, true);
  run<1>("DFMA, dependent chain (latency)", out, cyc, 64, false);
  run<2>("FP64 division, dependent (latency)", out, cyc, 8, false);
  run<3>("FP64 sqrt, dependent (latency)", out, cyc, 8, false);
  return 0;
}
