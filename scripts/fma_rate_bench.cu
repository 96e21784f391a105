// Micro-benchmark: issue rate of the FP32 forms the K1 generators could use -- FFMA (3 registers), FFMA with a
// constant-bank operand, FFMA2 (packed, 3 register pairs), FFMA2 with a broadcast scalar operand (the form the
// distance loop compiles to) -- for 1, 2 and 4 warps per scheduler.  Prints cycles per warp-instruction per SMSP.
#include <cstdio>
#include <cuda_runtime.h>
__constant__ float cst[256];
template <int MODE>
__global__ void k(float *out, long long *cyc, int iters, float seed) {
  float a[16];
  float2 p[8];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = seed + i + threadIdx.x;
#pragma unroll
  for (int i = 0; i < 8; ++i) p[i] = make_float2(seed + i, seed - i);
  float x0 = seed * 0.5f, x1 = seed * 0.25f;
  float2 q0 = make_float2(x0, x1), q1 = make_float2(x1, x0);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {         // FFMA 3-reg, 16 independent chains
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], x0, x1);
    } else if (MODE == 1) {  // FFMA with a constant operand
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], cst[i + 16 * r], x1);
    } else if (MODE == 2) {  // FFMA2, three register pairs
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i) p[i] = __ffma2_rn(p[i], q0, q1);
    } else if (MODE == 3) {  // FFMA2 with a broadcast scalar multiplier (x, x) * pair + pair
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i) p[i] = __ffma2_rn(make_float2(x0, x0), q0, p[i]);
    } else if (MODE == 4) {  // the distance loop's shape: acc[i] = (x_r, x_r) * t[i] + acc[i], 4 accumulators per x
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i) p[i] = __ffma2_rn(make_float2(a[r + (i >> 2)], a[r + (i >> 2)]), (i & 1) ? q0 : q1, p[i]);
    } else if (MODE == 5) {  // FMUL2 + MUFU mix is not measured here; FADD2
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i) p[i] = __fadd2_rn(p[i], q0);
    }
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
#pragma unroll
  for (int i = 0; i < 8; ++i) s += p[i].x + p[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE>
void run(const char *name, float *out, long long *cyc) {
  for (int warps : {4, 8, 16}) {         // per CTA, one CTA per SM: 1, 2, 4 warps per scheduler
    const int iters = 2000;
    k<MODE><<<148, warps * 32>>>(out, cyc, iters, 1.0f);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double mx = 0;
    for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
    const double inst_per_smsp = (double)iters * 64 * (warps / 4);
    printf("%-46s warps/SMSP %d  cycles per warp-instruction per SMSP = %.2f\n", name, warps / 4, mx / inst_per_smsp);
  }
}
int main() {
  float *out; long long *cyc;
  cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
  float h[256]; for (int i = 0; i < 256; ++i) h[i] = 1.0f + i * 1e-3f;
  cudaMemcpyToSymbol(cst, h, sizeof(h));
  run<0>("FFMA  r, r, r", out, cyc);
  run<1>("FFMA  r, c[], r", out, cyc);
  run<2>("FFMA2 rr, rr, rr", out, cyc);
  run<3>("FFMA2 (x,x), rr, rr  [same x]", out, cyc);
  run<4>("FFMA2 (x_r,x_r), rr, rr  [distance-loop shape]", out, cyc);
  run<5>("FADD2 rr, rr", out, cyc);
  return 0;
}
