// Micro-benchmark: do the pipes the K1 generators use overlap?  MUFU alone, MUFU interleaved with FFMA2 (1 : 4),
// the conversion instructions (F2FP to fp16 / e4m3, HADD2.F32), and the generators' per-element mix.
#include <cstdint>
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float *out, long long *cyc, int iters, float seed) {
  float2 p[8];
  float m[8];
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) { p[i] = make_float2(seed + i + threadIdx.x, seed - i); m[i] = seed * (i + 1) * 0.01f; }
  float2 q0 = make_float2(seed * 0.5f, seed * 0.25f), q1 = make_float2(0.999f, 1.001f);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {          // 8 MUFU.EX2
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(m[i]));
    } else if (MODE == 1) {   // 8 MUFU + 32 FFMA2
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(m[i]));
#pragma unroll
        for (int r = 0; r < 4; ++r) p[(i + r) & 7] = __ffma2_rn(p[(i + r) & 7], q1, q0);
      }
    } else if (MODE == 2) {   // 32 FFMA2 only
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int r = 0; r < 4; ++r) p[(i + r) & 7] = __ffma2_rn(p[(i + r) & 7], q1, q0);
    } else if (MODE == 3) {   // 8 x (F2FP.F16 pack + 2 HADD2.F32 back)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const __half2 h = __float22half2_rn(p[i]);
        const float2 f = __half22float2(h);
        p[i] = f;
        acc ^= *reinterpret_cast<const uint32_t *>(&h);
      }
    } else if (MODE == 4) {   // 8 x F2FP e4m3x2
#pragma unroll
      for (int i = 0; i < 8; ++i) acc += (uint32_t)__nv_cvt_float2_to_fp8x2(p[i], __NV_SATFINITE, __NV_E4M3) + it;
    } else if (MODE == 5) {   // 8 x sqrt
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("sqrt.approx.ftz.f32 %0, %0;" : "+f"(m[i]));
    } else if (MODE == 6) {   // per-float2 generator mix: 2 sqrt, 2 ex2, 3 FFMA2/FMUL2 poly, split (F2FP, 2 HADD2, FMUL2, FFMA2, 2 F2FP.E4M3) + 10 FFMA2 distance
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float2 r2 = p[i];
#pragma unroll
        for (int r = 0; r < 10; ++r) r2 = __ffma2_rn(q1, q0, r2);
        float2 rad, ex;
        asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rad.x) : "f"(fabsf(r2.x)));
        asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rad.y) : "f"(fabsf(r2.y)));
        const float2 arg = __fmul2_rn(rad, make_float2(-3.2259955597f, -3.2259955597f));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex.x) : "f"(arg.x));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex.y) : "f"(arg.y));
        const float2 poly = __ffma2_rn(rad, make_float2(2.2360679775f, 2.2360679775f), __ffma2_rn(r2, make_float2(1.6666666667f, 1.6666666667f), make_float2(1.0f, 1.0f)));
        const float2 kv = __fmul2_rn(poly, ex);
        const __half2 h = __float22half2_rn(kv);
        const float2 hf = __half22float2(h);
        const float2 res = __ffma2_rn(hf, make_float2(-4096.0f, -4096.0f), __fmul2_rn(kv, make_float2(4096.0f, 4096.0f)));
        acc ^= *reinterpret_cast<const uint32_t *>(&h);
        acc += (uint32_t)__nv_cvt_float2_to_fp8x2(res, __NV_SATFINITE, __NV_E4M3);
        acc += (uint32_t)__nv_cvt_float2_to_fp8x2(hf, __NV_SATFINITE, __NV_E4M3) << 16;
        p[i] = make_float2(kv.x + 1.0f, kv.y + 2.0f);
      }
    }
  }
  long long t1 = clock64();
  float s = (float)acc;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += p[i].x + p[i].y + m[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE>
void run(const char *name, float *out, long long *cyc) {
  for (int warps : {4, 8, 16}) {
    const int iters = 4000;
    k<MODE><<<148, warps * 32>>>(out, cyc, iters, 1.0f);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double mx = 0;
    for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("%-58s warps/SMSP %d  cycles per loop iteration per warp = %.1f  (x warps/SMSP: %.1f per SMSP-iteration)\n", name, warps / 4, mx / iters, mx / iters / (warps / 4));
  }
}
int main() {
  float *out; long long *cyc;
  cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
  run<0>("8 MUFU.EX2", out, cyc);
  run<5>("8 MUFU.SQRT", out, cyc);
  run<2>("32 FFMA2", out, cyc);
  run<1>("8 MUFU.EX2 + 32 FFMA2 interleaved", out, cyc);
  run<3>("8 x (F2FP.F16x2 + 2 HADD2.F32)", out, cyc);
  run<4>("8 x F2FP.E4M3x2", out, cyc);
  run<6>("8 x generator mix per float2 (32 MUFU, ~136 FMA-pipe, 40 cvt)", out, cyc);
  return 0;
}
