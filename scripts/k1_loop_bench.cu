// Micro-benchmark: the K1 generator loop of posterior_fast8.cu (distances by FFMA2, Matern-5/2 by MUFU, fp16 / e4m3
// planes, swizzled STS) WITHOUT the pipeline around it -- no mbarriers, no TMA, no MMA.  8 or 16 warps per SM, the
// same lane -> (row, column group) map, train slice and operand stage in shared memory.  Prints cycles per K-block and
// the cost of each part by ablation (ABL bit 0: no distance FFMA2s, 1: no MUFU, 2: no conversions, 3: no stores,
// 4: no mean).  Tells how much of the kernel's 3.0 k cycles per block is the instruction stream itself.
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <cuda_runtime.h>
#define FK 64
#define FM 128
__device__ __forceinline__ void sts128(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ void sts64(uint32_t a, uint32_t x, uint32_t y) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};\n" ::"r"(a), "r"(x), "r"(y) : "memory");
}
extern __shared__ __align__(1024) unsigned char smem[];
template <int DP, int R, int GW, int ABL>
#ifndef LBT
#define LBT (GW * 32)
#endif
__global__ void __launch_bounds__(LBT, 1) k(float *out, long long *cyc, int n_blocks, float seed) {
  constexpr int ROWS = R;
  constexpr int XT_STRIDE = (DP + 2) * FK;
  unsigned char *sA = smem;                                   // 3 stages x 32 KB
  float *xt = (float *)(smem + 3 * 32768);                    // 4 slices
  float *xc = xt + 4 * XT_STRIDE;                             // [DP][128]
  const int tid = threadIdx.x, lane = tid & 31, gw = tid >> 5;
  for (int i = tid; i < 4 * XT_STRIDE; i += blockDim.x) xt[i] = seed * 0.01f * (float)((i * 37) % 101) - 0.5f;
  for (int i = tid; i < DP * FM; i += blockDim.x) xc[i] = seed * 0.02f * (float)((i * 53) % 97) - 1.0f;
  __syncthreads();
  const int qq = gw & 3, rg = gw >> 2;
  const int gbit = lane & 1;
  const int g = 2 * qq + gbit;
  const int r16 = ((lane >> 1) & 3) * 2 + ((lane >> 3) & 1) + 8 * (lane >> 4);
  const int row0 = 16 * R * rg + r16;
  uint32_t off_hi[ROWS], off_c8[ROWS];
#pragma unroll
  for (int rr = 0; rr < ROWS; ++rr) {
    const int row = row0 + 16 * rr;
    off_hi[rr] = (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + (((g ^ (row & 7)) & 7) << 4));
    off_c8[rr] = (uint32_t)((row >> 3) * 512 + (row & 7) * 64 + ((((g >> 1) ^ (row >> 1)) & 3) << 4) + (g & 1) * 8);
  }
  const uint32_t sA_u = (uint32_t)__cvta_generic_to_shared(sA);
  float mu_acc[ROWS];
  float x[ROWS][DP], a2[ROWS];
#pragma unroll
  for (int rr = 0; rr < ROWS; ++rr) {
    float acc = 0.f;
    mu_acc[rr] = 0.f;
#pragma unroll
    for (int j = 0; j < DP; ++j) { x[rr][j] = xc[j * FM + row0 + 16 * rr]; acc = fmaf(x[rr][j], x[rr][j], acc); }
    a2[rr] = 0.25f * acc;
  }
  uint32_t sa = 0, sl = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int kb = 0; kb < n_blocks; ++kb) {
    const float *xs = xt + sl * XT_STRIDE + 8 * g;
    const uint32_t st_u = sA_u + sa * 32768;
    float2 r2[ROWS][4];
    {
      const float4 n0 = *(const float4 *)(xs + (DP + 1) * FK);
      const float4 n1 = *(const float4 *)(xs + (DP + 1) * FK + 4);
#pragma unroll
      for (int rr = 0; rr < ROWS; ++rr) {
        const float2 aa = make_float2(a2[rr], a2[rr]);
        r2[rr][0] = __fadd2_rn(aa, make_float2(n0.x, n0.y));
        r2[rr][1] = __fadd2_rn(aa, make_float2(n0.z, n0.w));
        r2[rr][2] = __fadd2_rn(aa, make_float2(n1.x, n1.y));
        r2[rr][3] = __fadd2_rn(aa, make_float2(n1.z, n1.w));
      }
    }
    if (!(ABL & 1)) {
#pragma unroll
      for (int j = 0; j < DP; ++j) {
        const float4 t0_ = *(const float4 *)(xs + j * FK);
        const float4 t1_ = *(const float4 *)(xs + j * FK + 4);
#pragma unroll
        for (int rr = 0; rr < ROWS; ++rr) {
          const float2 xx = make_float2(x[rr][j], x[rr][j]);
          r2[rr][0] = __ffma2_rn(xx, make_float2(t0_.x, t0_.y), r2[rr][0]);
          r2[rr][1] = __ffma2_rn(xx, make_float2(t0_.z, t0_.w), r2[rr][1]);
          r2[rr][2] = __ffma2_rn(xx, make_float2(t1_.x, t1_.y), r2[rr][2]);
          r2[rr][3] = __ffma2_rn(xx, make_float2(t1_.z, t1_.w), r2[rr][3]);
        }
      }
    }
    const float4 al0 = *(const float4 *)(xs + DP * FK);
    const float4 al1 = *(const float4 *)(xs + DP * FK + 4);
#pragma unroll
    for (int rr = 0; rr < ROWS; ++rr) {
      float2 kv[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float2 rad, ex;
        if (!(ABL & 2)) {
          asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rad.x) : "f"(fabsf(r2[rr][e].x)));
          asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rad.y) : "f"(fabsf(r2[rr][e].y)));
        } else { rad = __fmul2_rn(r2[rr][e], make_float2(0.37f, 0.37f)); }
        const float2 arg = __fmul2_rn(rad, make_float2(-3.2259955597f, -3.2259955597f));
        if (!(ABL & 2)) {
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex.x) : "f"(arg.x));
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex.y) : "f"(arg.y));
        } else { ex = __ffma2_rn(arg, make_float2(0.11f, 0.11f), make_float2(0.9f, 0.9f)); }
        const float2 poly = __ffma2_rn(rad, make_float2(2.2360679775f, 2.2360679775f),
                                       __ffma2_rn(r2[rr][e], make_float2(1.6666666667f, 1.6666666667f), make_float2(1.0f, 1.0f)));
        kv[e] = __fmul2_rn(poly, ex);
      }
      if (!(ABL & 16)) {
        float2 m2 = __fmul2_rn(kv[0], make_float2(al0.x, al0.y));
        m2 = __ffma2_rn(kv[1], make_float2(al0.z, al0.w), m2);
        m2 = __ffma2_rn(kv[2], make_float2(al1.x, al1.y), m2);
        m2 = __ffma2_rn(kv[3], make_float2(al1.z, al1.w), m2);
        mu_acc[rr] += m2.x + m2.y;
      }
      uint32_t hi[4], c1[2], c2[2];
      if (!(ABL & 4)) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const __half2 h = __float22half2_rn(kv[e]);
          const float2 hf = __half22float2(h);
          const float2 res = __ffma2_rn(hf, make_float2(-4096.0f, -4096.0f), __fmul2_rn(kv[e], make_float2(4096.0f, 4096.0f)));
          hi[e] = *reinterpret_cast<const uint32_t *>(&h);
          const uint32_t l8 = (uint32_t)__nv_cvt_float2_to_fp8x2(res, __NV_SATFINITE, __NV_E4M3);
          const uint32_t a8 = (uint32_t)__nv_cvt_float2_to_fp8x2(hf, __NV_SATFINITE, __NV_E4M3);
          if (e & 1) { c1[e >> 1] |= l8 << 16; c2[e >> 1] |= a8 << 16; }
          else { c1[e >> 1] = l8; c2[e >> 1] = a8; }
        }
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) hi[e] = __float_as_uint(kv[e].x) ^ __float_as_uint(kv[e].y);
        c1[0] = hi[0] + hi[1]; c1[1] = hi[2] + hi[3]; c2[0] = hi[0] ^ hi[2]; c2[1] = hi[1] ^ hi[3];
      }
      if (!(ABL & 8)) {
        sts128(st_u + off_hi[rr], hi[0], hi[1], hi[2], hi[3]);
        sts64(st_u + 16384 + off_c8[rr], c1[0], c1[1]);
        sts64(st_u + 24576 + off_c8[rr], c2[0], c2[1]);
      } else {
        mu_acc[rr] += __uint_as_float((hi[0] ^ hi[1] ^ hi[2] ^ hi[3] ^ c1[0] ^ c1[1] ^ c2[0] ^ c2[1]) & 0x3f800000u);
      }
    }
    __syncwarp();
    if (++sa == 3) sa = 0;
    if (++sl == 4) sl = 0;
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int rr = 0; rr < ROWS; ++rr) s += mu_acc[rr];
  out[blockIdx.x * blockDim.x + tid] = s + (float)sA[tid];
  if (tid == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int DP, int R, int GW, int ABL>
void run(const char *name, float *out, long long *cyc) {
  const size_t sm = 3 * 32768 + 4 * (DP + 2) * FK * 4 + DP * FM * 4;
  cudaFuncSetAttribute(k<DP, R, GW, ABL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  const int nb = 4000;
  k<DP, R, GW, ABL><<<148, GW * 32, sm>>>(out, cyc, nb, 1.0f);
  k<DP, R, GW, ABL><<<148, GW * 32, sm>>>(out, cyc, nb, 1.0f);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double mx = 0;
  for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
  printf("%-52s GW=%2d R=%d  cycles per K-block (128 x 64 values) = %7.1f   %s\n", name, GW, R, mx / nb, cudaGetErrorString(e));
}
int main() {
  float *out; long long *cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  run<10, 4, 8, 0>("full loop body", out, cyc);
  run<10, 2, 16, 0>("full loop body", out, cyc);
  run<10, 4, 8, 1>("- distance FFMA2s", out, cyc);
  run<10, 4, 8, 2>("- MUFU (sqrt, ex2 -> FMA stand-ins)", out, cyc);
  run<10, 4, 8, 4>("- conversions (fp16 / e4m3 / residual)", out, cyc);
  run<10, 4, 8, 8>("- stores", out, cyc);
  run<10, 4, 8, 16>("- mean", out, cyc);
  run<10, 4, 8, 3>("- distances - MUFU", out, cyc);
  run<10, 4, 8, 6>("- MUFU - conversions", out, cyc);
  run<10, 4, 8, 5>("- distances - conversions", out, cyc);
  run<10, 4, 8, 12>("- conversions - stores", out, cyc);
  run<10, 2, 16, 2>("- MUFU", out, cyc);
  run<10, 2, 16, 1>("- distance FFMA2s", out, cyc);
  return 0;
}
