// Micro-benchmark of tcgen05.mma issue patterns on sm_100a (ground truth for DESIGN.md section 6):
// cycles per kind::f16 UMMA (M = 128, K = 16) for different N and accumulator / collector patterns.
// Operands are whatever is in shared memory (values irrelevant), one CTA per SM, one issuing thread.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_sdesc(uint32_t saddr) {
  const uint32_t lo = ((saddr & 0x3FFFF) >> 4) | (1u << 16);
  const uint32_t hi = 64u | (1u << 14) | (2u << 29);
  return ((uint64_t)hi << 32) | lo;
}
#define MMA(QUAL, d, a, b, i, acc)                                                                        \
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16" QUAL          \
               " [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(i), "r"(acc) : "memory")
#define MMA_TS(d, a, b, i, acc)                                                                          \
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16"                 \
               " [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(a), "l"(b), "r"(i), "r"(acc) : "memory")
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred P1;\nelect.sync _|P1, 0xffffffff;\nselp.u32 %0, 1, 0, P1;\n}\n" : "=r"(pred));
  return pred;
}
extern __shared__ __align__(1024) unsigned char sm[];
template <int pattern>
__global__ void __launch_bounds__(128, 1) k(int n_mma, int N, long long *out) {
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  unsigned char *base = sm + ((1024u - (smem_u32(sm) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(&bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  for (int e = threadIdx.x; e < 160 * 1024 / 4; e += blockDim.x) ((uint32_t *)base)[e] = 0x3c003c00u;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tslot)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tm = tslot;
  if (warp == 1 && elect_one()) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t a0 = smem_u32(base), b0 = smem_u32(base + 32 * 1024);
    long long t0 = clock64();
#pragma unroll 12
    for (int i = 0; i < n_mma; ++i) {
      const int ks = i & 3;
      uint32_t d = tm;
      uint32_t a = a0 + ks * 32, b = b0 + ks * 32;
      if (pattern == 1) d = tm + (i & 1) * 256;                 // alternate accumulators every MMA
      if (pattern == 2) d = tm + ((i / 12) & 1) * 256;          // switch every 12
      if (pattern == 6) { b = b0 + ((i / 4) % 3) * 32 * 1024 + ks * 32; }   // 3 different B tiles, same accumulator
      if (pattern == 7) { a = a0 + ((i >> 2) & 1) * 16 * 1024 + ks * 32; }
      const uint64_t da = make_sdesc(a), db = make_sdesc(b);
      if (pattern == 8) { MMA_TS(tm, tm + 256 + ((i >> 2) & 3) * 64 + ks * 8, db, idesc, 1u); }   // A operand in TMEM
      else if (pattern == 9) { MMA_TS(tm + ((i / 12) & 1) * 128, tm + 256 + ((i >> 2) & 3) * 64 + ks * 8, make_sdesc(b0 + ((i / 4) % 3) * 32 * 1024 + ks * 32), idesc, 1u); }
      else if (pattern == 4) { if (i & 1) MMA(".collector::a::lastuse", d, da, db, idesc, 1u); else MMA(".collector::a::fill", d, da, db, idesc, 1u); }
      else MMA("", d, da, db, idesc, 1u);
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    out[blockIdx.x] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tm), "r"(512u));
}
template <int P>
static void launch1(int grid, int n_mma, int N, long long *out) {
  cudaFuncSetAttribute(k<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  k<P><<<grid, 128, 200 * 1024>>>(n_mma, N, out);
}
static void launch(int pattern, int grid, int n_mma, int N, long long *out) {
  switch (pattern) {
    case 0: launch1<0>(grid, n_mma, N, out); break;
    case 1: launch1<1>(grid, n_mma, N, out); break;
    case 2: launch1<2>(grid, n_mma, N, out); break;
    case 4: launch1<4>(grid, n_mma, N, out); break;
    case 6: launch1<6>(grid, n_mma, N, out); break;
    case 7: launch1<7>(grid, n_mma, N, out); break;
    case 8: launch1<8>(grid, n_mma, N, out); break;
    case 9: launch1<9>(grid, n_mma, N, out); break;
  }
}
int main() {
  long long *out;
  cudaMalloc(&out, 148 * 8);
  const int n_mma = 12000;
  struct { int pattern, N; const char *name; } cases[] = {
      {0, 128, "same accumulator, N=128"}, {0, 256, "same accumulator, N=256"}, {0, 64, "same accumulator, N=64"},
      {1, 128, "alternate 2 accumulators each MMA, N=128"}, {2, 128, "switch accumulator every 12, N=128"},
      {4, 128, "same accumulator, A keep/reuse pairs, N=128"}, {6, 128, "same accumulator, 3 B tiles, N=128"},
      {7, 128, "same accumulator, 2 A tiles, N=128"},
      {8, 128, "A in TMEM, N=128"}, {8, 256, "A in TMEM, N=256"}, {8, 64, "A in TMEM, N=64"},
      {9, 128, "A in TMEM, 2 accumulators, 3 B tiles, N=128"}};
  for (auto &c : cases) {
    for (int grid : {1, 148}) {
      launch(c.pattern, grid, n_mma, c.N, out);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[148];
      cudaMemcpy(h, out, grid * 8, cudaMemcpyDeviceToHost);
      long long mx = 0;
      for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
      printf("%-48s grid=%3d  cycles/MMA = %7.1f  (%s)\n", c.name, grid, (double)mx / n_mma, cudaGetErrorString(e));
    }
  }
  return 0;
}
