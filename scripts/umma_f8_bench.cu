// Micro-benchmark: cycles per tcgen05.mma.cta_group::2 (M = 256 over a CTA pair) for the operand formats of the f8c
// kernel -- kind::f16 (K = 16, SWIZZLE_128B) and kind::f8f6f4 e4m3 (K = 32) with SWIZZLE_64B or SWIZZLE_128B
// operand tiles -- alone and in the kernel's 4 + 2 + 2 mix per 64-deep K-block.  Companion of umma_pair_bench.cu.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t sdesc128(uint32_t saddr) {
  const uint32_t lo = ((saddr & 0x3FFFF) >> 4) | (1u << 16);
  const uint32_t hi = 64u | (1u << 14) | (2u << 29);
  return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ uint64_t sdesc64(uint32_t saddr) {
  const uint32_t lo = ((saddr & 0x3FFFF) >> 4) | (1u << 16);
  const uint32_t hi = 32u | (1u << 14) | (4u << 29);
  return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred P1;\nelect.sync _|P1, 0xffffffff;\nselp.u32 %0, 1, 0, P1;\n}\n" : "=r"(pred));
  return pred;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void mma_f16(uint32_t tm, uint64_t da, uint64_t db, uint32_t idesc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tm), "l"(da), "l"(db), "r"(idesc), "r"(1u) : "memory");
}
__device__ __forceinline__ void mma_f8(uint32_t tm, uint64_t da, uint64_t db, uint32_t idesc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n}\n" ::"r"(tm), "l"(da), "l"(db), "r"(idesc), "r"(1u) : "memory");
}
extern __shared__ __align__(1024) unsigned char sm[];
// MODE 0: f16 only (4 per unit)   1: e4m3 SW64 only (4 per unit)   2: e4m3 SW128 only (4 per unit)
//      3: kernel mix 4 x f16 + 4 x e4m3 SW64    4: mix with e4m3 SW128    5: mix, e4m3 planes interleaved in ONE SW128 tile
template <int MODE>
__global__ void __launch_bounds__(128, 1) k(int n_unit, int N, long long *out) {
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  unsigned char *base = sm + ((1024u - (smem_u32(sm) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5;
  uint32_t crank; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(&bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  for (int e = threadIdx.x; e < 192 * 1024 / 4; e += blockDim.x) ((uint32_t *)base)[e] = 0x38383838u;   // fp16 0.52.., e4m3 1.0
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tslot)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::);
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  cluster_sync_all();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tm = tslot;
  if (crank == 0 && warp == 1 && elect_one()) {
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
    const uint32_t a0 = smem_u32(base), b0 = smem_u32(base + 96 * 1024);
#pragma unroll 2
    for (int i = 0; i < n_unit; ++i) {
      const uint32_t st = (uint32_t)(i % 3) * 32 * 1024;
      const uint32_t a = a0 + st, b = b0 + st;
      if (MODE == 0 || MODE >= 3) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) mma_f16(tm, sdesc128(a + ks * 32), sdesc128(b + ks * 32), idesc);
      }
      if (MODE == 1 || MODE == 3) {
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) mma_f8(tm, sdesc64(a + 16384 + ks * 32), sdesc64(b + 16384 + ks * 32), idesc);
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) mma_f8(tm, sdesc64(a + 24576 + ks * 32), sdesc64(b + 24576 + ks * 32), idesc);
      }
      if (MODE == 2 || MODE == 4) {      // two separate SW128 tiles of 128 rows x 128 B, half of each row used
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) mma_f8(tm, sdesc128(a + ks * 32), sdesc128(b + ks * 32), idesc);
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) mma_f8(tm, sdesc128(a + 16384 + ks * 32), sdesc128(b + 16384 + ks * 32), idesc);
      }
      if (MODE == 5) {                    // one SW128 tile: 128 B rows = 4 x 32 e4m3
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) mma_f8(tm, sdesc128(a + 16384 + ks * 32), sdesc128(b + 16384 + ks * 32), idesc);
      }
    }
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(smem_u32(&bar)), "h"((uint16_t)3) : "memory");
  }
  if (warp == 1 && (threadIdx.x & 31) == 0) {
    long long t0 = clock64();
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    out[blockIdx.x] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(tm), "r"(512u));
}
template <int MODE>
static void run(int grid, int n_unit, int N, long long *out, const char *name) {
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 200 * 1024;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, k<MODE>, n_unit, N, out);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, out, grid * 8, cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
  printf("%-44s N=%3d grid=%3d  cycles/unit = %7.1f  (%s)\n", name, N, grid, (double)mx / n_unit, cudaGetErrorString(e));
}
int main() {
  long long *out;
  cudaMalloc(&out, 148 * 8);
  for (int N : {256, 192, 128, 64})
    for (int grid : {2, 148}) {
      run<0>(grid, 3000, N, out, "4 x f16 K=16 SW128 (ideal 4 x N/2)");
      run<1>(grid, 3000, N, out, "4 x e4m3 K=32 SW64 (ideal 4 x N/2)");
      run<2>(grid, 3000, N, out, "4 x e4m3 K=32 SW128 half rows");
      run<3>(grid, 3000, N, out, "mix 4 f16 + 4 e4m3 SW64 (ideal 8 x N/2)");
      run<4>(grid, 3000, N, out, "mix 4 f16 + 4 e4m3 SW128 half rows");
      run<5>(grid, 3000, N, out, "mix 4 f16 + 4 e4m3 one SW128 tile");
    }
  return 0;
}
