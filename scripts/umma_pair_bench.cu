// Micro-benchmark: cycles per tcgen05.mma.cta_group::2.kind::f16 (M = 256 over a CTA pair, K = 16) for N = 64..256,
// operands in shared memory (SS) -- companion of umma_bench.cu.  One cluster of 2 CTAs per TPC, leader issues.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_sdesc(uint32_t saddr) {
  const uint32_t lo = ((saddr & 0x3FFFF) >> 4) | (1u << 16);
  const uint32_t hi = 64u | (1u << 14) | (2u << 29);
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
extern __shared__ __align__(1024) unsigned char sm[];
template <int NB>   // NB: number of distinct B tiles cycled through (1 or 3)
__global__ void __launch_bounds__(128, 1) k(int n_mma, int N, long long *out) {
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  unsigned char *base = sm + ((1024u - (smem_u32(sm) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5;
  uint32_t crank; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(&bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  for (int e = threadIdx.x; e < 160 * 1024 / 4; e += blockDim.x) ((uint32_t *)base)[e] = 0x3c003c00u;
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
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
    const uint32_t a0 = smem_u32(base), b0 = smem_u32(base + 32 * 1024);
    long long t0 = clock64();
#pragma unroll 12
    for (int i = 0; i < n_mma; ++i) {
      const int ks = i & 3;
      const uint32_t a = a0 + ks * 32, b = b0 + (NB == 1 ? 0 : ((i >> 2) % NB) * 32 * 1024) + ks * 32;
      const uint64_t da = make_sdesc(a), db = make_sdesc(b);
      asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tm), "l"(da), "l"(db), "r"(idesc), "r"(1u) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(smem_u32(&bar)), "h"((uint16_t)3) : "memory");
  }
  // both CTAs wait for the commit (it is multicast)
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
template <int NB>
static void run(int grid, int n_mma, int N, long long *out, const char *name) {
  cudaFuncSetAttribute(k<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 200 * 1024;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, k<NB>, n_mma, N, out);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, out, grid * 8, cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
  printf("%-40s N=%3d grid=%3d  cycles/MMA = %7.1f  (%s)\n", name, N, grid, (double)mx / n_mma, cudaGetErrorString(e));
}
int main() {
  long long *out;
  cudaMalloc(&out, 148 * 8);
  for (int N : {256, 128, 64})
    for (int grid : {2, 148}) {
      run<1>(grid, 12000, N, out, "pair M=256, one B tile");
      run<3>(grid, 12000, N, out, "pair M=256, 3 B tiles");
    }
  return 0;
}
