// K3, diagonal block: see gp_refresh.cu.  Separate translation unit because of its compile time: the NVVM
// optimiser needs five minutes for the two fully unrolled 64-step loops at -O3 (everything else in the library
// builds in under a minute).  -Xcicc -O1 compiles it in three seconds but leaves the 64-element inverse column
// in local memory: measured 7.9 ms instead of 4.5 ms for the two refreshes of the C5 workload, so -O3 stays.
// (no project headers on purpose: this file must not rebuild when common.cuh changes)
#include <cuda_runtime.h>

#define NB 64      // == OMBO_NB (common.cuh); checked by a static_assert in gp_refresh.cu

// ------------------------------------------------------------------------------------------
// diagonal block: Cholesky of a 64x64 block + inverse of its triangular factor, 128 threads in two roles.
// Warps 0-1 (thread t = ROW t of the block in registers, fully unrolled, compile-time register indices): per
// elimination step only the current column travels through shared memory (two named barriers of two warps), so the
// 64 dependent steps cost a square root + a division + ~100 cycles each.
// Warps 2-3 (thread t = COLUMN t of X = L^-1 in registers): row i of X needs row i of L, which is final after
// elimination step i -- so the substitution runs one step behind the factorisation instead of after it (a counter in
// shared memory says how many rows of L are complete).  Both are latency chains of the same length (sqrt 100 + div 132
// cycles per pivot against up to 63 dependent DFMAs + a division per row), and overlapped they cost one of them.
// Every sum keeps its order: L and L^-1 are bit-identical to the one-after-the-other kernel.
__device__ __forceinline__ void potrf_bar64() { asm volatile("bar.sync 1, 64;\n" ::: "memory"); }

__global__ void __launch_bounds__(128) k_potrf_diag(double *__restrict__ A, int ld, int kb, int n,
                                                    double *__restrict__ dinv, int *__restrict__ status) {
  __shared__ double Ls[NB][NB + 1];
  __shared__ double col[NB];
  __shared__ double pivot;
  __shared__ int rows_done;
  const int tid = threadIdx.x;
  double *blk = A + ((size_t)kb * NB) * ld + (size_t)kb * NB;
  // the block travels through shared memory in both directions: a thread reading ITS row straight from global
  // memory touches 64 different sectors per instruction (row stride ld), 64 times over
  for (int e = tid; e < NB * NB; e += 128) Ls[e / NB][e % NB] = blk[(size_t)(e / NB) * ld + (e % NB)];
  if (tid == 0) rows_done = 0;
  __syncthreads();
  if (tid < NB) {
    const int t = tid;
    double a[NB];
#pragma unroll
    for (int k = 0; k < NB; ++k) a[k] = Ls[t][k];
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      if (t == j) {
        double p = a[j];
        if (!(p > 0.0)) { atomicCAS(status, 0, kb * NB + j + 1); p = 1.0; }
        pivot = sqrt(p);
      }
      potrf_bar64();
      const double piv = pivot;
      if (t == j) a[j] = piv;
      if (t > j) a[j] = a[j] / piv;
      col[t] = a[j];                 // column j of L (entries t < j are never read)
      if (t >= j) Ls[t][j] = a[j];   // ... and into the factor the other two warps read row by row
      potrf_bar64();
      if (t == 0) { __threadfence_block(); *(volatile int *)&rows_done = j + 1; }   // row j of L is complete
      // rows below the pivot: the whole remaining row is updated, also its entries above the diagonal (k > t) -- they
      // are never read (zeros are stored there below), and leaving the per-element predicate out removes two
      // FSELs and a compare per DFMA from a loop that two warps execute back to back
      if (t > j) {
#pragma unroll
        for (int k = j + 1; k < NB; ++k) a[k] -= a[j] * col[k];
      }
    }
    // zeros above the diagonal (the staging copy of A is still there), then the coalesced store of L
#pragma unroll
    for (int k = 0; k < NB; ++k)
      if (k > t) Ls[t][k] = 0.0;
    potrf_bar64();
    for (int e = t; e < NB * NB; e += NB) blk[(size_t)(e / NB) * ld + (e % NB)] = Ls[e / NB][e % NB];
  } else {
    // inverse: thread t owns COLUMN t of X = L^-1 in registers; L is read from smem by broadcast
    const int t = tid - NB;
    double x[NB];
#pragma unroll
    for (int i = 0; i < NB; ++i) {
      while (*(volatile int *)&rows_done <= i) {}
      __threadfence_block();          // (also keeps the loads of row i behind the poll)
      double sacc = (i == t) ? 1.0 : 0.0;
      // x[k] = 0 for k < t, so those terms need no predicate (same sum, same order for the entries that count)
#pragma unroll
      for (int k = 0; k < i; ++k) sacc -= Ls[i][k] * x[k];
      x[i] = (i >= t) ? sacc / Ls[i][i] : 0.0;
    }
#pragma unroll
    for (int i = 0; i < NB; ++i) dinv[(size_t)kb * NB * NB + (size_t)i * NB + t] = x[i];
  }
}


void ombo_launch_potrf_diag(double *A, int ld, int kb, int n, double *dinv, int *status, cudaStream_t s) {
  k_potrf_diag<<<1, 128, 0, s>>>(A, ld, kb, n, dinv, status);
}
