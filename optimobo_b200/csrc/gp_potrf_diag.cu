// K3, diagonal block: see gp_refresh.cu.  Separate translation unit because of its compile time: the NVVM
// optimiser needs five minutes for the two fully unrolled 64-step loops at -O3 (everything else in the library
// builds in under a minute).  -Xcicc -O1 compiles it in three seconds but leaves the 64-element inverse column
// in local memory: measured 7.9 ms instead of 4.5 ms for the two refreshes of the C5 workload, so -O3 stays.
// (no project headers on purpose: this file must not rebuild when common.cuh changes)
#include <cuda_runtime.h>

#define NB 64      // == OMBO_NB (common.cuh); checked by a static_assert in gp_refresh.cu

// ------------------------------------------------------------------------------------------
// diagonal block: Cholesky of a 64x64 block + inverse of its triangular factor, 64 threads.
// Thread t keeps ROW t of the block in registers (fully unrolled, compile-time register indices); per
// elimination step only the current column travels through shared memory (two barriers of two warps),
// so the 64 dependent steps cost ~100 cycles each instead of three 256-thread barriers + smem sweeps.
__global__ void __launch_bounds__(64) k_potrf_diag(double *__restrict__ A, int ld, int kb, int n,
                                                   double *__restrict__ dinv, int *__restrict__ status) {
  __shared__ double Ls[NB][NB + 1];
  __shared__ double col[NB];
  __shared__ double pivot;
  const int t = threadIdx.x;
  double *blk = A + ((size_t)kb * NB) * ld + (size_t)kb * NB;
  // the block travels through shared memory in both directions: a thread reading ITS row straight from global
  // memory touches 64 different sectors per instruction (row stride ld), 64 times over
  for (int e = t; e < NB * NB; e += NB) Ls[e / NB][e % NB] = blk[(size_t)(e / NB) * ld + (e % NB)];
  __syncthreads();
  double a[NB];
#pragma unroll
  for (int k = 0; k < NB; ++k) a[k] = Ls[t][k];
  __syncthreads();
#pragma unroll
  for (int j = 0; j < NB; ++j) {
    if (t == j) {
      double p = a[j];
      if (!(p > 0.0)) { atomicCAS(status, 0, kb * NB + j + 1); p = 1.0; }
      pivot = sqrt(p);
    }
    __syncthreads();
    const double piv = pivot;
    if (t == j) a[j] = piv;
    if (t > j) a[j] = a[j] / piv;
    col[t] = a[j];                 // column j of L (entries t < j are never read)
    __syncthreads();
    // rows below the pivot: the whole remaining row is updated, also its entries above the diagonal (k > t) -- they
    // are never read (the store below writes zeros there), and leaving the per-element predicate out removes two
    // FSELs and a compare per DFMA from a loop that two warps execute back to back
    if (t > j) {
#pragma unroll
      for (int k = j + 1; k < NB; ++k) a[k] -= a[j] * col[k];
    }
  }
#pragma unroll
  for (int k = 0; k < NB; ++k) Ls[t][k] = (k <= t) ? a[k] : 0.0;
  __syncthreads();
  for (int e = t; e < NB * NB; e += NB) blk[(size_t)(e / NB) * ld + (e % NB)] = Ls[e / NB][e % NB];
  // inverse: thread t owns COLUMN t of X = L^-1 in registers; L is read from smem by broadcast
  double x[NB];
#pragma unroll
  for (int i = 0; i < NB; ++i) {
    double sacc = (i == t) ? 1.0 : 0.0;
    // x[k] = 0 for k < t, so those terms need no predicate (same sum, same order for the entries that count)
#pragma unroll
    for (int k = 0; k < i; ++k) sacc -= Ls[i][k] * x[k];
    x[i] = (i >= t) ? sacc / Ls[i][i] : 0.0;
  }
#pragma unroll
  for (int i = 0; i < NB; ++i) dinv[(size_t)kb * NB * NB + (size_t)i * NB + t] = x[i];
}


void ombo_launch_potrf_diag(double *A, int ld, int kb, int n, double *dinv, int *status, cudaStream_t s) {
  k_potrf_diag<<<1, 64, 0, s>>>(A, ld, kb, n, dinv, status);
}
