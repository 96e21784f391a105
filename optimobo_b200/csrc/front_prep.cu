// Candidate-independent, once-per-iteration preparation on the device:
// first non-dominated front (reference: pygmo fast_non_dominated_sorting, util_functions.py:64-77),
// exact 2-D / 3-D dominated hypervolume (pymoo HV optimisers.py:217-219, pygmo hypervolume.compute
// util_functions.py:198-199) and the 2-D cell decomposition EMO scores against (emo.py:55-152).
// All inputs are tiny (n <= a few thousand rows): the kernels are latency-bound, single-block where
// an order matters, and sum in the SAME order as the host restatement (optimobo_b200/host_prep.py)
// so the results are bit-identical to it.
#include "common.cuh"

#define PREP_MAX_OBJ 8
#define PREP_THREADS 1024

// mask[j] = 1 unless some row i has Y_i <= Y_j in every objective and Y_i < Y_j in one (minimisation)
__global__ void k_pareto_mask(const double *__restrict__ Y, int n, int k, unsigned char *__restrict__ mask) {
  extern __shared__ double tile[];               // (blockDim.x, k) rows of the sweep
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  double yj[PREP_MAX_OBJ];
  for (int c = 0; c < k; ++c) yj[c] = j < n ? Y[(size_t)j * k + c] : 0.0;
  bool dominated = false;
  for (int i0 = 0; i0 < n; i0 += blockDim.x) {
    const int rows = min((int)blockDim.x, n - i0);
    for (int e = threadIdx.x; e < rows * k; e += blockDim.x) tile[e] = Y[(size_t)i0 * k + e];
    __syncthreads();
    if (j < n && !dominated) {
      for (int ii = 0; ii < rows; ++ii) {
        bool le = true, lt = false;
        for (int c = 0; c < k; ++c) {
          const double v = tile[ii * k + c];
          le = le && (v <= yj[c]);
          lt = lt || (v < yj[c]);
        }
        if (le && lt) { dominated = true; break; }
      }
    }
    __syncthreads();
  }
  if (j < n) mask[j] = dominated ? 0 : 1;
}

// ---- hypervolume -------------------------------------------------------------------------------
// ws layout (doubles): [0,p) a-sorted, [p,2p) b-sorted, [2p,3p) rank3 of the (f1,f2)-sorted entry (as double),
// [3p,4p) f3 by rank3, [4p,5p) slice areas, ws[5p] = number of valid points
__device__ __forceinline__ bool inside(const double *P, int i, int k, const double *r) {
  for (int c = 0; c < k; ++c) if (!(P[(size_t)i * k + c] <= r[c])) return false;
  return true;
}

__global__ void __launch_bounds__(PREP_THREADS, 1)
k_hypervolume(const double *__restrict__ P, int p, int k, double r0, double r1, double r2, double *__restrict__ ws,
              double *__restrict__ out) {
  const double r[3] = {r0, r1, r2};
  double *sa = ws, *sb = ws + p, *sr = ws + 2 * (size_t)p, *q3 = ws + 3 * (size_t)p, *area = ws + 4 * (size_t)p;
  __shared__ int nvalid;
  if (threadIdx.x == 0) nvalid = 0;
  __syncthreads();
  // rank sort (O(p^2), deterministic and stable): points beyond the reference point contribute nothing
  for (int i = threadIdx.x; i < p; i += blockDim.x) {
    if (!inside(P, i, k, r)) continue;
    const double a = P[(size_t)i * k], b = P[(size_t)i * k + 1], z = k == 3 ? P[(size_t)i * k + 2] : 0.0;
    int rank3 = 0;
    if (k == 3) {
      for (int j = 0; j < p; ++j) {
        if (!inside(P, j, k, r)) continue;
        const double zj = P[(size_t)j * k + 2];
        if (zj < z || (zj == z && j < i)) ++rank3;                 // np.argsort(kind="stable") on f3
      }
    }
    int pos = 0;
    for (int j = 0; j < p; ++j) {
      if (j == i || !inside(P, j, k, r)) continue;
      const double aj = P[(size_t)j * k], bj = P[(size_t)j * k + 1];
      bool before = aj < a || (aj == a && bj < b);
      if (!before && aj == a && bj == b) {
        if (k == 3) {                                               // lexsort is stable: tie -> position in the f3 order
          const double zj = P[(size_t)j * k + 2];
          before = zj < z || (zj == z && j < i);
        } else before = j < i;
      }
      if (before) ++pos;
    }
    sa[pos] = a; sb[pos] = b; sr[pos] = (double)rank3;
    if (k == 3) q3[rank3] = z;
    atomicAdd(&nvalid, 1);
  }
  __syncthreads();
  const int nv = nvalid;
  if (k == 2) {
    if (threadIdx.x == 0) {
      double total = 0.0, floor_ = r1;
      for (int s = 0; s < nv; ++s)
        if (sb[s] < floor_) { total = __dadd_rn(total, __dmul_rn(r0 - sa[s], floor_ - sb[s])); floor_ = sb[s]; }   // no FMA: numpy order
      *out = total;
    }
    return;
  }
  // 3-D: slice i (in f3 order) = 2-D hypervolume of the points with rank3 <= i, times the gap to the next f3
  for (int i = threadIdx.x; i < nv; i += blockDim.x) {
    const double top = i + 1 < nv ? q3[i + 1] : r2;
    double a2 = 0.0;
    if (top > q3[i]) {
      double floor_ = r1;
      for (int s = 0; s < nv; ++s)
        if (sr[s] <= (double)i && sb[s] < floor_) { a2 = __dadd_rn(a2, __dmul_rn(r0 - sa[s], floor_ - sb[s])); floor_ = sb[s]; }
    }
    area[i] = a2;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double total = 0.0;
    for (int i = 0; i < nv; ++i) {
      const double top = i + 1 < nv ? q3[i + 1] : r2;
      if (top > q3[i]) total = __dadd_rn(total, __dmul_rn(area[i], top - q3[i]));
    }
    *out = total;
  }
}

// cells (p+1, 2, 2): [c][0] = upper corner, [c][1] = lower corner of the staircase of the f1-sorted front
__global__ void __launch_bounds__(PREP_THREADS, 1)
k_cells_2d(const double *__restrict__ PF, int p, double ideal0, double ideal1, double max0, double max1,
           double *__restrict__ ws, double *__restrict__ cells) {
  double *s0 = ws, *s1 = ws + p;
  for (int i = threadIdx.x; i < p; i += blockDim.x) {
    const double a = PF[2 * (size_t)i];
    int pos = 0;
    for (int j = 0; j < p; ++j) {
      const double aj = PF[2 * (size_t)j];
      if (aj < a || (aj == a && j < i)) ++pos;                      // np.argsort(kind="stable") on f1
    }
    s0[pos] = a; s1[pos] = PF[2 * (size_t)i + 1];
  }
  __syncthreads();
  for (int c = threadIdx.x; c <= p; c += blockDim.x) {
    double *o = cells + 4 * (size_t)c;
    if (c == 0) { o[0] = s0[0]; o[1] = fmax(s1[0], max1); o[2] = ideal0; o[3] = ideal1; }
    else {
      o[0] = c < p ? s0[c] : fmax(s0[p - 1], max0);
      o[1] = s1[c - 1];
      o[2] = s0[c - 1];
      o[3] = ideal1;
    }
  }
}

int ombo_pareto_mask_impl(ombo_ctx *ctx, const double *Y, int n, int k, unsigned char *mask, cudaStream_t s) {
  const int threads = 256;
  k_pareto_mask<<<(n + threads - 1) / threads, threads, (size_t)threads * k * sizeof(double), s>>>(Y, n, k, mask);
  ctx->launches += 1;
  OMBO_CUDA(cudaGetLastError());
  return OMBO_OK;
}

int ombo_hypervolume_impl(ombo_ctx *ctx, const double *P, int p, int k, const double *ref, double *out, cudaStream_t s) {
  int rc = ombo_ws_reserve(&ctx->ws_partial, &ctx->ws_partial_bytes, (5 * (size_t)p + 8) * sizeof(double));
  if (rc) return rc;
  k_hypervolume<<<1, PREP_THREADS, 0, s>>>(P, p, k, ref[0], ref[1], k == 3 ? ref[2] : 0.0, (double *)ctx->ws_partial, out);
  ctx->launches += 1;
  OMBO_CUDA(cudaGetLastError());
  return OMBO_OK;
}

int ombo_cells_2d_impl(ombo_ctx *ctx, const double *PF, int p, const double *ideal, const double *maxp, double *cells,
                       cudaStream_t s) {
  int rc = ombo_ws_reserve(&ctx->ws_partial, &ctx->ws_partial_bytes, (2 * (size_t)p + 8) * sizeof(double));
  if (rc) return rc;
  k_cells_2d<<<1, PREP_THREADS, 0, s>>>(PF, p, ideal[0], ideal[1], maxp[0], maxp[1], (double *)ctx->ws_partial, cells);
  ctx->launches += 1;
  OMBO_CUDA(cudaGetLastError());
  return OMBO_OK;
}
