// Joint posterior samples of a GP over a candidate set: the device side of
// TuRBO's Thompson sampling, `GP.posterior_samples(X_cand, size=batch_size)` (turbo.py:75-117, GPy).
//
//   mu    = K*^T alpha                                   K* = k(X_train, X_cand)
//   Sigma = K** - V V^T + diag_add I                     V = K*^T L^-T  (= (L^-1 K*)^T),  K** = k(X_cand, X_cand)
//   out   = mu + chol(Sigma) Z                           Z (m, S) standard normals supplied by the caller
//
// All FP64 (the m x m covariance of nearby candidates is ill-conditioned).  m <= a few thousand, so the O(m^2 n)
// covariance and the O(m^3 / 3) Cholesky dominate; the Cholesky is the blocked K3 factorisation of gp_refresh.cu
// (ombo_potrf_lower_impl), the products are a 64x64-tiled FP64 GEMM that skips the tiles and K ranges the
// triangular structure makes zero.  Workspace comes from the context (ws_scratch).
#include "common.cuh"

#define JT 64                      // tile edge

__device__ __forceinline__ double joint_kernel(double r2, double sf2, int kernel) {
  if (kernel == OMBO_KERNEL_MATERN52) {
    const double r = sqrt(r2), s5 = 2.23606797749978969641 * r;
    return sf2 * (1.0 + s5 + (5.0 / 3.0) * r2) * exp(-s5);
  }
  return sf2 * exp(-0.5 * r2);
}

// xcs[t][c] = Xc[c][t] / ell[t]   (d, m_pad) SoA, zero padded
__global__ void k_joint_scale(const double *__restrict__ Xc, int m, int d, int m_pad, const double *__restrict__ ell,
                              double *__restrict__ xcs) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= m_pad) return;
  for (int t = 0; t < d; ++t) xcs[(size_t)t * m_pad + c] = c < m ? Xc[(size_t)c * d + t] / ell[t] : 0.0;
}

// K[a][b] = k(A_a, B_b) for a < na, b < nb (SoA inputs with leading dimensions lda / ldb), else 0; optional diagonal:
// + diag_add where a == b (both valid), 1 on the padded part of the diagonal (keeps the Cholesky well defined)
__global__ void k_joint_cov(const double *__restrict__ A, int na, int lda, const double *__restrict__ B, int nb, int ldb,
                            int d, double sf2, int kernel, int rows, int cols, int with_diag, double diag_add,
                            double *__restrict__ K) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  const int a = blockIdx.y * blockDim.y + threadIdx.y;
  if (a >= rows || b >= cols) return;
  double v = 0.0;
  if (a < na && b < nb) {
    double r2 = 0.0;
    for (int t = 0; t < d; ++t) {
      const double df = A[(size_t)t * lda + a] - B[(size_t)t * ldb + b];
      r2 += df * df;
    }
    v = joint_kernel(r2, sf2, kernel);
    if (with_diag && a == b) v += diag_add;
  } else if (with_diag && a == b) {
    v = 1.0;
  }
  K[(size_t)a * cols + b] = v;
}

// mu[c] = sum_i Kx[c][i] alpha[i]   (one warp per row, fixed-order reduction)
__global__ void k_joint_mean(const double *__restrict__ Kx, int rows, int np, const double *__restrict__ alpha,
                             double *__restrict__ mu) {
  const int c = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32, lane = threadIdx.x & 31;
  if (c >= rows) return;
  double acc = 0.0;
  for (int i = lane; i < np; i += 32) acc += Kx[(size_t)c * np + i] * alpha[i];
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) mu[c] = acc;
}

// C (M x N, ldc) = cc * C + cab * A (M x K, lda) * B (N x K, ldb)^T, 64x64 tiles, 4x4 per thread.
//   lower_only: tiles above the diagonal are skipped (symmetric result, only the lower triangle is used)
//   b_lower:    B is lower triangular (B[j][i] = 0 for i > j): the K loop of tile column tj stops at (tj + 1) * 64
__global__ void __launch_bounds__(256)
k_joint_gemm_abt(const double *__restrict__ A, int lda, const double *__restrict__ B, int ldb, double *__restrict__ C,
                 int ldc, int K, double cc, double cab, int lower_only, int b_lower) {
  const int ti = blockIdx.y, tj = blockIdx.x;
  if (lower_only && tj > ti) return;
  constexpr int KS = 32;                         // K step: 2 x 64 x 33 doubles = 33 KB of static shared memory
  __shared__ double As[JT][KS + 1], Bs[JT][KS + 1];
  const int tid = threadIdx.x, tr = tid / 16, tc = tid % 16;
  double acc[4][4] = {};
  const int kend = b_lower ? min(K, (tj + 1) * JT) : K;
  for (int k0 = 0; k0 < kend; k0 += KS) {
    for (int e = tid; e < JT * KS; e += 256) {
      const int r = e / KS, c = e % KS;
      As[r][c] = A[(size_t)(ti * JT + r) * lda + k0 + c];
      Bs[r][c] = B[(size_t)(tj * JT + r) * ldb + k0 + c];
    }
    __syncthreads();
#pragma unroll 4
    for (int t = 0; t < KS; ++t) {
      double av[4], bv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { av[u] = As[tr + 16 * u][t]; bv[u] = Bs[tc + 16 * u][t]; }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[u][v] += av[u] * bv[v];
    }
    __syncthreads();
  }
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      double *p = C + (size_t)(ti * JT + tr + 16 * u) * ldc + (tj * JT + tc + 16 * v);
      *p = (cc == 0.0 ? 0.0 : cc * *p) + cab * acc[u][v];
    }
}

// out[c][s] = mu[c] + sum_{c' <= c} L[c][c'] Z[c'][s]   (one warp per row; lanes stride over c', fixed-order reduce)
__global__ void k_joint_sample(const double *__restrict__ L, int ld, int m, const double *__restrict__ mu,
                               const double *__restrict__ Z, int S, double *__restrict__ out) {
  const int c = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32, lane = threadIdx.x & 31;
  if (c >= m) return;
  for (int s0 = 0; s0 < S; s0 += 4) {
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int j = lane; j <= c; j += 32) {
      const double l = L[(size_t)c * ld + j];
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (s0 + q < S) acc[q] += l * Z[(size_t)j * S + s0 + q];
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      for (int o = 16; o > 0; o >>= 1) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], o);
      if (lane == 0 && s0 + q < S) out[(size_t)c * S + s0 + q] = mu[c] + acc[q];
    }
  }
}

int ombo_joint_samples_impl(ombo_ctx *ctx, const ombo_gp &g, const double *Xc, int m, const double *Z, int S,
                            double diag_add, double *out, cudaStream_t s) {
  const GpDev gp = gp_dev_view(g);
  const int np = gp.n_pad, d = gp.d, mp = ombo_round_up(m, JT);
  const size_t n_xcs = (size_t)d * mp, n_kx = (size_t)mp * np, n_sig = (size_t)mp * mp;
  const size_t n_dinv = (size_t)(mp / JT) * JT * JT;
  const size_t doubles = n_xcs + 2 * n_kx + n_sig + n_dinv + mp + 8;
  int rc = ombo_ws_reserve(&ctx->ws_scratch, &ctx->ws_scratch_bytes, doubles * sizeof(double));
  if (rc) return rc;
  double *xcs = (double *)ctx->ws_scratch, *Kx = xcs + n_xcs, *V = Kx + n_kx, *Sig = V + n_kx, *dinv = Sig + n_sig;
  double *mu = dinv + n_dinv;
  int *status = (int *)(mu + mp);
  OMBO_CUDA(cudaMemsetAsync(status, 0, 16, s));
  k_joint_scale<<<(mp + 255) / 256, 256, 0, s>>>(Xc, m, d, mp, gp.ell, xcs);
  const dim3 tb(16, 16);
  // K*^T (m_pad x n_pad): rows = candidates
  k_joint_cov<<<dim3(np / 16, mp / 16), tb, 0, s>>>(xcs, m, mp, gp.xs, gp.n, np, d, gp.sigma_f2, gp.kernel, mp, np, 0, 0.0, Kx);
  k_joint_mean<<<(mp + 7) / 8, 256, 0, s>>>(Kx, mp, np, gp.alpha, mu);
  // V = K*^T L^-T : V[c][j] = sum_{i <= j} Kx[c][i] Linv[j][i]
  k_joint_gemm_abt<<<dim3(np / JT, mp / JT), 256, 0, s>>>(Kx, np, gp.Linv, np, V, np, np, 0.0, 1.0, 0, 1);
  // Sigma = K** + diag_add I - V V^T (lower triangle)
  k_joint_cov<<<dim3(mp / 16, mp / 16), tb, 0, s>>>(xcs, m, mp, xcs, m, mp, d, gp.sigma_f2, gp.kernel, mp, mp, 1, diag_add, Sig);
  k_joint_gemm_abt<<<dim3(mp / JT, mp / JT), 256, 0, s>>>(V, np, V, np, Sig, mp, np, 1.0, -1.0, 1, 0);
  ctx->launches += 6;
  OMBO_CUDA(cudaGetLastError());
  rc = ombo_potrf_lower_impl(ctx, Sig, mp, m, dinv, status, s);
  if (rc) return rc;
  k_joint_sample<<<(m + 7) / 8, 256, 0, s>>>(Sig, mp, m, mu, Z, S, out);
  ctx->launches += 1;
  OMBO_CUDA(cudaGetLastError());
  int hstatus[4] = {0, 0, 0, 0};
  OMBO_CUDA(cudaMemcpyAsync(hstatus, status, 16, cudaMemcpyDeviceToHost, s));
  OMBO_CUDA(cudaStreamSynchronize(s));
  if (hstatus[0] != 0) {
    ombo_set_error("joint_samples: posterior covariance not positive definite at candidate %d (raise the jitter)",
                   hstatus[0] - 1);
    return OMBO_ERR_NOT_PD;
  }
  return OMBO_OK;
}
