// Marginal likelihood and its gradient on the device: replaces the O(n^3)
// work inside GPy's `model.optimize(max_f_eval=1000)` (optimisers.py:230 and the sites of section 0.1).
// The L-BFGS iteration itself stays on the host (scipy); every function evaluation is
//   K3 refresh (L, L^-1, alpha)  ->  W = L^-T L^-1 - alpha alpha^T  ->  g_k = 1/2 sum_ij W_ij dK_ij/dtheta_k
// with theta = (log sigma_f2, log ell_1..d), all FP64.
#include "common.cuh"

// W = Linv^T Linv - alpha alpha^T  (full symmetric n_pad x n_pad; Linv is lower triangular)
__global__ void __launch_bounds__(256) k_kinv_minus_aat(const double *__restrict__ Linv, const double *__restrict__ alpha,
                                                        int np, double *__restrict__ W) {
  __shared__ double As[16][65];
  __shared__ double Bs[16][65];
  const int a0 = blockIdx.y * 64, b0 = blockIdx.x * 64;
  const int tr = threadIdx.x / 16, tc = threadIdx.x % 16;
  double acc[4][4] = {};
  const int j_start = max(a0, b0);              // rows j < max(a, b) contribute zeros
  for (int j0 = j_start; j0 < np; j0 += 16) {
    for (int e = threadIdx.x; e < 16 * 64; e += 256) {
      const int t = e / 64, c = e % 64;
      As[t][c] = Linv[(size_t)(j0 + t) * np + a0 + c];
      Bs[t][c] = Linv[(size_t)(j0 + t) * np + b0 + c];
    }
    __syncthreads();
#pragma unroll
    for (int t = 0; t < 16; ++t) {
      double av[4], bv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { av[u] = As[t][tr + 16 * u]; bv[u] = Bs[t][tc + 16 * u]; }
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
      const int a = a0 + tr + 16 * u, b = b0 + tc + 16 * v;
      W[(size_t)a * np + b] = acc[u][v] - alpha[a] * alpha[b];
    }
}

// per-block partial sums of W_ij * dK_ij/dtheta_k, k = 0..d (k = 0: log sigma_f2)
__global__ void __launch_bounds__(256) k_nlml_grad_partial(const double *__restrict__ W, const double *__restrict__ xs,
                                                           int n, int d, int np, double sf2, int kernel,
                                                           double *__restrict__ partial) {
  __shared__ double red[8][OMBO_MAX_DIM + 1];
  const int j = blockIdx.x * 16 + (threadIdx.x & 15);
  const int i = blockIdx.y * 16 + (threadIdx.x >> 4);
  double g[OMBO_MAX_DIM + 1];
  for (int k = 0; k <= d; ++k) g[k] = 0.0;
  if (i < n && j < n) {
    double diff2[OMBO_MAX_DIM], r2 = 0.0;
    for (int k = 0; k < d; ++k) {
      const double df = xs[(size_t)k * np + i] - xs[(size_t)k * np + j];
      diff2[k] = df * df;
      r2 += diff2[k];
    }
    const double w = W[(size_t)i * np + j];
    double k0, dk;
    if (kernel == OMBO_KERNEL_MATERN52) {
      const double s5 = 2.23606797749978969641, r = sqrt(r2), e = exp(-s5 * r);
      k0 = (1.0 + s5 * r + (5.0 / 3.0) * r2) * e;
      dk = (5.0 / 3.0) * (1.0 + s5 * r) * e;
    } else {
      k0 = exp(-0.5 * r2);
      dk = k0;
    }
    g[0] = w * sf2 * k0;
    for (int k = 0; k < d; ++k) g[1 + k] = w * sf2 * dk * diff2[k];
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int k = 0; k <= d; ++k) {
    double v = g[k];
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp][k] = v;
  }
  __syncthreads();
  if (threadIdx.x <= d) {
    double v = 0.0;
    for (int w8 = 0; w8 < 8; ++w8) v += red[w8][threadIdx.x];
    partial[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * (OMBO_MAX_DIM + 1) + threadIdx.x] = v;
  }
}

// out[0] = NLML, out[1 + k] = dNLML/dtheta_k; fixed summation order (bit-reproducible)
__global__ void __launch_bounds__(256) k_nlml_finish(const double *__restrict__ partial, int n_blocks, int d,
                                                     const double *__restrict__ L, const double *__restrict__ alpha,
                                                     const double *__restrict__ y, int n, int np, double *__restrict__ out) {
  __shared__ double red[256];
  for (int k = 0; k <= d + 1; ++k) {
    double v = 0.0;
    if (k <= d) {
      for (int b = threadIdx.x; b < n_blocks; b += 256) v += partial[(size_t)b * (OMBO_MAX_DIM + 1) + k];
    } else {
      for (int i = threadIdx.x; i < n; i += 256) v += 0.5 * y[i] * alpha[i] + log(L[(size_t)i * np + i]);
    }
    red[threadIdx.x] = v;
    __syncthreads();
    for (int o = 128; o; o >>= 1) {
      if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      if (k <= d) out[1 + k] = 0.5 * red[0];
      else out[0] = red[0] + 0.5 * (double)n * 1.83787706640934548356;   // log(2 pi)
    }
    __syncthreads();
  }
}

int ombo_nlml_grad_impl(ombo_ctx *ctx, const ombo_gp_spec *sp, void *state, double *out_host, cudaStream_t s) {
  int rc = ombo_refresh_impl(ctx, sp, state, s);      // also reports OMBO_ERR_NOT_PD
  if (rc) return rc;
  const GpLayout lay = gp_layout(sp->n, sp->d);
  const int n = sp->n, d = sp->d, np = lay.n_pad;
  char *b = (char *)state;
  double *L = (double *)(b + lay.off_L), *Linv = (double *)(b + lay.off_Linv);
  double *alpha = (double *)(b + lay.off_alpha), *xs = (double *)(b + lay.off_xs);
  const dim3 gg((n + 15) / 16, (n + 15) / 16);
  const int n_blocks = gg.x * gg.y;
  // workspace: W (np x np) + partials + out
  size_t want = (size_t)np * np * 8 + (size_t)n_blocks * (OMBO_MAX_DIM + 1) * 8 + 64 * 8;
  // stream-ordered allocation (the device's default pool keeps freed blocks: ombo_ctx_create raises its release
  // threshold), not the context's scratch: likelihood evaluations of different GPs run concurrently on different
  // streams and host threads (one fit per objective), and each needs its own W
  double *W = nullptr;
  OMBO_CUDA(cudaMallocAsync((void **)&W, want, s));
  double *partial = W + (size_t)np * np;
  double *out = partial + (size_t)n_blocks * (OMBO_MAX_DIM + 1);
  k_kinv_minus_aat<<<dim3(np / 64, np / 64), 256, 0, s>>>(Linv, alpha, np, W);
  k_nlml_grad_partial<<<gg, 256, 0, s>>>(W, xs, n, d, np, sp->sigma_f2, sp->kernel, partial);
  k_nlml_finish<<<1, 256, 0, s>>>(partial, n_blocks, d, L, alpha, sp->y, n, np, out);
  ctx->launches += 3;
  OMBO_CUDA(cudaGetLastError());
  OMBO_CUDA(cudaMemcpyAsync(out_host, out, (size_t)(d + 2) * 8, cudaMemcpyDeviceToHost, s));
  OMBO_CUDA(cudaFreeAsync(W, s));
  OMBO_CUDA(cudaStreamSynchronize(s));
  return OMBO_OK;
}
