// K1 + K2, FP64 "reference-tolerance" mode (rtol 1e-6 against the reference's numpy path).
// Replaces model.predict(Xnew) -> (mean, var) of GPy GPRegression (reference call sites:
// util_functions.py:19,156,188,309; optimisers.py:336; parego.py:137; cparego.py:63,461,479;
// keep.py:129,149; emo.py:204) for m candidates at once.
//
// One persistent CTA owns a tile of 64 candidates at a time:
//   phase A (K1): k*[c][i] = sigma_f2 * k(|| (x_c - x_i)/ell ||) by direct differences in FP64,
//                 written to a CTA-private scratch tile [64][n_pad] that stays L2-resident
//                 (grid * 64 * n_pad * 8 B  ~ 75-150 MB at n_pad = 1024), and the mean
//                 mu_c = sum_i k*[c][i] alpha[i] accumulated on the way;
//   phase B (K2): V = K* . L^-T on the FP64 tensor pipe (mma.sync m8n8k4 DMMA), 128 columns of
//                 L^-1 at a time, K-loop stopped at the diagonal (L^-1 is lower triangular, so
//                 only ~n^2/2 MACs per candidate are issued), and inside the diagonal 128 x 128 block every
//                 warp stops at the end of ITS 32 columns -- with the warps mapped so that each of the four
//                 schedulers hosts one short and one long column group; the epilogue squares and
//                 row-reduces the accumulators, so only sum_j v_j^2 per candidate survives:
//                 var = max(sigma_f2 - sum_j v_j^2, floor) + sigma_n2.
// tcgen05 has no f64 kind, so the FP64 path is mma.sync by necessity.
#include "candidates.cuh"

#define TM 64
#define NC 128
#define KC 16
#define SROW 20   // padded row stride (doubles): 4 rows x 8 words tile the 32 banks per half-warp
#define TCH 128   // training points staged per phase-A chunk

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ double kstar_of_r2(double r2, double sf2, int kernel) {
  if (kernel == OMBO_KERNEL_MATERN52) {
    double r = sqrt(r2);
    const double s5 = 2.23606797749978969641;
    return sf2 * (1.0 + s5 * r + (5.0 / 3.0) * r2) * exp(-s5 * r);
  }
  return sf2 * exp(-0.5 * r2);
}

extern __shared__ __align__(16) unsigned char smem_raw[];

__global__ void __launch_bounds__(256, 2)
k_posterior_fp64(GpDev gp, PoolDev pool, long long m, double *__restrict__ scratch_all,
                 double *__restrict__ mu_out, double *__restrict__ var_out, int want_var) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int d = gp.d, n = gp.n, np = gp.n_pad;
  double *scratch = scratch_all + (size_t)blockIdx.x * TM * np;

  // smem carve-up (phase A and phase B overlay the same bytes)
  double *As = (double *)smem_raw;                 // [2][TM][SROW]
  double *Bs = As + 2 * TM * SROW;                 // [2][NC][SROW]
  double *xc = (double *)smem_raw;                 // phase A: [TM][d]
  double *xt = xc + TM * d;                        // phase A: [d][TCH]
  __shared__ double red_mu[TM];
  __shared__ double red_ss[4][TM];   // one slot per N-warp: fixed-order sum, bit-reproducible

  const long long num_tiles = (m + TM - 1) / TM;
  for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const long long c0 = tile * TM;
    // ---------------- phase A: K* tile + mean ------------------------------------------
    __syncthreads();
    for (int e = tid; e < TM * d; e += 256) {
      int c = e / d, j = e % d;
      long long cg = c0 + c;
      xc[e] = (cg < m) ? ombo_pool_coord(pool, cg, j) / gp.ell[j] : 0.0;
    }
    if (tid < TM) red_mu[tid] = 0.0;
    for (int i0 = 0; i0 < np; i0 += TCH) {
      __syncthreads();
      for (int e = tid; e < d * TCH; e += 256) {
        int j = e / TCH, i = e % TCH;
        xt[e] = gp.xs[(size_t)j * np + i0 + i];
      }
      __syncthreads();
      double al[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) al[q] = gp.alpha[i0 + lane + 32 * q];
#pragma unroll 1
      for (int t = 0; t < 8; ++t) {
        const int c = warp + 8 * t;
        double r2[4] = {0.0, 0.0, 0.0, 0.0};
        double mu_part = 0.0;
        for (int j = 0; j < d; ++j) {
          double x = xc[c * d + j];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            double df = x - xt[j * TCH + lane + 32 * q];
            r2[q] += df * df;
          }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          int i = i0 + lane + 32 * q;
          double k = (i < n) ? kstar_of_r2(r2[q], gp.sigma_f2, gp.kernel) : 0.0;
          if (want_var) scratch[(size_t)c * np + i] = k;
          mu_part += k * al[q];
        }
        for (int o = 16; o; o >>= 1) mu_part += __shfl_xor_sync(0xffffffffu, mu_part, o);
        if (lane == 0) red_mu[c] += mu_part;     // candidate c belongs to this warp only
      }
    }
    __threadfence();
    __syncthreads();

    // ---------------- phase B: V = K* . Linv^T, fused sum of squares ---------------------
    // warp w runs on scheduler w & 3.  The second row of warps takes its column groups in reverse, so that every
    // scheduler hosts column groups {q, 3 - q}: in the diagonal steps below group q multiplies q + 1 quarters of the
    // K range, and with the plain map all of the long ones sat on scheduler 3
    const int wm = warp >> 2, wn = wm == 0 ? (warp & 3) : 3 - (warp & 3);
    const int g8 = lane >> 2, t4 = lane & 3;
    double ss[4] = {0.0, 0.0, 0.0, 0.0};
    for (int j0 = 0; j0 < (want_var ? np : 0); j0 += NC) {
      if (j0 >= n) break;                       // remaining rows of Linv are padding (zero)
      double acc[4][4][2];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
      const int steps = (j0 + NC) / KC;
      auto issue = [&](int ks, int buf) {
        const int k0 = ks * KC;
        // A: 64 rows x 128 B = 512 x 16 B
        for (int e = tid; e < TM * 8; e += 256) {
          int row = e >> 3, seg = e & 7;
          cp_async16(As + ((size_t)buf * TM + row) * SROW + seg * 2, scratch + (size_t)row * np + k0 + seg * 2);
        }
        for (int e = tid; e < NC * 8; e += 256) {
          int row = e >> 3, seg = e & 7;
          cp_async16(Bs + ((size_t)buf * NC + row) * SROW + seg * 2,
                     gp.Linv + (size_t)(j0 + row) * np + k0 + seg * 2);
        }
        cp_async_commit();
      };
      auto mma_step = [&](int buf) {
        const double *Ab = As + (size_t)buf * TM * SROW + (size_t)(32 * wm + g8) * SROW + t4;
        const double *Bb = Bs + (size_t)buf * NC * SROW + (size_t)(32 * wn + g8) * SROW + t4;
#pragma unroll
        for (int kk = 0; kk < KC / 4; ++kk) {
          double a[4], b[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) { a[u] = Ab[u * 8 * SROW + kk * 4]; b[u] = Bb[u * 8 * SROW + kk * 4]; }
#pragma unroll
          for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int v = 0; v < 4; ++v) dmma884(acc[u][v][0], acc[u][v][1], a[u], b[v]);
        }
      };
      // k < j0: every warp multiplies.  The last 128 / KC steps run through the diagonal block of L^-1, where the warp
      // holding columns [j0 + 32 wn, + 32) only needs k < j0 + 32 wn + 32 (the rest of its B rows is zero): a separate
      // loop, so that the main loop stays free of the per-warp branch
      const int steps_all = j0 / KC, my_steps = (j0 + 32 * wn + 32) / KC;
      issue(0, 0);
      int ks = 0;
      for (; ks < steps_all; ++ks) {
        const int buf = ks & 1;
        issue(ks + 1, buf ^ 1);                   // (steps_all < steps: there is always a next step here)
        cp_async_wait<1>();
        __syncthreads();
        mma_step(buf);
        __syncthreads();
      }
      for (; ks < steps; ++ks) {
        const int buf = ks & 1;
        if (ks + 1 < steps) { issue(ks + 1, buf ^ 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
        __syncthreads();
        if (ks < my_steps) mma_step(buf);
        __syncthreads();
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) ss[u] += acc[u][v][0] * acc[u][v][0] + acc[u][v][1] * acc[u][v][1];
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      double s = ss[u];
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      if (t4 == 0) red_ss[wn][32 * wm + 8 * u + g8] = s;
    }
    __syncthreads();
    if (tid < TM) {
      long long cg = c0 + tid;
      if (cg < m) {
        double v = gp.sigma_f2 - (((red_ss[0][tid] + red_ss[1][tid]) + red_ss[2][tid]) + red_ss[3][tid]);
        v = fmax(v, gp.var_floor) + gp.sigma_n2;
        mu_out[cg] = red_mu[tid];
        var_out[cg] = want_var ? v : nan("");
      }
    }
  }
}

int ombo_posterior_fp64(ombo_ctx *ctx, const GpDev &gp, const PoolDev &pool, long long m, double *mu,
                        double *var, bool want_var, cudaStream_t s) {
  if (m <= 0) return OMBO_OK;
  const size_t smem_b = (size_t)(2 * TM * SROW + 2 * NC * SROW) * 8;
  const size_t smem_a = (size_t)(TM * gp.d + gp.d * TCH) * 8;
  const size_t smem = smem_a > smem_b ? smem_a : smem_b;
  OMBO_CUDA(cudaFuncSetAttribute(k_posterior_fp64, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));   // per (function, device)
  long long tiles = (m + TM - 1) / TM;
  int grid = (int)(tiles < 2LL * ctx->num_sms ? tiles : 2LL * ctx->num_sms);
  size_t want = (size_t)grid * TM * gp.n_pad * 8;
  int rc = ombo_ws_reserve(&ctx->ws_scratch, &ctx->ws_scratch_bytes, want);
  if (rc) return rc;
  {
    ProfScope prof(ctx, s);
    k_posterior_fp64<<<grid, 256, smem, s>>>(gp, pool, m, (double *)ctx->ws_scratch, mu, var, want_var ? 1 : 0);
  }
  ctx->launches += 1;
  OMBO_CUDA(cudaGetLastError());
  return OMBO_OK;
}
