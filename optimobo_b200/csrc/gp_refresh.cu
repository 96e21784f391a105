// K3 -- GP refresh: K = k(X,X) + (sigma_n2 + jitter) I, L = chol(K), L^-1, alpha = K^-1 y,
// plus the fast-path operand planes.  Replaces the state GPy builds inside
// GPRegression(...) after .optimize() (reference call sites: optimisers.py:226-231 etc.;
// GPy exact_gaussian_inference.py / pdinv / dpotrs).
//
// Blocked right-looking FP64 Cholesky with 64x64 blocks: the diagonal block is factorised
// AND inverted by one CTA in shared memory; the panel solve and the trailing update are
// 64x64x64 GEMM tiles spread over the whole GPU.  The triangular inverse is column-block
// parallel (16 columns per CTA, n_pad/16 CTAs), each CTA sweeping down its columns with
// the pre-inverted diagonal blocks, so no step is a scalar substitution.
#include "common.cuh"

#define NB OMBO_NB
#define LDS (NB + 1)

// ------------------------------------------------------------------------------------------
// per-dimension mean of the training inputs (fixed-order tree: bit-reproducible)
__global__ void __launch_bounds__(256) k_center(const double *__restrict__ X, int n, int d, double *__restrict__ center) {
  __shared__ double part[256];
  const int j = blockIdx.x;
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) s += X[(size_t)i * d + j];
  part[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o; o >>= 1) {
    if (threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) center[j] = part[0] / (double)n;
}

__global__ void k_prepare(const double *__restrict__ X, const double *__restrict__ y, int n, int d,
                          int n_pad, const double *__restrict__ ell_dev, const double *__restrict__ center,
                          double *__restrict__ xs, float *__restrict__ xs32, float *__restrict__ b2,
                          double *__restrict__ ypad) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pad) return;
  for (int j = 0; j < d; ++j) {
    double v = (i < n) ? X[(size_t)i * d + j] / ell_dev[j] : 0.0;
    xs[(size_t)j * n_pad + i] = v;
  }
  // fast path: CENTRED scaled inputs in FP32 and their squared norms.  The fast K1 uses
  // r^2 = |a|^2 + |b|^2 - 2 a.b (one FMA per pair and dimension); centring keeps |a|, |b| small so the
  // cancellation error stays ~1e-7 like the direct-difference form.
  float nb = 0.f;
  for (int j = 0; j < 32; ++j) {
    float v = (i < n && j < d) ? (float)((X[(size_t)i * d + j] - center[j]) / ell_dev[j]) : 0.f;
    xs32[(size_t)j * n_pad + i] = v;
    nb = fmaf(v, v, nb);
  }
  b2[i] = nb;
  ypad[i] = (i < n) ? y[i] : 0.0;
}

__device__ __forceinline__ double kernel_of_r2(double r2, double sf2, int kernel) {
  if (kernel == OMBO_KERNEL_MATERN52) {
    double r = sqrt(r2);
    const double s5 = 2.23606797749978969641;
    return sf2 * (1.0 + s5 * r + (5.0 / 3.0) * r2) * exp(-s5 * r);
  }
  return sf2 * exp(-0.5 * r2);
}

__global__ void k_build_K(const double *__restrict__ xs, int n, int d, int n_pad, double sf2,
                          double diag_add, int kernel, double *__restrict__ K) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  int i = blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= n_pad || j >= n_pad) return;
  double v;
  if (i >= n || j >= n) {
    v = (i == j) ? 1.0 : 0.0;
  } else {
    double r2 = 0.0;
    for (int t = 0; t < d; ++t) {
      double df = xs[(size_t)t * n_pad + i] - xs[(size_t)t * n_pad + j];
      r2 += df * df;
    }
    v = kernel_of_r2(r2, sf2, kernel);
    if (i == j) v += diag_add;
  }
  K[(size_t)i * n_pad + j] = v;
}

// k_potrf_diag (Cholesky + inverse of one 64x64 diagonal block) lives in gp_potrf_diag.cu: its fully unrolled
// 64-step register kernels take the NVVM optimiser five minutes at -O3 -- kept apart so that edits here build fast
void ombo_launch_potrf_diag(double *A, int ld, int kb, int n, double *dinv, int *status, cudaStream_t s);
static_assert(OMBO_NB == 64, "gp_potrf_diag.cu hard-codes the 64x64 diagonal block");

// C(64x64) (+)= sign * As(64x64) * Bs(64x64)^T with both operands k-contiguous in smem
__device__ __forceinline__ void gemm64_abt(double (*As)[LDS], double (*Bs)[LDS], double acc[4][4],
                                           int tr, int tc) {
#pragma unroll 4
  for (int t = 0; t < NB; ++t) {
    double av[4], bv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) { av[u] = As[tr + 16 * u][t]; bv[u] = Bs[tc + 16 * u][t]; }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int v = 0; v < 4; ++v) acc[u][v] += av[u] * bv[v];
  }
}

// L_ik = A_ik * Dinv_k^T   (i > k)
__global__ void __launch_bounds__(256) k_trsm_panel(double *__restrict__ A, int ld, int kb,
                                                    const double *__restrict__ dinv) {
  extern __shared__ __align__(16) double dyn_smem[];
  double (*As)[LDS] = (double (*)[LDS])dyn_smem;
  double (*Bs)[LDS] = (double (*)[LDS])(dyn_smem + NB * LDS);
  const int tid = threadIdx.x;
  const int ib = kb + 1 + blockIdx.x;
  double *blk = A + ((size_t)ib * NB) * ld + (size_t)kb * NB;
  const double *dv = dinv + (size_t)kb * NB * NB;
  for (int e = tid; e < NB * NB; e += 256) {
    As[e / NB][e % NB] = blk[(size_t)(e / NB) * ld + (e % NB)];
    Bs[e / NB][e % NB] = dv[e];
  }
  __syncthreads();
  const int tr = tid / 16, tc = tid % 16;
  double acc[4][4] = {};
  gemm64_abt(As, Bs, acc, tr, tc);
  __syncthreads();
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v) blk[(size_t)(tr + 16 * u) * ld + (tc + 16 * v)] = acc[u][v];
}

// A_ij -= L_ik * L_jk^T   (k < j <= i)
__global__ void __launch_bounds__(256) k_syrk_update(double *__restrict__ A, int ld, int kb) {
  const int ib = kb + 1 + blockIdx.y;
  const int jb = kb + 1 + blockIdx.x;
  if (jb > ib) return;
  extern __shared__ __align__(16) double dyn_smem[];
  double (*As)[LDS] = (double (*)[LDS])dyn_smem;
  double (*Bs)[LDS] = (double (*)[LDS])(dyn_smem + NB * LDS);
  const int tid = threadIdx.x;
  const double *Li = A + ((size_t)ib * NB) * ld + (size_t)kb * NB;
  const double *Lj = A + ((size_t)jb * NB) * ld + (size_t)kb * NB;
  for (int e = tid; e < NB * NB; e += 256) {
    As[e / NB][e % NB] = Li[(size_t)(e / NB) * ld + (e % NB)];
    Bs[e / NB][e % NB] = Lj[(size_t)(e / NB) * ld + (e % NB)];
  }
  __syncthreads();
  const int tr = tid / 16, tc = tid % 16;
  double acc[4][4] = {};
  gemm64_abt(As, Bs, acc, tr, tc);
  double *C = A + ((size_t)ib * NB) * ld + (size_t)jb * NB;
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v) C[(size_t)(tr + 16 * u) * ld + (tc + 16 * v)] -= acc[u][v];
}

// ------------------------------------------------------------------------------------------
// triangular inverse, TC columns per CTA.  X = L^-1, block row by block row:
//   T = I[ib rows, these columns] - sum_{jb < ib} L[ib][jb] X[jb],   X[ib] = Dinv[ib] T
// The CTA walks the tile products (ib, jb = kb .. ib-1) and the diagonal product (ib, jb = ib) as ONE sequence of
// steps and double-buffers their operands with cp.async: the 32 KB tile of L (or of Dinv) and the 64 x 16 tile of X for
// step s + 1 arrive while step s multiplies (round 1 loaded them between two barriers, and the first column block
// runs 120 + 16 such steps back to back).  Every sum keeps its order (jb ascending, then t ascending), so the result is
// bit-identical to the unpipelined kernel.
#define TC 8       // columns per CTA: n_pad / 8 CTAs (128 at n = 1024, one per SM); 16 halves the CTAs and doubles the chain
#define TCV (TC / 4) // columns per thread
__device__ __forceinline__ void cp_async8(void *smem, const void *gmem) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(gmem));
}
extern __shared__ __align__(16) double trtri_smem[];
__global__ void __launch_bounds__(256) k_trtri_cols(const double *__restrict__ L, int ld, int nb,
                                                    const double *__restrict__ dinv,
                                                    double *__restrict__ Linv) {
  double (*Lb)[NB][LDS] = (double (*)[NB][LDS])trtri_smem;                              // [2] L tile / Dinv tile
  double (*Xb)[NB][TC + 1] = (double (*)[NB][TC + 1])(trtri_smem + 2 * NB * LDS);       // [2] X rows of block jb / T
  const int tid = threadIdx.x;
  const int c0 = blockIdx.x * TC;
  const int kb = c0 / NB;
  const int r = tid % NB;          // output row inside the block row
  const int cg = (tid / NB) * TCV; // first of this thread's output columns
  auto issue = [&](int ib, int jb, int buf, bool with_x) {
    const double *src = jb < ib ? L + ((size_t)ib * NB) * ld + (size_t)jb * NB : dinv + (size_t)ib * NB * NB;
    const int stride = jb < ib ? ld : NB;
    for (int e = tid; e < NB * NB; e += 256) cp_async8(&Lb[buf][e / NB][e % NB], src + (size_t)(e / NB) * stride + (e % NB));
    if (with_x)
      for (int e = tid; e < NB * TC; e += 256)
        cp_async8(&Xb[buf][e / TC][e % TC], Linv + ((size_t)jb * NB + e / TC) * ld + c0 + (e % TC));
    asm volatile("cp.async.commit_group;\n" ::);
  };
  int ib = kb, jb = kb, buf = 0;
  issue(ib, jb, 0, false);
  double T[TCV];
#pragma unroll
  for (int v = 0; v < TCV; ++v) T[v] = (ib * NB + r == c0 + cg + v) ? 1.0 : 0.0;
  while (ib < nb) {
    const int nib = jb < ib ? ib : ib + 1, njb = jb < ib ? jb + 1 : kb;
    const bool has_next = nib < nb;
    // X[njb] is final unless THIS step is the diagonal product that writes it (only at the very first step)
    const bool next_x = has_next && njb < nib && !(jb == ib && njb == ib);
    if (has_next) {
      issue(nib, njb, buf ^ 1, next_x);
      asm volatile("cp.async.wait_group 1;\n" ::);
    } else {
      asm volatile("cp.async.wait_group 0;\n" ::);
    }
    __syncthreads();
    if (jb < ib) {
      if (ib == kb + 1 && jb == kb) {        // the one X tile that could not be prefetched
        for (int e = tid; e < NB * TC; e += 256)
          Xb[buf][e / TC][e % TC] = Linv[((size_t)jb * NB + e / TC) * ld + c0 + (e % TC)];
        __syncthreads();
      }
#pragma unroll 8
      for (int t = 0; t < NB; ++t) {
        double l = Lb[buf][r][t];
#pragma unroll
        for (int v = 0; v < TCV; ++v) T[v] -= l * Xb[buf][t][cg + v];
      }
    } else {
      // X_ib = Dinv_ib * T
#pragma unroll
      for (int v = 0; v < TCV; ++v) Xb[buf][r][cg + v] = T[v];
      __syncthreads();
      double o[TCV];
#pragma unroll
      for (int v = 0; v < TCV; ++v) o[v] = 0.0;
#pragma unroll 8
      for (int t = 0; t < NB; ++t) {
        double l = Lb[buf][r][t];
#pragma unroll
        for (int v = 0; v < TCV; ++v) o[v] += l * Xb[buf][t][cg + v];
      }
#pragma unroll
      for (int v = 0; v < TCV; ++v) Linv[((size_t)ib * NB + r) * ld + c0 + cg + v] = o[v];
      __threadfence_block();
#pragma unroll
      for (int v = 0; v < TCV; ++v) T[v] = ((ib + 1) * NB + r == c0 + cg + v) ? 1.0 : 0.0;
    }
    __syncthreads();
    ib = nib; jb = njb; buf ^= 1;
  }
}

// t[j] = sum_i M[j][i] * v[i], i <= j   (one warp per row)
__global__ void k_trmv_lower(const double *__restrict__ M, int ld, int n_pad, const double *__restrict__ v,
                             double *__restrict__ out) {
  int row = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  int lane = threadIdx.x % 32;
  if (row >= n_pad) return;
  double s = 0.0;
  for (int i = lane; i <= row; i += 32) s += M[(size_t)row * ld + i] * v[i];
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[row] = s;
}

// out[i] = sum_{j >= i} M[j][i] * t[j]    (block (32 cols x 8 row groups))
__global__ void k_trmv_lower_t(const double *__restrict__ M, int ld, int n_pad, const double *__restrict__ t,
                               double *__restrict__ out) {
  __shared__ double part[8][33];
  int col = blockIdx.x * 32 + threadIdx.x;
  double s = 0.0;
  if (col < n_pad)
    for (int j = col + threadIdx.y; j < n_pad; j += 8) s += M[(size_t)j * ld + col] * t[j];
  part[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && col < n_pad) {
    double a = 0.0;
    for (int g = 0; g < 8; ++g) a += part[g][threadIdx.x];
    out[col] = a;
  }
}

// max |L^-1| over the valid lower triangle and min / max of diag(L) (bit patterns of non-negative doubles order
// like the values): out[0] = max |L^-1|, out[1] = max L_ii, out[2] = min L_ii (initialised to +inf bits by the host)
__global__ void k_absmax_lower(const double *__restrict__ Linv, const double *__restrict__ L, int n, int n_pad,
                               unsigned long long *__restrict__ out) {
  size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  double v = 0.0;
  if (e < (size_t)n_pad * n_pad) {
    const int j = (int)(e / n_pad), i = (int)(e % n_pad);
    if (j < n && i <= j) v = fabs(Linv[e]);
    if (j < n && i == j) {
      const unsigned long long dg = (unsigned long long)__double_as_longlong(fabs(L[e]));
      atomicMax(out + 1, dg);
      atomicMin(out + 2, dg);
    }
  }
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0 && v > 0.0) atomicMax(out, (unsigned long long)__double_as_longlong(v));
}

// zero the padding of L^-1 and emit the fast-path operands B = s * sigma_f2 * L^-1 as planes.
// The format is chosen here, on the device, from the conditioning proxy kappa = (max L_ii / min L_ii)^2:
//   f8c  (kappa <= f8c_kappa, caller allows it): hi plane fp16 with s = the power of two that brings max |B| into
//        [8192, 16384); the lo field holds two e4m3 planes, e4m3(2^-12 B) and e4m3(B - fp16(B)) -- the operands of
//        the two correction products of posterior_fast8.cu (2.0 tensor units per MAC);
//   bf16 (kappa <= 100 otherwise): bf16 hi + lo planes, s = 1 (16-bit x3 split, 3.0 units);
//   fp16 (kappa  > 100, caller allows it): fp16 hi + lo planes with the same power-of-two scale; the generators then
//        also use direct-difference distances.
// bscale[1] = s, [2] = 1 / s^2, [3] = format (0 bf16, 1 fp16, 2 f8c).
#include <cuda_fp8.h>
__device__ __forceinline__ int plane_format(const double *bscale, int allow, double f8c_kappa) {
  const double ratio = bscale[2] > 0.0 ? bscale[1] / bscale[2] : 1.0;       // max L_ii / min L_ii
  const double kappa = ratio * ratio;
  if ((allow & OMBO_GP_F8C_PLANES) && kappa <= f8c_kappa) return 2;
  if ((allow & OMBO_GP_FP16_PLANES) && kappa > 100.0) return 1;
  return 0;
}
__global__ void k_finalize(double *__restrict__ Linv, int n, int n_pad, double sf2, int allow, double f8c_kappa,
                           unsigned short *__restrict__ bhi, unsigned short *__restrict__ blo, double *__restrict__ bscale,
                           const double *__restrict__ alpha, float *__restrict__ alpha32) {
  size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t total = (size_t)n_pad * n_pad;
  if (e >= total) return;
  const int fmt = plane_format(bscale, allow, f8c_kappa);
  const double mx = bscale[0] * sf2;
  const double sc = (fmt != 0 && mx > 0.0) ? exp2(floor(log2(16384.0 / mx))) : 1.0;
  int j = (int)(e / n_pad), i = (int)(e % n_pad);
  double v = Linv[e];
  if (j >= n || i >= n || i > j) { v = 0.0; Linv[e] = 0.0; }
  const double t = v * sf2 * sc;
  if (fmt == 2) {
    const __half h = __float2half_rn((float)t);
    bhi[e] = __half_as_ushort(h);
    unsigned char *c1 = (unsigned char *)blo, *c2 = c1 + total;
    c1[e] = (unsigned char)__nv_cvt_float_to_fp8((float)(t * (1.0 / 4096.0)), __NV_SATFINITE, __NV_E4M3);
    c2[e] = (unsigned char)__nv_cvt_float_to_fp8((float)(t - (double)__half2float(h)), __NV_SATFINITE, __NV_E4M3);
  } else if (fmt == 1) {
    const __half h = __float2half_rn((float)t);
    const __half l = __float2half_rn((float)(t - (double)__half2float(h)));
    bhi[e] = __half_as_ushort(h);
    blo[e] = __half_as_ushort(l);
  } else {
    const __nv_bfloat16 h = __float2bfloat16_rn((float)t);
    const __nv_bfloat16 l = __float2bfloat16_rn((float)(t - (double)__bfloat162float(h)));
    bhi[e] = __bfloat16_as_ushort(h);
    blo[e] = __bfloat16_as_ushort(l);
  }
  if (e < (size_t)n_pad) alpha32[e] = (float)(alpha[e] * sf2);
}

// the per-GP constants the fast kernels read: [1] = s, [2] = 1 / s^2, [3] = format (after every thread of
// k_finalize has read the raw extrema: a separate launch)
__global__ void k_publish_scale(double *__restrict__ bscale, double sf2, int allow, double f8c_kappa, int *__restrict__ status) {
  const int fmt = plane_format(bscale, allow, f8c_kappa);
  const double mx = bscale[0] * sf2;
  const double sc = (fmt != 0 && mx > 0.0) ? exp2(floor(log2(16384.0 / mx))) : 1.0;
  bscale[1] = sc; bscale[2] = 1.0 / (sc * sc); bscale[3] = (double)fmt;
  status[1] = fmt;          // read back by the host with the PD flag: 1 -> OMBO_GP_FP16_PLANES, 2 -> OMBO_GP_F8C_PLANES
}

// ------------------------------------------------------------------------------------------
// In-place blocked Cholesky of the lower triangle of an (np x np, np % 64 == 0) FP64 matrix; dinv receives the
// inverted 64x64 diagonal blocks, *status the first non-positive pivot (+1) or 0.  Used by the refresh below and
// by posterior_joint.cu (posterior covariance of a candidate set).  The status word is left on the device.
int ombo_potrf_lower_impl(ombo_ctx *ctx, double *L, int np, int n, double *dinv, int *status, cudaStream_t s) {
  const int nb = np / NB;
  const size_t sm2 = (size_t)2 * NB * LDS * sizeof(double);
  // per (function, device): one process may drive several devices, so no process-wide "already set" flag
  OMBO_CUDA(cudaFuncSetAttribute(k_trsm_panel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2));
  OMBO_CUDA(cudaFuncSetAttribute(k_syrk_update, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2));
  for (int kb = 0; kb < nb; ++kb) {
    ombo_launch_potrf_diag(L, np, kb, n, dinv, status, s);
    ctx->launches += 1;
    int rem = nb - kb - 1;
    if (rem > 0) {
      k_trsm_panel<<<rem, 256, sm2, s>>>(L, np, kb, dinv);
      k_syrk_update<<<dim3(rem, rem), 256, sm2, s>>>(L, np, kb);
      ctx->launches += 2;
    }
  }
  OMBO_CUDA(cudaGetLastError());
  return OMBO_OK;
}

int ombo_refresh_impl(ombo_ctx *ctx, const ombo_gp_spec *sp, void *state, cudaStream_t s) {
  const GpLayout lay = gp_layout(sp->n, sp->d);
  const int n = sp->n, d = sp->d, np = lay.n_pad, nb = np / NB;
  char *b = (char *)state;
  double *L = (double *)(b + lay.off_L), *Linv = (double *)(b + lay.off_Linv);
  double *alpha = (double *)(b + lay.off_alpha), *xs = (double *)(b + lay.off_xs);
  int *status = (int *)(b + lay.off_status);
  double *dinv = (double *)(b + lay.off_dinv), *tmp = (double *)(b + lay.off_tmp);
  double *ell_dev = (double *)(b + lay.off_inv_ell);
  unsigned short *bhi = (unsigned short *)(b + lay.off_bhi), *blo = (unsigned short *)(b + lay.off_blo);
  double *bscale = (double *)(b + lay.off_bscale);
  float *xs32 = (float *)(b + lay.off_xs32), *alpha32 = (float *)(b + lay.off_alpha32);
  float *b2 = (float *)(b + lay.off_b2);
  double *center = (double *)(b + lay.off_center);

  OMBO_CUDA(cudaMemcpyAsync(ell_dev, sp->ell, sizeof(double) * d, cudaMemcpyHostToDevice, s));
  OMBO_CUDA(cudaMemsetAsync(status, 0, 16, s));
  OMBO_CUDA(cudaMemsetAsync(Linv, 0, (size_t)np * np * 8, s));
  k_center<<<d, 256, 0, s>>>(sp->X, n, d, center);
  k_prepare<<<(np + 127) / 128, 128, 0, s>>>(sp->X, sp->y, n, d, np, ell_dev, center, xs, xs32, b2, alpha /*ypad*/);
  ctx->launches += 1;
  dim3 tb(16, 16), gb(np / 16, np / 16);
  k_build_K<<<gb, tb, 0, s>>>(xs, n, d, np, sp->sigma_f2, sp->sigma_n2 + sp->jitter, sp->kernel, L);
  ctx->launches += 2;
  int rc = ombo_potrf_lower_impl(ctx, L, np, n, dinv, status, s);
  if (rc) return rc;
  const size_t sm_tri = (size_t)(2 * NB * LDS + 2 * NB * (TC + 1)) * sizeof(double);
  OMBO_CUDA(cudaFuncSetAttribute(k_trtri_cols, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_tri));
  k_trtri_cols<<<np / TC, 256, sm_tri, s>>>(L, np, nb, dinv, Linv);
  // alpha = Linv^T (Linv y); ypad currently lives in alpha
  k_trmv_lower<<<(np + 7) / 8, 256, 0, s>>>(Linv, np, np, alpha, tmp);
  k_trmv_lower_t<<<(np + 31) / 32, dim3(32, 8), 0, s>>>(Linv, np, np, tmp, alpha);
  size_t total = (size_t)np * np;
  {
    static const unsigned long long init[4] = {0ull, 0ull, 0x7ff0000000000000ull, 0ull};      // max, max, min (+inf), -
    OMBO_CUDA(cudaMemcpyAsync(bscale, init, 32, cudaMemcpyHostToDevice, s));
  }
  // fp16 planes only for callers that announce (spec->reserved bit 1) that they will pass the reported format on
  int allow = sp->reserved & (OMBO_GP_FP16_PLANES | OMBO_GP_F8C_PLANES);
  if (d > 12 || ctx->knobs.no_f8c) allow &= ~OMBO_GP_F8C_PLANES;       // the f8c kernel is instantiated for d <= 12
  const double f8c_kappa = ctx->knobs.f8c_kappa;
  k_absmax_lower<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(Linv, L, n, np, (unsigned long long *)bscale);
  k_finalize<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(Linv, n, np, sp->sigma_f2, allow, f8c_kappa, bhi, blo, bscale, alpha, alpha32);
  k_publish_scale<<<1, 1, 0, s>>>(bscale, sp->sigma_f2, allow, f8c_kappa, status);
  ctx->launches += 6;
  OMBO_CUDA(cudaGetLastError());
  int hstatus[4] = {0, 0, 0, 0};
  OMBO_CUDA(cudaMemcpyAsync(hstatus, status, 16, cudaMemcpyDeviceToHost, s));
  OMBO_CUDA(cudaStreamSynchronize(s));
  if (hstatus[0] != 0) {
    ombo_set_error("gp_refresh: matrix not positive definite at row %d (raise jitter and retry)", hstatus[0] - 1);
    return OMBO_ERR_NOT_PD;
  }
  return OMBO_OK;
}
