// K4 + K5 -- per-candidate acquisition value from the GP posteriors, and the arg-max.
// One thread per candidate, FP64 throughout (the reference is float64 numpy/scipy); the sorted
// Pareto-front stripes / cells / cached samples are staged in shared memory once per block.
//
// Formulas restate (file:line into the reference tree, optimobo/):
//   EHVI2D            util_functions.py:136-167 (EHVI) + :81-128 (EHVI_2D_aux) + :130-133 (psi_cal)
//   EHVI3D            util_functions.py:170-214
//   EXPECTED_DECOMP   util_functions.py:285-327 with the 12 scalarisations of scalarisations.py
//   EI                optimisers.py:325-344, parego.py:126-145, keep.py:118-137, cparego.py:450-469
//   CONSTRAINED_EI    cparego.py:471-496
//   PARETO_EI         keep.py:142-150
//   HV_POI            emo.py:176-228
// Arg-max replaces the inner optimiser call sites (differential_evolution at optimisers.py:87,
// :118,:366; cparego.py:94,:544; emo.py:240; EA loops parego.py:242-270, keep.py:260-292):
// warp shuffle -> block (smem) -> per-block partial -> one final block; ties resolve to the
// lowest global index (np.argmax), NaN is treated as -inf.
#include "common.cuh"
#include "acq_math.cuh"

#define ACQ_THREADS 256


// ---- scalarisations (scalarisations.py) --------------------------------------------------
__device__ double scalarise(const ombo_acq &a, const double *F) {
  const int k = a.n_obj;
  double Fp[OMBO_MAX_OBJ];
  for (int i = 0; i < k; ++i) Fp[i] = (F[i] - a.ideal[i]) / (a.maxp[i] - a.ideal[i]);
  const double *w = a.weights;
  switch (a.scalarisation) {
    case OMBO_SC_WEIGHTED_SUM: {            // :43-50
      double s = 0.0;
      for (int i = 0; i < k; ++i) s += Fp[i] * w[i];
      return s;
    }
    case OMBO_SC_TCHEBICHEFF: {             // :66-72
      double mx = w[0] * Fp[0];
      for (int i = 1; i < k; ++i) mx = fmax(mx, w[i] * Fp[i]);
      return mx;
    }
    case OMBO_SC_AUG_TCHEBICHEFF: {         // :92-109
      double mx = fabs(Fp[0]) * w[0], sum = fabs(Fp[0]);
      for (int i = 1; i < k; ++i) { mx = fmax(mx, fabs(Fp[i]) * w[i]); sum += fabs(Fp[i]); }
      return mx + a.sc_params[0] * sum;
    }
    case OMBO_SC_MOD_TCHEBICHEFF: {         // :130-149
      double sum = 0.0;
      for (int i = 0; i < k; ++i) sum += fabs(Fp[i]);
      double right = a.sc_params[0] * sum;
      double mx = (fabs(Fp[0]) + right) * w[0];
      for (int i = 1; i < k; ++i) mx = fmax(mx, (fabs(Fp[i]) + right) * w[i]);
      return mx;
    }
    case OMBO_SC_EXP_WEIGHTED: {            // :165-173
      double s = 0.0, p = a.sc_params[0];
      for (int i = 0; i < k; ++i) s += exp(p * w[i] - 1.0) * exp(p * Fp[i]);
      return s;
    }
    case OMBO_SC_WEIGHTED_NORM: {           // :190-197
      double s = 0.0, p = a.sc_params[0];
      for (int i = 0; i < k; ++i) s += pow(fabs(Fp[i]), p) * w[i];
      return pow(s, 1.0 / p);
    }
    case OMBO_SC_WEIGHTED_POWER: {          // :212-219
      double s = 0.0, p = a.sc_params[0];
      for (int i = 0; i < k; ++i) s += pow(Fp[i], p) * w[i];
      return s;
    }
    case OMBO_SC_WEIGHTED_PRODUCT: {        // :230-238
      double s = 1.0;
      for (int i = 0; i < k; ++i) s *= pow(Fp[i] + 100000.0, w[i]);
      return s;
    }
    case OMBO_SC_PBI:
    case OMBO_SC_IPBI:
    case OMBO_SC_QPBI: {                    // :257-273, :291-310, :325-351
      double nw = 0.0;
      for (int i = 0; i < k; ++i) nw += w[i] * w[i];
      nw = sqrt(nw);
      double d1 = 0.0;
      for (int i = 0; i < k; ++i) d1 += Fp[i] * (w[i] / nw);
      double d2 = 0.0;
      for (int i = 0; i < k; ++i) { double e = Fp[i] - d1 * (w[i] / nw); d2 += e * e; }
      d2 = sqrt(d2);
      const double theta = a.sc_params[0];
      if (a.scalarisation == OMBO_SC_PBI) return d1 + theta * d2;
      if (a.scalarisation == OMBO_SC_IPBI) return theta * d2 - d1;
      double span = 0.0;
      for (int i = 0; i < k; ++i) span += a.maxp[i] - a.ideal[i];
      double d_star = a.sc_params[1] * ((1.0 / a.sc_params[2]) * (1.0 / (double)k) * span);
      return d1 + theta * d2 * (d2 / d_star);
    }
    case OMBO_SC_APD: {                     // :375-397
      double nf = 0.0;
      bool allz = true, wz = true;
      for (int i = 0; i < k; ++i) { nf += Fp[i] * Fp[i]; allz = allz && (Fp[i] == 0.0); wz = wz && (w[i] == 0.0); }
      nf = sqrt(nf);                        // norm taken before the 1e-5 patch (:384)
      double n1 = 0.0, n2 = 0.0, dot = 0.0;
      double f[OMBO_MAX_OBJ], ww[OMBO_MAX_OBJ];
      for (int i = 0; i < k; ++i) { f[i] = allz ? 1e-5 : Fp[i]; ww[i] = wz ? 1e-5 : w[i]; n1 += f[i] * f[i]; n2 += ww[i] * ww[i]; }
      n1 = sqrt(n1); n2 = sqrt(n2);
      for (int i = 0; i < k; ++i) dot += (f[i] / n1) * (ww[i] / n2);
      double th = acos(fmin(1.0, fmax(-1.0, dot)));
      return (1.0 + (double)k * (a.sc_params[0] / a.sc_params[1]) * (th / a.sc_params[2])) * nf;
    }
  }
  return nan("");
}

// ---- EI ----------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T ei_value(T mu, T var, T best, T eps) {
  T s = sqrt_t(var + eps);
  T g = (best - mu) / (s + (T)1e-10);
  return s * (g * Phi(g) + phi(g));
}

extern __shared__ __align__(16) double acq_smem[];

// T = double: the reference's arithmetic (FP64 precision mode, all goldens).  T = float: the fast precision mode
// (tolerance 1e-3) -- same formulas on the FP32 pipe with erfcf / MUFU exp; the expected decomposition always runs
// in double (ExponentialWeightedCriterion overflows FP32, WeightedProduct's 1e5 offset eats its digits).
template <typename T>
__global__ void __launch_bounds__(ACQ_THREADS)
k_acquire(ombo_acq a, int n_gp, const double *__restrict__ mu, const double *__restrict__ var,
          long long m, long long ld, long long index_base, double *__restrict__ out_acq,
          ombo_best *__restrict__ partials) {
  // stage the per-iteration constants
  T *cst = reinterpret_cast<T *>(acq_smem);
  int n_stage = 0;
  const double *src = nullptr;
  if (a.kind == OMBO_ACQ_EHVI2D) { n_stage = 2 * (a.n_pf + 2); src = a.stripes; }
  else if (a.kind == OMBO_ACQ_HV_POI) { n_stage = a.n_cells * 2 * a.n_obj; src = a.cells; }
  else if (a.kind == OMBO_ACQ_EHVI3D || a.kind == OMBO_ACQ_EXPECTED_DECOMP) { n_stage = a.n_samples * a.n_obj; src = a.cache; }
  for (int e = threadIdx.x; e < n_stage; e += ACQ_THREADS) cst[e] = (T)src[e];
  __syncthreads();

  const long long c = (long long)blockIdx.x * ACQ_THREADS + threadIdx.x;
  double val = -INFINITY;
  if (c < m) {
    T mus[OMBO_MAX_GP], vars[OMBO_MAX_GP];
    for (int g = 0; g < n_gp; ++g) { mus[g] = (T)mu[g * ld + c]; vars[g] = (T)var[g * ld + c]; }
    T r = (T)nan("");
    switch (a.kind) {
      case OMBO_ACQ_EHVI2D: {
        const int P = a.n_pf;
        r = ehvi2d_value<T>(mus[0], mus[1], vars[0], vars[1], cst, cst + (P + 2), P, a.semantics == OMBO_SEM_EXACT,
                            (T)a.cache_c00, (T)a.cache_c01);
      } break;
      case OMBO_ACQ_EHVI3D:
      case OMBO_ACQ_EXPECTED_DECOMP: {
        const int k = a.n_obj, S = a.n_samples;
        const bool exact = (a.semantics == OMBO_SEM_EXACT);
        T sd[OMBO_MAX_OBJ];
        for (int i = 0; i < k; ++i) sd[i] = sqrt_t(exact ? vars[i] : vars[0]);   // util_functions.py:233
        T total = 0;
        for (int j = 0; j < S; ++j) {
          T Y[OMBO_MAX_OBJ];
          for (int i = 0; i < k; ++i) Y[i] = cst[j * k + i] * sd[i] + mus[i];
          if (a.kind == OMBO_ACQ_EHVI3D) {
            bool inside = true;
            T hv = 1;
            for (int i = 0; i < k; ++i) { T df = (T)a.ref[i] - Y[i]; inside = inside && (df >= (T)0); hv = (i == 0) ? df : hv * df; }
            hv -= (T)a.best;    // Sminus = HV(PF), hoisted (util_functions.py:198-199)
            if (inside && hv > (T)0) total += hv;
          } else {
            double Yd[OMBO_MAX_OBJ];
            for (int i = 0; i < k; ++i) Yd[i] = (double)Y[i];
            double g = scalarise(a, Yd);
            total += (T)fmax(0.0, a.best - g);
            if (isnan(g)) total = (T)nan("");
          }
        }
        r = total / (T)S;
      } break;
      case OMBO_ACQ_EI:
        r = ei_value<T>(mus[0], vars[0], (T)a.best, (T)a.var_eps[0]);
        break;
      case OMBO_ACQ_CONSTRAINED_EI: {
        r = ei_value<T>(mus[0], vars[0], (T)a.best, (T)a.var_eps[0]);
        T pof = 1;
        for (int g = 1; g < n_gp; ++g) pof *= Phi(((T)0 - mus[g]) / sqrt_t(vars[g] + (T)a.var_eps[g]));
        r = r * pof;
      } break;
      case OMBO_ACQ_PARETO_EI:
        r = mus[0] * ei_value<T>(mus[1], vars[1], (T)a.best, (T)a.var_eps[1]);
        break;
      case OMBO_ACQ_HV_POI: {
        const int k = 2;     // emo.py:21 "Only works for 2D so far"
        T sd[2];
        for (int i = 0; i < k; ++i) sd[i] = sqrt_t(vars[i] + (T)a.var_eps[i]);
        T poi = 0, imp = 0;
        for (int cix = 0; cix < a.n_cells; ++cix) {
          const T *U = cst + (size_t)cix * 2 * a.n_obj, *Lo = U + a.n_obj;
          T p = 1, vol = 1;
          bool valid = true;
          for (int i = 0; i < k; ++i) {
            T pi = Phi((U[i] - mus[i]) / sd[i]) - Phi((Lo[i] - mus[i]) / sd[i]);
            p = (i == 0) ? pi : p * pi;
            valid = valid && (U[i] > mus[i]);
            T e = U[i] - fmax_t(Lo[i], mus[i]);
            vol = (i == 0) ? e : vol * e;
          }
          poi += p;
          if (valid) imp += vol;
        }
        r = poi * imp;
      } break;
      default:
        r = 0;
    }
    if (out_acq) out_acq[c] = (double)r;
    val = isnan(r) ? -INFINITY : (double)r;
  }
  // ---- K5: block arg-max ----
  BestPair b;
  b.v = val;
  b.i = (c < m) ? index_base + c : 0x7fffffffffffffffLL;
  b = warp_best(b);
  __shared__ double sv[ACQ_THREADS / 32];
  __shared__ long long si[ACQ_THREADS / 32];
  if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = b.v; si[threadIdx.x >> 5] = b.i; }
  __syncthreads();
  if (threadIdx.x < 32) {
    BestPair q;
    q.v = (threadIdx.x < ACQ_THREADS / 32) ? sv[threadIdx.x] : -INFINITY;
    q.i = (threadIdx.x < ACQ_THREADS / 32) ? si[threadIdx.x] : 0x7fffffffffffffffLL;
    q = warp_best(q);
    if (threadIdx.x == 0) { partials[blockIdx.x].value = q.v; partials[blockIdx.x].index = q.i; }
  }
}

// merges the per-block partials with the running best (chunks / previous passes)
__global__ void __launch_bounds__(1024) k_argmax_final(const ombo_best *__restrict__ partials, int n_partials,
                                                       ombo_best *__restrict__ best) {
  BestPair b;
  b.v = -INFINITY; b.i = 0x7fffffffffffffffLL;
  for (int e = threadIdx.x; e < n_partials; e += 1024) {
    double v = partials[e].value; long long i = partials[e].index;
    if (better(v, i, b.v, b.i)) { b.v = v; b.i = i; }
  }
  b = warp_best(b);
  __shared__ double sv[32];
  __shared__ long long si[32];
  if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = b.v; si[threadIdx.x >> 5] = b.i; }
  __syncthreads();
  if (threadIdx.x < 32) {
    BestPair q; q.v = sv[threadIdx.x]; q.i = si[threadIdx.x];
    q = warp_best(q);
    if (threadIdx.x == 0) {
      double cv = best->value; long long ci = best->index;
      if (ci < 0 || better(q.v, q.i, cv, ci)) { best->value = q.v; best->index = q.i; }
    }
  }
}

// the 12 scalarisation functions on explicit objective vectors (device twin of Scalarisation.__call__,
// scalarisations.py:20-27): out[r] = g(F[r, :], w)
__global__ void k_scalarise(ombo_acq a, const double *__restrict__ F, long long m, double *__restrict__ out) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= m) return;
  double f[OMBO_MAX_OBJ];
  for (int i = 0; i < a.n_obj; ++i) f[i] = F[r * a.n_obj + i];
  out[r] = scalarise(a, f);
}

int ombo_scalarise_impl(ombo_ctx *ctx, const ombo_acq *acq, const double *F, long long m, double *out, cudaStream_t s) {
  if (m <= 0) return OMBO_OK;
  k_scalarise<<<(unsigned)((m + 127) / 128), 128, 0, s>>>(*acq, F, m, out);
  ctx->launches += 1;
  OMBO_CUDA(cudaGetLastError());
  return OMBO_OK;
}

__global__ void k_best_init(ombo_best *best) { best->value = -INFINITY; best->index = -1; }

__global__ void k_pack_key(const ombo_best *__restrict__ best, long long *__restrict__ key) {
  float f = (float)best->value;
  unsigned bits = __float_as_uint(f);
  unsigned ord = (bits & 0x80000000u) ? ~bits : (bits | 0x80000000u);
  long long idx = best->index;
  unsigned low = (idx < 0 || idx > 0xFFFFFFFFll) ? 0u : (0xFFFFFFFFu - (unsigned)idx);
  long long hi = (long long)ord - 2147483648ll;            // signed, order-preserving
  *key = (hi << 32) | (long long)low;
}

int ombo_pack_key_impl(ombo_ctx *ctx, const ombo_best *best_dev, long long *key_dev, cudaStream_t s) {
  k_pack_key<<<1, 1, 0, s>>>(best_dev, key_dev);
  ctx->launches += 1;
  OMBO_CUDA(cudaGetLastError());
  return OMBO_OK;
}

int ombo_best_init(ombo_ctx *ctx, ombo_best *best_dev, cudaStream_t s) {
  k_best_init<<<1, 1, 0, s>>>(best_dev);
  ctx->launches += 1;
  OMBO_CUDA(cudaGetLastError());
  return OMBO_OK;
}

int ombo_argmax_merge(ombo_ctx *ctx, int n_partials, ombo_best *best_dev, cudaStream_t s) {
  k_argmax_final<<<1, 1024, 0, s>>>((const ombo_best *)ctx->ws_partial, n_partials, best_dev);
  ctx->launches += 1;
  OMBO_CUDA(cudaGetLastError());
  return OMBO_OK;
}

int ombo_acquire(ombo_ctx *ctx, const ombo_acq *acq, int n_gp, const double *mu, const double *var,
                 long long m, long long ld, long long index_base, double *out_acq, ombo_best *best_dev,
                 cudaStream_t s, bool fp32) {
  if (m <= 0) return OMBO_OK;
  int n_stage = 0;
  if (acq->kind == OMBO_ACQ_EHVI2D) n_stage = 2 * (acq->n_pf + 2);
  else if (acq->kind == OMBO_ACQ_HV_POI) n_stage = acq->n_cells * 2 * acq->n_obj;
  else if (acq->kind == OMBO_ACQ_EHVI3D || acq->kind == OMBO_ACQ_EXPECTED_DECOMP) n_stage = acq->n_samples * acq->n_obj;
  size_t smem = (size_t)n_stage * 8;
  OMBO_CHECK(smem <= 200 * 1024, "acquisition constants (%zu B) exceed shared memory", smem);
  // FP32 arithmetic only in the fast precision mode, and never for the expected decomposition (see k_acquire)
  fp32 = fp32 && acq->kind != OMBO_ACQ_EXPECTED_DECOMP && !ctx->knobs.acq_fp64;
  if (smem > 48 * 1024) {   // per (function, device): set whenever the opt-in is needed
    OMBO_CUDA(cudaFuncSetAttribute(k_acquire<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    OMBO_CUDA(cudaFuncSetAttribute(k_acquire<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  int blocks = (int)((m + ACQ_THREADS - 1) / ACQ_THREADS);
  int rc = ombo_ws_reserve(&ctx->ws_partial, &ctx->ws_partial_bytes, (size_t)blocks * sizeof(ombo_best));
  if (rc) return rc;
  if (fp32)
    k_acquire<float><<<blocks, ACQ_THREADS, smem, s>>>(*acq, n_gp, mu, var, m, ld, index_base, out_acq,
                                                       (ombo_best *)ctx->ws_partial);
  else
    k_acquire<double><<<blocks, ACQ_THREADS, smem, s>>>(*acq, n_gp, mu, var, m, ld, index_base, out_acq,
                                                        (ombo_best *)ctx->ws_partial);
  ctx->launches += 1;
  if (best_dev) {
    k_argmax_final<<<1, 1024, 0, s>>>((const ombo_best *)ctx->ws_partial, blocks, best_dev);
    ctx->launches += 1;
  }
  OMBO_CUDA(cudaGetLastError());
  return OMBO_OK;
}
