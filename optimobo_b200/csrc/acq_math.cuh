// Per-candidate acquisition arithmetic shared by K4 (acquisition.cu) and by the fast posterior kernel's fused
// epilogue (posterior_fast8.cu): the normal cdf / pdf pair and the closed-form 2-D EHVI.
// Restates util_functions.py:136-167 (EHVI) + :81-128 (EHVI_2D_aux) + :130-133 (psi_cal) of the reference tree.
#pragma once

__device__ __forceinline__ double Phi(double t) { return 0.5 * erfc(-t * 0.70710678118654752440); }
__device__ __forceinline__ double phi(double t) { return exp(-(t * t) / 2.0) / 2.50662827463100050242; }
// FP32 twins for the fast precision mode (tolerance 1e-3): erfcf keeps its relative accuracy in the tails,
// exp goes through MUFU.EX2 -- the FP64 pipe of the B200 is ~60x narrower than the FP32 one
__device__ __forceinline__ float Phi(float t) { return 0.5f * erfcf(-t * 0.70710678f); }
__device__ __forceinline__ float phi(float t) { return __expf(-0.5f * t * t) * 0.39894228f; }
__device__ __forceinline__ float sqrt_t(float x) { return sqrtf(x); }
__device__ __forceinline__ double sqrt_t(double x) { return sqrt(x); }
__device__ __forceinline__ float fmax_t(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ double fmax_t(double a, double b) { return fmax(a, b); }

// y1 / y2: the P + 2 stripe bounds per objective (host_prep.ehvi_stripes: front sorted by f1, reference point and
// ideal sentinel at the ends).  exact = 0: the reference's arithmetic ('sigma' = model 0's variance times the
// flattened sample covariance, util_functions.py:163,167,:115; P stripes); exact = 1: each model's own standard
// deviation and the (P+1)-th stripe the reference drops (y1[P+1] = -inf taken as a limit).
template <typename T>
__device__ __forceinline__ T ehvi2d_value(T m0, T m1, T v0, T v1, const T *__restrict__ y1, const T *__restrict__ y2,
                                          int P, bool exact, T c00, T c01) {
  const T s0 = exact ? sqrt_t(v0) : v0 * c00;
  const T s1 = exact ? sqrt_t(v1) : v0 * c01;
  T sum1 = 0, sum2 = 0;
  T tp = (y1[0] - m0) / s0;
  T cdf_p = Phi(tp), pdf_p = phi(tp);
  for (int i = 1; i <= P; ++i) {
    T t = (y1[i] - m0) / s0;
    T cdf_t = Phi(t), pdf_t = phi(t);
    T t2 = (y2[i] - m1) / s1;
    T psi2 = s1 * phi(t2) + (y2[i] - m1) * Phi(t2);
    sum1 = sum1 + (y1[i - 1] - y1[i]) * cdf_t * psi2;
    T psi_a = s0 * pdf_p + (y1[i - 1] - m0) * cdf_p;
    T psi_b = s0 * pdf_t + (y1[i - 1] - m0) * cdf_t;
    sum2 = sum2 + (psi_a - psi_b) * psi2;
    cdf_p = cdf_t; pdf_p = pdf_t;
  }
  if (exact) {
    T t2 = (y2[P + 1] - m1) / s1;
    T psi2 = s1 * phi(t2) + (y2[P + 1] - m1) * Phi(t2);
    T psi_a = s0 * pdf_p + (y1[P] - m0) * cdf_p;
    sum2 = sum2 + psi_a * psi2;
  }
  return sum1 + sum2;
}

// arg-max pair: ties resolve to the lowest global index (np.argmax)
struct BestPair { double v; long long i; };
__device__ __forceinline__ bool better(double v, long long i, double bv, long long bi) {
  return (v > bv) || (v == bv && i < bi);
}
__device__ __forceinline__ BestPair warp_best(BestPair b) {
  for (int o = 16; o; o >>= 1) {
    double ov = __shfl_xor_sync(0xffffffffu, b.v, o);
    long long oi = __shfl_xor_sync(0xffffffffu, b.i, o);
    if (better(ov, oi, b.v, b.i)) { b.v = ov; b.i = oi; }
  }
  return b;
}
