// Candidate pool access: explicit (m, d) buffers (f64 / f32) or the counter-based generator.
// The generator is the bit-exact twin of oracle/oracle.py::counter_uniform (splitmix64
// finaliser keyed by (seed, global_index * d + j)), so any shard on any GPU regenerates the
// same pool and the host can regenerate rows for parity checks.
#pragma once
#include "common.cuh"

__device__ __forceinline__ double ombo_counter_u01(unsigned long long seed, unsigned long long ctr) {
  unsigned long long z = ctr + (seed + 1ull) * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}

// coordinate j of local candidate row c (global index = index_base + c)
__device__ __forceinline__ double ombo_pool_coord(const PoolDev &p, long long c, int j) {
  if (p.X != nullptr) {
    if (p.dtype == 0) return ((const double *)p.X)[c * p.d + j];
    return (double)((const float *)p.X)[c * p.d + j];
  }
  unsigned long long g = (unsigned long long)(p.index_base + c);
  double u = ombo_counter_u01(p.seed, g * (unsigned long long)p.d + (unsigned long long)j);
  // no FMA contraction: must round like numpy's lo + span * u
  return __dadd_rn(p.lo[j], __dmul_rn(p.span[j], u));
}
