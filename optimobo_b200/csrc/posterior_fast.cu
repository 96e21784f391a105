// K1 + K2 fast mode (bf16x3 split on tcgen05) -- placeholder until the tensor-core path lands.
#include "common.cuh"
int ombo_fast_path_built() { return 0; }
int ombo_posterior_fast(ombo_ctx *, const GpDev &, const PoolDev &, long long, double *, double *, cudaStream_t) {
  ombo_set_error("fast precision mode is not built yet");
  return OMBO_ERR_UNSUPPORTED;
}
