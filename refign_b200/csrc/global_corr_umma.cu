// tcgen05 (UMMA) TF32 path of the global correlation -- placeholder until the
// tensor-core kernel lands; reports "unsupported" so callers take the FFMA path.
#include "rf_common.cuh"
namespace rf {
bool global_corr_umma_supported(int, long, long, const void*, const void*, const void*) { return false; }
int global_corr_umma(const float*, const float*, float*, float*, float*, int, int, long, long, cudaStream_t) {
  set_error("global_corr_umma: not built");
  return RF_EINVAL;
}
}  // namespace rf
