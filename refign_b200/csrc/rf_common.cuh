// Shared helpers for the refign_b200 sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/refign_b200.h"

namespace rf {

void set_error(const char* fmt, ...);

#define RF_REQUIRE(cond, ...)        \
  do {                               \
    if (!(cond)) {                   \
      rf::set_error(__VA_ARGS__);    \
      return RF_EINVAL;              \
    }                                \
  } while (0)

// Checks the launch (not the completion) of the preceding kernel.
#define RF_CHECK_LAUNCH(name)                                                  \
  do {                                                                         \
    cudaError_t e_ = cudaGetLastError();                                       \
    if (e_ != cudaSuccess) {                                                   \
      rf::set_error("%s: launch failed: %s", name, cudaGetErrorString(e_));    \
      return RF_ECUDA;                                                         \
    }                                                                          \
  } while (0)

#define RF_CUDA(call)                                                          \
  do {                                                                         \
    cudaError_t e_ = (call);                                                   \
    if (e_ != cudaSuccess) {                                                   \
      rf::set_error("%s: %s", #call, cudaGetErrorString(e_));                  \
      return RF_ECUDA;                                                         \
    }                                                                          \
  } while (0)

constexpr int kNumSMs = 148;  // B200

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ long long warp_sum_i64(long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// 16-byte async copy global->shared; bytes==0 zero-fills the destination.
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem),
               "r"(src_bytes));
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem, int src_bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(s), "l"(gmem),
               "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// streaming (evict-first) vector store for write-once outputs
__device__ __forceinline__ void st_cs_f4(float* p, float4 v) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};\n" ::"l"(p), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w));
}

}  // namespace rf
