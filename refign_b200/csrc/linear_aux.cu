// Helpers around the library GEMMs of the MiT / DAFormer Linear layers (sm_100a):
//   * rf_colsum      : bias gradient  db[c] = sum_rows g[row, c]   (the reference / autograd runs a generic
//                      reduce kernel per Linear, ~20 us each at these shapes; this is one coalesced pass);
//   * rf_cast_bf16   : fp32 -> bf16 copy of a flat parameter buffer (the bf16 "shadow" weights the
//                      tensor-core GEMMs read; refreshed once per step after AdamW / EMA instead of
//                      ~3 000 per-tensor autocast casts per step).
// The GEMMs themselves stay on cuBLASLt (plain library GEMMs).
#include <cuda_bf16.h>

#include "rf_common.cuh"

namespace rf {

// CTA = 8 column groups (8 columns each: one 128-byte line of bf16 per row) x 32 row lanes, so narrow
// matrices (the 64-column stage-1 Linears) keep every lane busy; four rows are in flight per thread.
constexpr int CS_CG = 8, CS_RL = 32, CS_UNROLL = 4;

template <typename T>
__device__ __forceinline__ void cs_load8(const T* p, float (&v)[8]);
template <>
__device__ __forceinline__ void cs_load8<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[8]) {
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}
template <>
__device__ __forceinline__ void cs_load8<float>(const float* p, float (&v)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

template <typename T>
__global__ void __launch_bounds__(CS_CG * CS_RL)
colsum_kernel(const T* __restrict__ g, float* __restrict__ out, long rows, int cols, long strip) {
  __shared__ float red[CS_RL][CS_CG * 8 + 1];
  const int lane_cg = threadIdx.x % CS_CG, rl = threadIdx.x / CS_CG;
  const int cg = blockIdx.x * CS_CG + lane_cg;
  const bool live = cg * 8 < cols;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (live) {
    const long r0 = (long)blockIdx.y * strip;
    const long r1 = (r0 + strip < rows) ? r0 + strip : rows;
    const T* p = g + cg * 8;
    long r = r0 + rl;
    for (; r + (CS_UNROLL - 1) * CS_RL < r1; r += CS_UNROLL * CS_RL) {
      float v[CS_UNROLL][8];
#pragma unroll
      for (int u = 0; u < CS_UNROLL; ++u) cs_load8<T>(p + (r + u * CS_RL) * cols, v[u]);
#pragma unroll
      for (int u = 0; u < CS_UNROLL; ++u)
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] += v[u][k];
    }
    for (; r < r1; r += CS_RL) {
      float v[8];
      cs_load8<T>(p + r * cols, v);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] += v[k];
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) red[rl][lane_cg * 8 + k] = acc[k];
  __syncthreads();
  if (threadIdx.x < CS_CG * 8) {
    float t = 0.f;
#pragma unroll 8
    for (int i = 0; i < CS_RL; ++i) t += red[i][threadIdx.x];
    const int c = blockIdx.x * CS_CG * 8 + threadIdx.x;
    if (c < cols) atomicAdd(out + c, t);
  }
}

__global__ void __launch_bounds__(256)
cast_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long n) {
  const long n8 = n >> 3;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n8; i += (long)gridDim.x * blockDim.x) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(src) + 2 * i);
    const float4 b = __ldg(reinterpret_cast<const float4*>(src) + 2 * i + 1);
    uint4 u;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
    h[0] = __floats2bfloat162_rn(a.x, a.y);
    h[1] = __floats2bfloat162_rn(a.z, a.w);
    h[2] = __floats2bfloat162_rn(b.x, b.y);
    h[3] = __floats2bfloat162_rn(b.z, b.w);
    reinterpret_cast<uint4*>(dst)[i] = u;
  }
  for (long i = (n8 << 3) + blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    dst[i] = __float2bfloat16_rn(src[i]);
}

}  // namespace rf

using namespace rf;

extern "C" int rf_colsum(const void* g, float* out, int64_t rows, int cols, int dtype, int accumulate,
                         void* stream) {
  RF_REQUIRE(g && out && rows > 0 && cols > 0, "rf_colsum: bad argument");
  RF_REQUIRE(cols % 8 == 0, "rf_colsum: cols=%d must be a multiple of 8", cols);
  RF_REQUIRE(((uintptr_t)g & 15) == 0, "rf_colsum: input must be 16-byte aligned");
  RF_REQUIRE(dtype == 0 || dtype == 1, "rf_colsum: dtype must be 0 (f32) or 1 (bf16)");
  cudaStream_t st = (cudaStream_t)stream;
  if (!accumulate) RF_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * cols, st));
  const int gx = (cols / 8 + CS_CG - 1) / CS_CG;
  long strip = rows * gx / ((long)kNumSMs * 4);   // ~4 CTAs per SM, at least 8 rows per thread
  if (strip < CS_RL * 8) strip = CS_RL * 8;
  if (strip > 8192) strip = 8192;
  const long gy = (rows + strip - 1) / strip;
  RF_REQUIRE(gy <= 65535, "rf_colsum: too many row strips");
  dim3 grid((unsigned)gx, (unsigned)gy);
  if (dtype == 1)
    colsum_kernel<__nv_bfloat16><<<grid, CS_CG * CS_RL, 0, st>>>((const __nv_bfloat16*)g, out, rows, cols, strip);
  else
    colsum_kernel<float><<<grid, CS_CG * CS_RL, 0, st>>>((const float*)g, out, rows, cols, strip);
  RF_CHECK_LAUNCH("colsum_kernel");
  return RF_OK;
}

extern "C" int rf_cast_bf16(const float* src, void* dst, int64_t n, void* stream) {
  RF_REQUIRE(src && dst && n > 0, "rf_cast_bf16: bad argument");
  RF_REQUIRE((((uintptr_t)src | (uintptr_t)dst) & 15) == 0, "rf_cast_bf16: buffers must be 16-byte aligned");
  long blocks = ceil_div(n / 8 + 1, 256);
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  cast_bf16_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(src, (__nv_bfloat16*)dst, n);
  RF_CHECK_LAUNCH("cast_bf16_kernel");
  return RF_OK;
}
