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

// Space-to-depth of a channels-last token grid (and its inverse): the s x s patch of pixel rows that one
// spatial-reduction conv output reads becomes one contiguous row of s*s*C values, so the conv is a plain GEMM.
//   packed[b, hs, ws, i, j, :] = img[b, hs*s + i, ws*s + j, :]          (H = Hs*s, W = Ws*s)
// One thread moves one 16-byte chunk; consecutive threads walk the chunks of a pixel row, so both sides are
// accessed in contiguous C*elt-byte runs.
__global__ void __launch_bounds__(256)
space_to_depth_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, long nchunks, int chunks_per_pixel,
                      int Hs, int Ws, int s, int inverse) {
  const long idx = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (idx >= nchunks) return;
  const int c = (int)(idx % chunks_per_pixel);
  long p = idx / chunks_per_pixel;             // packed pixel index: (((b*Hs + hs)*Ws + ws)*s + i)*s + j
  const int j = (int)(p % s);
  p /= s;
  const int i = (int)(p % s);
  p /= s;
  const int ws = (int)(p % Ws);
  p /= Ws;
  const int hs = (int)(p % Hs);
  const long b = p / Hs;
  const long img = ((b * Hs * s + (long)hs * s + i) * ((long)Ws * s) + (long)ws * s + j) * chunks_per_pixel + c;
  if (inverse)
    dst[img] = __ldg(src + idx);
  else
    dst[idx] = __ldg(src + img);
}

// y = act(y + bias[channel]) in place, 16-byte vectors, for the frozen (no-grad) convolution stacks of the
// alignment network (VGG conv + bias + ReLU, BN-folded decoder convs + LeakyReLU): the library convolution
// leaves the bias to a broadcast add and the activation to another elementwise pass.
//   chan_inner == 1: channels-last memory (channel = element index % C), C % VEC == 0
//   chan_inner  > 1: NCHW memory (channel = (index / chan_inner) % C), chan_inner % VEC == 0
template <typename T, int VEC>
__global__ void __launch_bounds__(256)
bias_act_kernel(T* __restrict__ y, const float* __restrict__ bias, long nvec, int C, long chan_inner, int act,
                float slope) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < nvec; i += (long)gridDim.x * blockDim.x) {
    const long e0 = i * VEC;
    float v[VEC];
    if (sizeof(T) == 2) {
      const uint4 u = *reinterpret_cast<const uint4*>(y + e0);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
      for (int k = 0; k < VEC / 2; ++k) {
        const float2 f = __bfloat1622float2(h[k]);
        v[2 * k] = f.x;
        v[2 * k + 1] = f.y;
      }
    } else {
      const float4 f = *reinterpret_cast<const float4*>(y + e0);
      v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
    }
    if (chan_inner == 1) {
      const int c0 = (int)(e0 % C);
#pragma unroll
      for (int k = 0; k < VEC; ++k) v[k] += __ldg(bias + c0 + k);
    } else {
      const float b = __ldg(bias + (int)((e0 / chan_inner) % C));
#pragma unroll
      for (int k = 0; k < VEC; ++k) v[k] += b;
    }
    if (act == 1) {
#pragma unroll
      for (int k = 0; k < VEC; ++k) v[k] = fmaxf(v[k], 0.f);
    } else if (act == 2) {
#pragma unroll
      for (int k = 0; k < VEC; ++k) v[k] = v[k] > 0.f ? v[k] : v[k] * slope;
    }
    if (sizeof(T) == 2) {
      uint4 u;
      __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
      for (int k = 0; k < VEC / 2; ++k) h[k] = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
      *reinterpret_cast<uint4*>(y + e0) = u;
    } else {
      *reinterpret_cast<float4*>(y + e0) = make_float4(v[0], v[1], v[2], v[3]);
    }
  }
}

// scalar variant for channel runs that are not a multiple of the vector width (e.g. the 7 x 7 / 5 x 5 maps of
// the uncertainty decoder's per-pixel patch CNN)
template <typename T>
__global__ void __launch_bounds__(256)
bias_act_scalar_kernel(T* __restrict__ y, const float* __restrict__ bias, long n, int C, long chan_inner, int act,
                       float slope) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    float v = (float)y[i] + __ldg(bias + (int)((i / chan_inner) % C));
    if (act == 1) v = fmaxf(v, 0.f);
    else if (act == 2) v = v > 0.f ? v : v * slope;
    y[i] = (T)v;
  }
}

// 2x2 / stride 2 max-pool of a channels-last bf16 tensor [B,H,W,C] -> [B,H/2,W/2,C] (floor, as nn.MaxPool2d(2, 2)):
// one thread per 8 channels of one output pixel, four 16-byte loads, packed bf16 max (VGG.forward, vgg.py:108-120)
__global__ void __launch_bounds__(256) maxpool2x2_nhwc_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, long nchunks,
                                                              int cpp, int Ho, int Wo, int H, int W) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= nchunks) return;
  const int c = (int)(i % cpp);
  long pix = i / cpp;
  const int xo = (int)(pix % Wo);
  pix /= Wo;
  const int yo = (int)(pix % Ho);
  const long b = pix / Ho;
  const uint4* r0 = x + ((b * H + 2 * yo) * (long)W + 2 * xo) * cpp + c;
  const uint4* r1 = r0 + (long)W * cpp;
  uint4 v[4] = {__ldg(r0), __ldg(r0 + cpp), __ldg(r1), __ldg(r1 + cpp)};
  uint4 o;
  uint32_t* po = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    __nv_bfloat162 m = *reinterpret_cast<const __nv_bfloat162*>(reinterpret_cast<const uint32_t*>(&v[0]) + k);
#pragma unroll
    for (int q = 1; q < 4; ++q) m = __hmax2(m, *reinterpret_cast<const __nv_bfloat162*>(reinterpret_cast<const uint32_t*>(&v[q]) + k));
    po[k] = *reinterpret_cast<const uint32_t*>(&m);
  }
  y[i] = o;
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

extern "C" int rf_space_to_depth(const void* src, void* dst, int B, int H, int W, int row_bytes, int s, int inverse,
                                 void* stream) {
  RF_REQUIRE(src && dst && B > 0 && H > 0 && W > 0 && s > 0, "rf_space_to_depth: bad argument");
  RF_REQUIRE(H % s == 0 && W % s == 0, "rf_space_to_depth: H=%d, W=%d must be multiples of s=%d", H, W, s);
  RF_REQUIRE(row_bytes > 0 && row_bytes % 16 == 0, "rf_space_to_depth: row_bytes=%d must be a multiple of 16", row_bytes);
  RF_REQUIRE((((uintptr_t)src | (uintptr_t)dst) & 15) == 0, "rf_space_to_depth: buffers must be 16-byte aligned");
  const int cpp = row_bytes / 16;
  const long nchunks = (long)B * H * W * cpp;
  const long blocks = (nchunks + 255) / 256;
  RF_REQUIRE(blocks < (1l << 31), "rf_space_to_depth: tensor too large");
  space_to_depth_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const uint4*)src, (uint4*)dst, nchunks, cpp,
                                                                            H / s, W / s, s, inverse);
  RF_CHECK_LAUNCH("space_to_depth_kernel");
  return RF_OK;
}

extern "C" int rf_bias_act(void* y, const float* bias, int64_t numel, int C, int64_t chan_inner, int act, float slope,
                           int dtype, void* stream) {
  RF_REQUIRE(y && bias && numel > 0 && C > 0 && chan_inner > 0, "rf_bias_act: bad argument");
  RF_REQUIRE(dtype == 0 || dtype == 1, "rf_bias_act: dtype must be 0 (f32) or 1 (bf16)");
  RF_REQUIRE(act >= 0 && act <= 2, "rf_bias_act: act must be 0 (none), 1 (relu) or 2 (leaky relu)");
  const int vec = dtype == 1 ? 8 : 4;
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec_ok = (((uintptr_t)y & 15) == 0) && (numel % vec == 0) &&
                      (chan_inner == 1 ? (C % vec == 0) : (chan_inner % vec == 0));
  if (!vec_ok) {
    long sb = (numel + 255) / 256;
    if (sb > (long)kNumSMs * 32) sb = (long)kNumSMs * 32;
    if (dtype == 1)
      bias_act_scalar_kernel<__nv_bfloat16><<<(unsigned)sb, 256, 0, st>>>((__nv_bfloat16*)y, bias, numel, C, chan_inner, act, slope);
    else
      bias_act_scalar_kernel<float><<<(unsigned)sb, 256, 0, st>>>((float*)y, bias, numel, C, chan_inner, act, slope);
    RF_CHECK_LAUNCH("bias_act_scalar_kernel");
    return RF_OK;
  }
  const long nvec = numel / vec;
  long blocks = (nvec + 255) / 256;
  if (blocks > (long)kNumSMs * 16) blocks = (long)kNumSMs * 16;
  if (dtype == 1)
    bias_act_kernel<__nv_bfloat16, 8><<<(unsigned)blocks, 256, 0, st>>>((__nv_bfloat16*)y, bias, nvec, C, chan_inner, act, slope);
  else
    bias_act_kernel<float, 4><<<(unsigned)blocks, 256, 0, st>>>((float*)y, bias, nvec, C, chan_inner, act, slope);
  RF_CHECK_LAUNCH("bias_act_kernel");
  return RF_OK;
}

extern "C" int rf_maxpool2x2_nhwc_bf16(const void* x, void* y, int B, int H, int W, int C, void* stream) {
  RF_REQUIRE(x && y && B > 0 && H >= 2 && W >= 2 && C > 0, "rf_maxpool2x2_nhwc_bf16: bad argument");
  RF_REQUIRE(C % 8 == 0, "rf_maxpool2x2_nhwc_bf16: C=%d must be a multiple of 8", C);
  RF_REQUIRE((((uintptr_t)x | (uintptr_t)y) & 15) == 0, "rf_maxpool2x2_nhwc_bf16: buffers must be 16-byte aligned");
  const int cpp = C / 8, Ho = H / 2, Wo = W / 2;
  const long nchunks = (long)B * Ho * Wo * cpp;
  const long blocks = (nchunks + 255) / 256;
  RF_REQUIRE(blocks < (1l << 31), "rf_maxpool2x2_nhwc_bf16: tensor too large");
  maxpool2x2_nhwc_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const uint4*)x, (uint4*)y, nchunks, cpp, Ho, Wo, H, W);
  RF_CHECK_LAUNCH("maxpool2x2_nhwc_kernel");
  return RF_OK;
}
