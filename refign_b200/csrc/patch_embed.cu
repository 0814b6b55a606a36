// Stage-1 OverlapPatchEmbed of the MiT encoder for sm_100a: 7x7 / stride 4 / pad 3 convolution with 3
// input channels fused with the LayerNorm over the embedding channels.
// Restates OverlapPatchEmbed.forward (/root/reference/models/backbones/mix_transformer.py:236-242):
//   x = self.proj(x); x = x.flatten(2).transpose(1, 2); x = self.norm(x)       (LayerNorm eps 1e-5, :234)
// The reference runs a cuDNN convolution with C_in = 3 (poor tensor-core utilisation), an NCHW -> NLC
// transpose copy and a LayerNorm: three passes.  Here one CTA produces an 8 x 8 tile of tokens: the
// 35 x 35 x 3 input window and the [147][COUT] weights are staged in shared memory, each thread
// accumulates 16 channels of one token in registers (fp32), the LayerNorm statistics are a 4-lane
// shuffle reduction, and tokens leave as contiguous 256-byte rows [B, H/4 * W/4, COUT].
// HBM-bound: bytes = 4 B (3 H W + COUT H W / 16) (+ the pre-norm copy when training).
#include "rf_common.cuh"

namespace rf {

constexpr int PE_T = 8;                    // output tile edge (tokens)
constexpr int PE_IN = PE_T * 4 + 3;        // 35 input rows / columns per tile
constexpr int PE_TAPS = 147;               // 3 * 7 * 7

template <int COUT>
__global__ void __launch_bounds__(PE_T * PE_T * (COUT / 16))
patch_embed_ln_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                      const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ pre,
                      float* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out, int B,
                      int H, int W, int Ho, int Wo, float eps) {
  constexpr int G = COUT / 16;             // threads per token
  extern __shared__ float smem[];
  float* sW = smem;                        // [147][COUT]
  float* sIn = smem + PE_TAPS * COUT;      // [3][35][36]
  const int tid = threadIdx.x;
  const int b = blockIdx.z, ty0 = blockIdx.y * PE_T, tx0 = blockIdx.x * PE_T;
  // weights: global [COUT][3][7][7] -> smem [tap][COUT]
  for (int i = tid; i < PE_TAPS * COUT; i += blockDim.x) {
    const int co = i / PE_TAPS, tap = i % PE_TAPS;
    sW[tap * COUT + co] = __ldg(w + i);
  }
  const int iy0 = ty0 * 4 - 3, ix0 = tx0 * 4 - 3;
  for (int i = tid; i < 3 * PE_IN * PE_IN; i += blockDim.x) {
    const int c = i / (PE_IN * PE_IN), r = (i / PE_IN) % PE_IN, col = i % PE_IN;
    const int iy = iy0 + r, ix = ix0 + col;
    float v = 0.f;
    if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = __ldg(x + (((long)b * 3 + c) * H + iy) * W + ix);
    sIn[(c * PE_IN + r) * (PE_IN + 1) + col] = v;
  }
  __syncthreads();
  const int tok = tid / G, cg = tid % G;
  const int ty = tok / PE_T, tx = tok % PE_T;
  float acc[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) acc[k] = __ldg(bias + cg * 16 + k);
  for (int c = 0; c < 3; ++c)
    for (int ky = 0; ky < 7; ++ky) {
      const float* in = sIn + (c * PE_IN + ty * 4 + ky) * (PE_IN + 1) + tx * 4;
      const float* wr = sW + ((c * 7 + ky) * 7) * COUT + cg * 16;
#pragma unroll
      for (int kx = 0; kx < 7; ++kx) {
        const float a = in[kx];
        const float4* w4 = reinterpret_cast<const float4*>(wr + kx * COUT);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 ww = w4[q];
          acc[4 * q + 0] = fmaf(a, ww.x, acc[4 * q + 0]);
          acc[4 * q + 1] = fmaf(a, ww.y, acc[4 * q + 1]);
          acc[4 * q + 2] = fmaf(a, ww.z, acc[4 * q + 2]);
          acc[4 * q + 3] = fmaf(a, ww.w, acc[4 * q + 3]);
        }
      }
    }
  // LayerNorm over COUT channels = G consecutive lanes
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 16; ++k) s += acc[k];
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / (float)COUT;
  float sq = 0.f;
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const float d = acc[k] - mean;
    sq = fmaf(d, d, sq);
  }
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = rsqrtf(sq / (float)COUT + eps);
  const int oy = ty0 + ty, ox = tx0 + tx;
  if (oy < Ho && ox < Wo) {
    const long t = ((long)b * Ho + oy) * Wo + ox;
    float4* dst = reinterpret_cast<float4*>(y + t * COUT + cg * 16);
    float4* dpre = pre ? reinterpret_cast<float4*>(pre + t * COUT + cg * 16) : nullptr;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + cg * 16) + q);
      const float4 be = __ldg(reinterpret_cast<const float4*>(beta + cg * 16) + q);
      if (dpre) dpre[q] = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
      dst[q] = make_float4(fmaf((acc[4 * q] - mean) * rstd, g.x, be.x), fmaf((acc[4 * q + 1] - mean) * rstd, g.y, be.y),
                           fmaf((acc[4 * q + 2] - mean) * rstd, g.z, be.z), fmaf((acc[4 * q + 3] - mean) * rstd, g.w, be.w));
    }
    if (cg == 0 && mean_out) {
      mean_out[t] = mean;
      rstd_out[t] = rstd;
    }
  }
}

}  // namespace rf

using namespace rf;

extern "C" int rf_patch_embed_ln_fwd(const float* x, const float* weight, const float* bias, const float* gamma,
                                     const float* beta, float* pre_norm, float* y, float* mean, float* rstd, int B,
                                     int H, int W, int cout, float eps, void* stream) {
  RF_REQUIRE(x && weight && bias && gamma && beta && y, "rf_patch_embed_ln_fwd: null pointer");
  RF_REQUIRE(B > 0 && H >= 7 && W >= 7, "rf_patch_embed_ln_fwd: bad shape");
  RF_REQUIRE(cout == 64 || cout == 32, "rf_patch_embed_ln_fwd: embed dim %d not supported (32 or 64)", cout);
  RF_REQUIRE((mean == nullptr) == (rstd == nullptr), "rf_patch_embed_ln_fwd: mean and rstd go together");
  const int Ho = (H + 6 - 7) / 4 + 1, Wo = (W + 6 - 7) / 4 + 1;
  dim3 grid((Wo + PE_T - 1) / PE_T, (Ho + PE_T - 1) / PE_T, B);
  RF_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "rf_patch_embed_ln_fwd: grid too large");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = sizeof(float) * ((size_t)PE_TAPS * cout + 3 * PE_IN * (PE_IN + 1));
  static bool attr_set = false;
  if (!attr_set) {
    RF_CUDA(cudaFuncSetAttribute(patch_embed_ln_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)(sizeof(float) * (PE_TAPS * 64 + 3 * PE_IN * (PE_IN + 1)))));
    attr_set = true;
  }
  if (cout == 64)
    patch_embed_ln_kernel<64><<<grid, PE_T * PE_T * 4, smem, st>>>(x, weight, bias, gamma, beta, pre_norm, y, mean, rstd,
                                                                   B, H, W, Ho, Wo, eps);
  else
    patch_embed_ln_kernel<32><<<grid, PE_T * PE_T * 2, smem, st>>>(x, weight, bias, gamma, beta, pre_norm, y, mean, rstd,
                                                                   B, H, W, Ho, Wo, eps);
  RF_CHECK_LAUNCH("patch_embed_ln_kernel");
  return RF_OK;
}
