// Stage-1 OverlapPatchEmbed of the MiT encoder for sm_100a: 7x7 / stride 4 / pad 3 convolution with 3
// input channels fused with the LayerNorm over the embedding channels.
// Restates OverlapPatchEmbed.forward (/root/reference/models/backbones/mix_transformer.py:236-242):
//   x = self.proj(x); x = x.flatten(2).transpose(1, 2); x = self.norm(x)       (LayerNorm eps 1e-5, :234)
// The reference runs a cuDNN convolution with C_in = 3 (poor tensor-core utilisation), an NCHW -> NLC
// transpose copy and a LayerNorm: three passes.  Here a CTA produces 8 x 8 tiles of tokens: the
// 35 x 35 x 3 input window and the [147][COUT] weights are staged in shared memory, each thread
// accumulates 8 channels of 4 tokens in registers (fp32, packed FFMA2), the LayerNorm statistics are a
// shuffle reduction, and tokens leave as contiguous 256-byte rows [B, H/4 * W/4, COUT].
// Bytes = 4 B (3 H W + COUT H W / 16) (+ the pre-norm copy when training); with 147 FMAs per output the
// kernel sits at the fp32-FMA ridge rather than the HBM roofline (2 B Ho Wo COUT 147 flop).
#include "rf_common.cuh"

namespace rf {

constexpr int PE_T = 8;                    // output tile edge (tokens)
constexpr int PE_IN = PE_T * 4 + 3;        // 35 input rows / columns per tile
constexpr int PE_INP = PE_IN + 1;          // padded input row pitch (36 floats: 16-byte aligned rows)
constexpr int PE_TAPS = 147;               // 3 * 7 * 7
constexpr int PE_TOK = 4;                  // tokens per thread (consecutive along x)

// One CTA walks a strided list of 8 x 8 token tiles with the [147][COUT] weights resident in shared memory
// (staged once per CTA, pitch COUT + 4 so the transposing store is 4-way instead of 32-way bank-conflicted).
// Thread (cg = tid % (COUT/8), tg = tid / (COUT/8)) owns 8 channels of 4 x-consecutive tokens: per (channel-in,
// ky) row it reads 19 inputs (5 LDS.128, shared by the 7 kx taps of the 4 tokens) and 7 x 2 LDS.128 of weights
// for 7 x 16 packed FFMA2 (scalar input x channel-pair weight), i.e. the kernel is FMA-issue bound, not
// shared-memory bound.  LayerNorm statistics: shuffle reduction over the COUT/8 lanes of a token.
template <int COUT>
__global__ void __launch_bounds__(16 * (COUT / 8))
patch_embed_ln_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                      const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ pre,
                      float* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out, int B,
                      int H, int W, int Ho, int Wo, float eps, int tiles_x, int tiles_y) {
  constexpr int G = COUT / 8;              // threads per token group (8 channels each)
  constexpr int NT = 16 * G;
  constexpr int WP = COUT + 4;             // weight row pitch
  extern __shared__ __align__(16) float smem[];
  float* sW = smem;                        // [147][WP]
  float* sIn = smem + PE_TAPS * WP;        // [3][35][36]
  const int tid = threadIdx.x;
  for (int i = tid; i < PE_TAPS * COUT; i += NT) {   // global [COUT][3][7][7] -> smem [tap][COUT]
    const int co = i / PE_TAPS, tap = i - co * PE_TAPS;
    sW[tap * WP + co] = __ldg(w + i);
  }
  const int cg = tid % G, tg = tid / G;
  const int ty = tg >> 1, tx4 = (tg & 1) * PE_TOK;   // tile row, first of the 4 tokens
  float bv[8], gv[8], bev[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    bv[k] = __ldg(bias + cg * 8 + k);
    gv[k] = __ldg(gamma + cg * 8 + k);
    bev[k] = __ldg(beta + cg * 8 + k);
  }
  const int ntiles = tiles_x * tiles_y * B;
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int b = t / (tiles_x * tiles_y), tr = t - b * (tiles_x * tiles_y);
    const int ty0 = (tr / tiles_x) * PE_T, tx0 = (tr % tiles_x) * PE_T;
    const int iy0 = ty0 * 4 - 3, ix0 = tx0 * 4 - 3;
    __syncthreads();                       // previous tile's readers (and the weight staging) are done
    for (int i = tid; i < 3 * PE_IN * PE_INP; i += NT) {
      const int c = i / (PE_IN * PE_INP), r = (i / PE_INP) % PE_IN, col = i % PE_INP;
      const int iy = iy0 + r, ix = ix0 + col;
      float v = 0.f;
      if (col < PE_IN && iy >= 0 && iy < H && ix >= 0 && ix < W) v = __ldg(x + (((long)b * 3 + c) * H + iy) * W + ix);
      sIn[i] = v;
    }
    __syncthreads();
    float2 acc[PE_TOK][4];
#pragma unroll
    for (int q = 0; q < PE_TOK; ++q)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[q][k] = make_float2(bv[2 * k], bv[2 * k + 1]);
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int ky = 0; ky < 7; ++ky) {
        const float4* in4 = reinterpret_cast<const float4*>(sIn + (c * PE_IN + ty * 4 + ky) * PE_INP + tx4 * 4);
        float in[20];
#pragma unroll
        for (int v = 0; v < 5; ++v) {
          const float4 f = in4[v];
          in[4 * v] = f.x; in[4 * v + 1] = f.y; in[4 * v + 2] = f.z; in[4 * v + 3] = f.w;
        }
        const float* wr = sW + ((c * 7 + ky) * 7) * WP + cg * 8;
#pragma unroll
        for (int kx = 0; kx < 7; ++kx) {
          const float4 w0 = *reinterpret_cast<const float4*>(wr + kx * WP);
          const float4 w1 = *reinterpret_cast<const float4*>(wr + kx * WP + 4);
          const float2 wp[4] = {make_float2(w0.x, w0.y), make_float2(w0.z, w0.w), make_float2(w1.x, w1.y),
                                make_float2(w1.z, w1.w)};
#pragma unroll
          for (int q = 0; q < PE_TOK; ++q) {
            const float a = in[4 * q + kx];
            const float2 a2 = make_float2(a, a);
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[q][k] = __ffma2_rn(a2, wp[k], acc[q][k]);
          }
        }
      }
    // LayerNorm over COUT channels = G consecutive lanes, per token
#pragma unroll
    for (int q = 0; q < PE_TOK; ++q) {
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) s += acc[q][k].x + acc[q][k].y;
#pragma unroll
      for (int o = G / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      const float mean = s / (float)COUT;
      float sq = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float d0 = acc[q][k].x - mean, d1 = acc[q][k].y - mean;
        sq = fmaf(d0, d0, sq);
        sq = fmaf(d1, d1, sq);
      }
#pragma unroll
      for (int o = G / 2; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
      const float rstd = rsqrtf(sq / (float)COUT + eps);
      const int oy = ty0 + ty, ox = tx0 + tx4 + q;
      if (oy < Ho && ox < Wo) {
        const long tk = ((long)b * Ho + oy) * Wo + ox;
        float4* dst = reinterpret_cast<float4*>(y + tk * COUT + cg * 8);
        if (pre) {
          float4* dpre = reinterpret_cast<float4*>(pre + tk * COUT + cg * 8);
          dpre[0] = make_float4(acc[q][0].x, acc[q][0].y, acc[q][1].x, acc[q][1].y);
          dpre[1] = make_float4(acc[q][2].x, acc[q][2].y, acc[q][3].x, acc[q][3].y);
        }
        float o8[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          o8[2 * k] = fmaf((acc[q][k].x - mean) * rstd, gv[2 * k], bev[2 * k]);
          o8[2 * k + 1] = fmaf((acc[q][k].y - mean) * rstd, gv[2 * k + 1], bev[2 * k + 1]);
        }
        dst[0] = make_float4(o8[0], o8[1], o8[2], o8[3]);
        dst[1] = make_float4(o8[4], o8[5], o8[6], o8[7]);
        if (cg == 0 && mean_out) {
          mean_out[tk] = mean;
          rstd_out[tk] = rstd;
        }
      }
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
  const int tiles_x = (Wo + PE_T - 1) / PE_T, tiles_y = (Ho + PE_T - 1) / PE_T;
  const long ntiles = (long)tiles_x * tiles_y * B;
  RF_REQUIRE(ntiles < (1l << 31), "rf_patch_embed_ln_fwd: too many tiles");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = sizeof(float) * ((size_t)PE_TAPS * (cout + 4) + 3 * PE_IN * PE_INP);
  static bool attr_set = false;
  if (!attr_set) {
    RF_CUDA(cudaFuncSetAttribute(patch_embed_ln_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)(sizeof(float) * (PE_TAPS * (64 + 4) + 3 * PE_IN * PE_INP))));
    attr_set = true;
  }
  // CTAs keep the weights resident and walk tiles: ~4 CTAs per SM (55 KB of shared memory each)
  const int grid = (int)(ntiles < (long)kNumSMs * 4 ? ntiles : (long)kNumSMs * 4);
  if (cout == 64)
    patch_embed_ln_kernel<64><<<grid, 128, smem, st>>>(x, weight, bias, gamma, beta, pre_norm, y, mean, rstd, B, H, W,
                                                       Ho, Wo, eps, tiles_x, tiles_y);
  else
    patch_embed_ln_kernel<32><<<grid, 64, smem, st>>>(x, weight, bias, gamma, beta, pre_norm, y, mean, rstd, B, H, W,
                                                      Ho, Wo, eps, tiles_x, tiles_y);
  RF_CHECK_LAUNCH("patch_embed_ln_kernel");
  return RF_OK;
}
