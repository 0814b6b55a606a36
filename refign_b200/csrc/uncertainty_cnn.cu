// Fused per-pixel patch CNN of the UAWarpC UncertaintyModule (reference models/modules.py:534-561).
// The reference reshapes the correlation volume [B, s*s, H, W] into B*H*W single-channel s x s images and runs four
// "valid" 3x3 convolutions over that giant batch (conv_0 1->32 [+ 2x2 max-pool when s = 16], conv_1 32->32, conv_2
// 32->16, predict_uncertainty 16->6; eval-mode BatchNorm + LeakyReLU(0.1) after the first three), which materialises
// [B*H*W, 32, 7, 7] activations in HBM (822 MB in fp32 at 2 x 256^2) for 574 kFLOP of work per pixel.  Here one CTA
// takes 16 pixels at a time and keeps every intermediate in shared memory:
//   patches  fp32 [16][s*s]           gathered from the displacement planes (64-byte segments per plane)
//   conv_0   CUDA cores (K = 9 is too short for a tensor-core tile), BN folded, -> bf16 [16][49][32]
//   conv_1   mma.sync m16n8k16 bf16: rows = (pixel, output position) flattened (16 x 25 = 25 m-tiles, no padding
//            rows), K = (tap, cin) = 288, the A fragments are gathered by ldmatrix straight from the conv_0 tile
//            (the im2col matrix is never built), -> bf16 [16][25][32]
//   conv_2   same, rows = 16 x 9, N = 16, -> bf16 [16][9][16]
//   predict  CUDA cores, 6 x 144 dot products per pixel -> out bf16 [B, H, W, 6]
// HBM traffic is the volume read once + 12 bytes per pixel written.  bf16 activations / fp32 accumulation: the same
// precision as the library path under bf16 autocast (the fp32 parity mode keeps the library convolutions).
#include <cuda_bf16.h>

#include "rf_common.cuh"

namespace rf {

constexpr int UC_PIX = 16;           // pixels per CTA iteration
constexpr int UC_THREADS = 512;          // 16 warps: the phases are latency-bound (dependent LDSM -> HMMA -> STS chains), not throughput-bound
constexpr int UC_APITCH = 40;        // bf16 elements per activation row (32 channels + 8 pad: conflict-free ldmatrix rows)
constexpr int UC_WPITCH = 296;       // bf16 elements per filter row (288 + 8 pad)
// parameter block (bytes), built by refign_b200/modules.py: UncertaintyModule._fused_params
constexpr int UC_OFF_W0 = 0;                       // f32 [9][32]   conv_0 (BN folded), tap-major
constexpr int UC_OFF_B0 = UC_OFF_W0 + 9 * 32 * 4;  // f32 [32]
constexpr int UC_OFF_B1 = UC_OFF_B0 + 32 * 4;      // f32 [32]
constexpr int UC_OFF_B2 = UC_OFF_B1 + 32 * 4;      // f32 [16]
constexpr int UC_OFF_W3 = UC_OFF_B2 + 16 * 4;      // f32 [6][9][16] predict_uncertainty, k = position * 16 + channel
constexpr int UC_OFF_B3 = UC_OFF_W3 + 6 * 144 * 4; // f32 [8] (6 used)
constexpr int UC_OFF_W1 = UC_OFF_B3 + 8 * 4;       // bf16 [32][296]  conv_1, k = tap * 32 + cin
constexpr int UC_OFF_W2 = UC_OFF_W1 + 32 * UC_WPITCH * 2;   // bf16 [16][296]
constexpr int UC_PARAM_BYTES = UC_OFF_W2 + 16 * UC_WPITCH * 2;
static_assert(UC_OFF_W1 % 16 == 0 && UC_PARAM_BYTES % 16 == 0, "parameter block alignment");

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x2(uint32_t (&r)[2], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float leaky(float x, float slope) { return x > 0.f ? x : x * slope; }
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

template <int S>
struct UcSmem {
  float patch[UC_PIX][S * S];
  __align__(16) __nv_bfloat16 act0[UC_PIX * 49 * UC_APITCH];
  __align__(16) __nv_bfloat16 act1[UC_PIX * 25 * UC_APITCH];
  __align__(16) __nv_bfloat16 act2[UC_PIX * 9 * 16];
  __align__(16) uint8_t params[UC_PARAM_BYTES];
};

template <int S>
__global__ void __launch_bounds__(UC_THREADS, 1)
uncertainty_cnn_kernel(const float* __restrict__ corr, const uint8_t* __restrict__ params, __nv_bfloat16* __restrict__ out,
                       long npix, long hw, float slope) {
  extern __shared__ __align__(16) uint8_t uc_smem_raw[];
  UcSmem<S>& sm = *reinterpret_cast<UcSmem<S>*>(uc_smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < UC_PARAM_BYTES / 16; i += UC_THREADS)
    reinterpret_cast<uint4*>(sm.params)[i] = __ldg(reinterpret_cast<const uint4*>(params) + i);
  const float* w0 = reinterpret_cast<const float*>(sm.params + UC_OFF_W0);
  const float* b0 = reinterpret_cast<const float*>(sm.params + UC_OFF_B0);
  const float* b1 = reinterpret_cast<const float*>(sm.params + UC_OFF_B1);
  const float* b2 = reinterpret_cast<const float*>(sm.params + UC_OFF_B2);
  const float* w3 = reinterpret_cast<const float*>(sm.params + UC_OFF_W3);
  const float* b3 = reinterpret_cast<const float*>(sm.params + UC_OFF_B3);
  const uint32_t w1_s = (uint32_t)__cvta_generic_to_shared(sm.params + UC_OFF_W1);
  const uint32_t w2_s = (uint32_t)__cvta_generic_to_shared(sm.params + UC_OFF_W2);
  const uint32_t act0_s = (uint32_t)__cvta_generic_to_shared(sm.act0);
  const uint32_t act1_s = (uint32_t)__cvta_generic_to_shared(sm.act1);
  const long ngroups = (npix + UC_PIX - 1) / UC_PIX;

  for (long grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
    const long pix0 = grp * UC_PIX;
    __syncthreads();   // the previous group's readers are done with patch / act2 (and the parameters are staged)
    // ---- gather the 16 patches: plane-major loop, 16 consecutive pixels of a plane are one 64-byte segment
    for (int i = tid; i < S * S * UC_PIX; i += UC_THREADS) {
      const int p = i / UC_PIX, px = i % UC_PIX;
      const long pix = pix0 + px;
      float v = 0.f;
      if (pix < npix) {
        const long b = pix / hw, r = pix - b * hw;
        v = __ldg(corr + (b * (S * S) + p) * hw + r);
      }
      sm.patch[px][p] = v;
    }
    __syncthreads();
    // ---- conv_0 (+ folded BN + LeakyReLU [+ 2x2 max-pool for S = 16]) on the CUDA cores -> act0 bf16 [px][49][32]
    for (int item = tid; item < UC_PIX * 49; item += UC_THREADS) {
      const int px = item / 49, pos = item % 49, oy = pos / 7, ox = pos % 7;
      float acc[32];
      if (S == 9) {
        float v[9];
#pragma unroll
        for (int t = 0; t < 9; ++t) v[t] = sm.patch[px][(oy + t / 3) * 9 + ox + t % 3];
#pragma unroll
        for (int c = 0; c < 32; ++c) acc[c] = b0[c];
#pragma unroll
        for (int t = 0; t < 9; ++t) {
#pragma unroll
          for (int c = 0; c < 32; c += 4) {
            const float4 w = *reinterpret_cast<const float4*>(w0 + t * 32 + c);
            acc[c] = fmaf(v[t], w.x, acc[c]);
            acc[c + 1] = fmaf(v[t], w.y, acc[c + 1]);
            acc[c + 2] = fmaf(v[t], w.z, acc[c + 2]);
            acc[c + 3] = fmaf(v[t], w.w, acc[c + 3]);
          }
        }
#pragma unroll
        for (int c = 0; c < 32; ++c) acc[c] = leaky(acc[c], slope);
      } else {
        // S = 16: the four conv outputs under one pooling window share a 4 x 4 input region
        float v[16];
#pragma unroll
        for (int t = 0; t < 16; ++t) v[t] = sm.patch[px][(2 * oy + t / 4) * 16 + 2 * ox + t % 4];
#pragma unroll
        for (int c = 0; c < 32; ++c) acc[c] = -3.0e38f;
#pragma unroll
        for (int q = 0; q < 4; ++q) {   // (fully unrolled: every index into v[] is a compile-time constant)
          const int qy = q >> 1, qx = q & 1;
#pragma unroll
          for (int c = 0; c < 32; c += 4) {
            float4 a = *reinterpret_cast<const float4*>(b0 + c);
#pragma unroll
            for (int t = 0; t < 9; ++t) {
              const float4 w = *reinterpret_cast<const float4*>(w0 + t * 32 + c);
              const float x = v[(qy + t / 3) * 4 + qx + t % 3];
              a.x = fmaf(x, w.x, a.x);
              a.y = fmaf(x, w.y, a.y);
              a.z = fmaf(x, w.z, a.z);
              a.w = fmaf(x, w.w, a.w);
            }
            acc[c] = fmaxf(acc[c], leaky(a.x, slope));
            acc[c + 1] = fmaxf(acc[c + 1], leaky(a.y, slope));
            acc[c + 2] = fmaxf(acc[c + 2], leaky(a.z, slope));
            acc[c + 3] = fmaxf(acc[c + 3], leaky(a.w, slope));
          }
        }
      }
      uint4* dst = reinterpret_cast<uint4*>(sm.act0 + (size_t)item * UC_APITCH);
#pragma unroll
      for (int q = 0; q < 4; ++q)
        dst[q] = make_uint4(pack_bf16(acc[8 * q], acc[8 * q + 1]), pack_bf16(acc[8 * q + 2], acc[8 * q + 3]),
                            pack_bf16(acc[8 * q + 4], acc[8 * q + 5]), pack_bf16(acc[8 * q + 6], acc[8 * q + 7]));
    }
    __syncthreads();
    // ---- conv_1 on the tensor cores: rows = (pixel, 5 x 5 output position), N = 32, K = 9 taps x 32 channels
    {
      const int mi = lane >> 3, ri = lane & 7;
      for (int mt = warp; mt < UC_PIX * 25 / 16; mt += UC_THREADS / 32) {
        const int row = mt * 16 + (mi & 1) * 8 + ri;            // the A row this lane addresses for ldmatrix
        const int px = row / 25, pos = row % 25, py = pos / 5, pxx = pos % 5;
        const uint32_t a_base = act0_s + (uint32_t)(((px * 49 + py * 7 + pxx) * UC_APITCH + (mi >> 1) * 8) * 2);
        // B: lane addresses filter row n = (mi >> 1) * 8 + ri (+ 16 for the second pair of n-tiles), k half = mi & 1
        const uint32_t b_base = w1_s + (uint32_t)((((mi >> 1) * 8 + ri) * UC_WPITCH + (mi & 1) * 8) * 2);
        float d[4][4];
#pragma unroll
        for (int n = 0; n < 4; ++n)
#pragma unroll
          for (int e = 0; e < 4; ++e) d[n][e] = 0.f;
#pragma unroll 2
        for (int ks = 0; ks < 18; ++ks) {
          const int tap = ks >> 1;
          uint32_t a[4], bA[4], bB[4];
          ldmatrix_x4(a, a_base + (uint32_t)((((tap / 3) * 7 + tap % 3) * UC_APITCH + (ks & 1) * 16) * 2));
          ldmatrix_x4(bA, b_base + (uint32_t)(ks * 32));
          ldmatrix_x4(bB, b_base + (uint32_t)(16 * UC_WPITCH * 2 + ks * 32));
          mma_bf16_16816(d[0], a, bA[0], bA[1]);
          mma_bf16_16816(d[1], a, bA[2], bA[3]);
          mma_bf16_16816(d[2], a, bB[0], bB[1]);
          mma_bf16_16816(d[3], a, bB[2], bB[3]);
        }
        const int g = lane >> 2, t = lane & 3;
#pragma unroll
        for (int n = 0; n < 4; ++n) {
          const int col = n * 8 + 2 * t;
          const float bx = b1[col], by = b1[col + 1];
          uint32_t* r0 = reinterpret_cast<uint32_t*>(sm.act1 + (size_t)(mt * 16 + g) * UC_APITCH + col);
          uint32_t* r1 = reinterpret_cast<uint32_t*>(sm.act1 + (size_t)(mt * 16 + g + 8) * UC_APITCH + col);
          *r0 = pack_bf16(leaky(d[n][0] + bx, slope), leaky(d[n][1] + by, slope));
          *r1 = pack_bf16(leaky(d[n][2] + bx, slope), leaky(d[n][3] + by, slope));
        }
      }
    }
    __syncthreads();
    // ---- conv_2: rows = (pixel, 3 x 3 output position) = 9 m-tiles, N = 16 (two n-tiles -> 18 units over the 8 warps)
    {
      const int mi = lane >> 3, ri = lane & 7;
      for (int u = warp; u < (UC_PIX * 9 / 16) * 2; u += UC_THREADS / 32) {
        const int mt = u >> 1, nt = u & 1;
        const int row = mt * 16 + (mi & 1) * 8 + ri;
        const int px = row / 9, pos = row % 9, py = pos / 3, pxx = pos % 3;
        const uint32_t a_base = act1_s + (uint32_t)(((px * 25 + py * 5 + pxx) * UC_APITCH + (mi >> 1) * 8) * 2);
        const uint32_t b_base = w2_s + (uint32_t)(((nt * 8 + ri) * UC_WPITCH + (mi & 1) * 8) * 2);   // lanes 0-15 matter (x2)
        float d[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 2
        for (int ks = 0; ks < 18; ++ks) {
          const int tap = ks >> 1;
          uint32_t a[4], b[2];
          ldmatrix_x4(a, a_base + (uint32_t)((((tap / 3) * 5 + tap % 3) * UC_APITCH + (ks & 1) * 16) * 2));
          ldmatrix_x2(b, b_base + (uint32_t)(ks * 32));
          mma_bf16_16816(d, a, b[0], b[1]);
        }
        const int g = lane >> 2, t = lane & 3, col = nt * 8 + 2 * t;
        const float bx = b2[col], by = b2[col + 1];
        *reinterpret_cast<uint32_t*>(sm.act2 + (size_t)(mt * 16 + g) * 16 + col) = pack_bf16(leaky(d[0] + bx, slope), leaky(d[1] + by, slope));
        *reinterpret_cast<uint32_t*>(sm.act2 + (size_t)(mt * 16 + g + 8) * 16 + col) = pack_bf16(leaky(d[2] + bx, slope), leaky(d[3] + by, slope));
      }
    }
    __syncthreads();
    // ---- predict_uncertainty: 6 outputs per pixel, K = 9 positions x 16 channels (plain conv: bias, no activation)
    if (tid < UC_PIX * 6) {
      const int px = tid / 6, o = tid % 6;
      const long pix = pix0 + px;
      if (pix < npix) {
        float acc = b3[o];
        const __nv_bfloat162* a = reinterpret_cast<const __nv_bfloat162*>(sm.act2 + px * 144);
        const float* w = w3 + o * 144;
#pragma unroll 8
        for (int k = 0; k < 72; ++k) {
          const float2 v = __bfloat1622float2(a[k]);
          acc = fmaf(v.x, w[2 * k], acc);
          acc = fmaf(v.y, w[2 * k + 1], acc);
        }
        out[pix * 6 + o] = __float2bfloat16(acc);
      }
    }
  }
}

template <int S>
static int uc_launch(const float* corr, const void* params, void* out, long npix, long hw, float slope, cudaStream_t st) {
  static bool attr_set = false;
  const int smem = (int)sizeof(UcSmem<S>);
  if (!attr_set) {
    RF_CUDA(cudaFuncSetAttribute(uncertainty_cnn_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  const long ngroups = (npix + UC_PIX - 1) / UC_PIX;
  const int blocks = (int)(ngroups < kNumSMs ? ngroups : kNumSMs);
  uncertainty_cnn_kernel<S><<<blocks, UC_THREADS, smem, st>>>(corr, (const uint8_t*)params, (__nv_bfloat16*)out, npix, hw, slope);
  RF_CHECK_LAUNCH("uncertainty_cnn_kernel");
  return RF_OK;
}

}  // namespace rf

using namespace rf;

extern "C" int rf_uncertainty_cnn_param_bytes(void) { return UC_PARAM_BYTES; }

extern "C" int rf_uncertainty_cnn_fwd(const float* corr, const void* params, void* out_bf16, int B, int H, int W, int search_size,
                                      float slope, void* stream) {
  RF_REQUIRE(corr && params && out_bf16 && B > 0 && H > 0 && W > 0, "rf_uncertainty_cnn_fwd: bad arguments");
  RF_REQUIRE(search_size == 9 || search_size == 16, "rf_uncertainty_cnn_fwd: search_size must be 9 or 16 (got %d)", search_size);
  RF_REQUIRE(((uintptr_t)params & 15) == 0, "rf_uncertainty_cnn_fwd: the parameter block must be 16-byte aligned");
  const long hw = (long)H * W, npix = (long)B * hw;
  cudaStream_t st = (cudaStream_t)stream;
  return search_size == 9 ? uc_launch<9>(corr, params, out_bf16, npix, hw, slope, st)
                          : uc_launch<16>(corr, params, out_bf16, npix, hw, slope, st);
}
