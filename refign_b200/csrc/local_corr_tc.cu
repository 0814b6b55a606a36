// Local (windowed) correlation on the tensor cores -- the 9 x 9 volume of LocalFeatureCorrelationLayer
// (reference models/modules.py:266-274 -> correlation_sampler.cpp:62-90) as a BANDED GEMM:
//   out[b, dy, dx, y, x] = sum_c in1[b, c, y, x] * in2[b, c, y + dy - 4, x + dx - 4]        (zero outside the image)
// The FFMA kernel of local_corr.cu is bound by shared-memory operand reads (10 LDS.128 per 108 FMAs: 23 TFLOP/s fp32,
// 0.18 of the HBM roof).  Here one CTA owns a 4 x 32 tile of target pixels (= the 128 TMEM lanes) and multiplies it with
// the 12 x 40 halo of source pixels around it (N = 480 accumulator columns, K = C) with tcgen05.mma; 81 of the 480
// products per pixel are the wanted band, the rest is the price of a dense instruction (the tensor pipe has the room:
// 3 passes x 5.3 k cycles per tile against the 150 us the FFMA kernel takes for the whole volume).
// Precision: fp32 operands are split on the fly into bf16 hi + lo (x = hi + lo + O(2^-17 |x|)) and three bf16 MMAs
// accumulate hi*hi + hi*lo + lo*hi in fp32 -- the dropped lo*lo term and the split error are ~2^-16 relative to
// sum_c |a_c||b_c|, i.e. <= 2e-5 absolute on unit-norm features (a single TF32 pass would be 1e-3).
// Roles (384 threads): warp 6 streams raw fp32 boxes (16 channels x tile / halo rows; TMA zero-fills everything outside
// the image = the correlation's padding) into two 38 KiB buffers; warps 0-5 split them into bf16 hi / lo and write the
// UMMA 128-byte-swizzled MN-major operand layout (pixels are the contiguous dimension of NCHW; three 40 KiB stages);
// warp 7 issues the MMAs; warps 8-11 pull the band out of TMEM: warp = tile row, lane = tile column, so the column offset (dy, dx) is
// warp-uniform except for "+ lane", which a per-lane shared-memory bounce row resolves; optional fused ReLU + L2-norm
// over the 81 displacements (modules.py:272-273) happens on the 81 registers before the coalesced plane stores.
#include <cuda_bf16.h>

#include "rf_common.cuh"
#include "rf_sm100.cuh"
#include "rf_trace.cuh"

namespace rf {
using namespace sm100;

constexpr int LT_TH = 4, LT_TW = 32;                    // target tile
constexpr int LT_P = 9, LT_R = 4;                       // patch, radius
constexpr int LT_HH = LT_TH + 2 * LT_R, LT_HW = LT_TW + 2 * LT_R;   // 12 x 40 halo
constexpr int LT_N = LT_HH * LT_HW;                     // 480 accumulator columns
constexpr int LT_KC = 16;                               // channels per stage (= one K16 MMA step)
constexpr int LT_BLK = LT_KC * 128;                     // one MN-major block: 16 k rows x 64 pixels x 2 B
constexpr int LT_A_BYTES = 2 * LT_BLK;                  // 128 target pixels
constexpr int LT_B_BYTES = 8 * LT_BLK;                  // 480 halo pixels -> 8 blocks (the last half used)
constexpr int LT_STAGE = 2 * (LT_A_BYTES + LT_B_BYTES); // hi + lo: 40 KiB
#ifdef WS_TRACE
constexpr int LT_STAGES = 2;                            // (the trace log takes 22 KiB of static smem)
#else
constexpr int LT_STAGES = 3;                            // bf16 operand stages
#endif
constexpr int LT_RAW_A = LT_KC * LT_TH * LT_TW * 4;     // 8 KiB   fp32 [16][4][32]
constexpr int LT_RAW_B = LT_KC * LT_HH * LT_HW * 4;     // 30 KiB  fp32 [16][12][40]
constexpr int LT_RAW = LT_RAW_A + LT_RAW_B;
constexpr int LT_RAWS = 2;
constexpr int LT_BOUNCE_PITCH = 44;                     // words per lane row (40 used; 16-byte aligned, conflict-free)
constexpr int LT_BOUNCE_BYTES = 4 * 32 * LT_BOUNCE_PITCH * 4;
constexpr int LT_SMEM = LT_STAGES * LT_STAGE + LT_RAWS * LT_RAW + LT_BOUNCE_BYTES + 256 + 1024;
constexpr int LT_CONV_WARPS = 6;                        // 12 warps in all: registers are allocated per 4 warps
constexpr int LT_CONV = LT_CONV_WARPS * 32;
constexpr int LT_THREADS = LT_CONV + 64 + 128;          // + TMA warp + MMA warp + 4 epilogue warps
constexpr int LT_W_TMA = LT_CONV_WARPS, LT_W_MMA = LT_CONV_WARPS + 1;

struct LtBars {
  uint64_t raw_full[2], raw_empty[2];
  uint64_t full[LT_STAGES], empty[LT_STAGES];
  uint64_t acc_full, acc_empty;
  uint32_t tmem_base;
};

// byte offset of pixel m (multiple of 4), channel k inside an MN-major 128-byte-swizzled operand of 16-channel blocks
__device__ __forceinline__ uint32_t lt_off(int m, int k) {
  return (uint32_t)((m >> 6) * LT_BLK + k * 128 + ((((m & 63) >> 3) ^ (k & 7)) << 4) + ((m & 7) << 1));
}

template <bool FUSE>
__global__ void __launch_bounds__(LT_THREADS, 1)
local_corr_tc_kernel(const __grid_constant__ CUtensorMap tm1, const __grid_constant__ CUtensorMap tm2, float* __restrict__ out,
                     float* __restrict__ norm_out, int B, int C, int H, int W, int tiles_y, int tiles_x) {
  WS_T_INIT();
  WS_T(0);
  extern __shared__ uint8_t lt_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(lt_smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* raw = smem + LT_STAGES * LT_STAGE;
  float* bounce = reinterpret_cast<float*>(raw + LT_RAWS * LT_RAW);
  LtBars* bars = reinterpret_cast<LtBars*>(raw + LT_RAWS * LT_RAW + LT_BOUNCE_BYTES);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long plane = (long)H * W;
  const int ntiles = B * tiles_y * tiles_x, kchunks = C / LT_KC;

  if (warp == LT_W_MMA) {
    if (elect_one()) {
      tma_prefetch_desc(&tm1);
      tma_prefetch_desc(&tm2);
      for (int s = 0; s < LT_RAWS; ++s) {
        mbar_init(&bars->raw_full[s], 1);
        mbar_init(&bars->raw_empty[s], LT_CONV_WARPS);
      }
      for (int s = 0; s < LT_STAGES; ++s) {
        mbar_init(&bars->full[s], LT_CONV_WARPS);
        mbar_init(&bars->empty[s], 1);
      }
      mbar_init(&bars->acc_full, 1);
      mbar_init(&bars->acc_empty, 4);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<512>(&bars->tmem_base);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = uniform_u32(bars->tmem_base);

  if (warp < LT_CONV_WARPS) {
    // ------------------------------------------------------------------ converters: raw fp32 -> bf16 hi / lo, UMMA layout
    constexpr int NA = LT_KC * LT_TH * (LT_TW / 4), NB = LT_KC * LT_HH * (LT_HW / 4);   // 512 + 1920 float4 per stage
    constexpr int PER = (NA + NB + LT_CONV - 1) / LT_CONV;                                // 13 per thread
    // this thread's float4 slots of a stage: offset in the raw buffer and destination in the operand stage -- the same
    // for every stage of every tile (the tile origin only enters the TMA coordinates)
    uint32_t src[PER], dst[PER];
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int j = tid + i * LT_CONV;
      src[i] = 0;
      dst[i] = 0xffffffffu;
      if (j < NA) {
        const int ch = j / (LT_TH * (LT_TW / 4)), rem = j % (LT_TH * (LT_TW / 4)), ty = rem / (LT_TW / 4), q = rem % (LT_TW / 4);
        src[i] = (uint32_t)(((ch * LT_TH + ty) * LT_TW + 4 * q) * 4);
        dst[i] = lt_off(ty * LT_TW + 4 * q, ch);
      } else if (j < NA + NB) {
        const int jj = j - NA;
        const int ch = jj / (LT_HH * (LT_HW / 4)), rem = jj % (LT_HH * (LT_HW / 4)), hy = rem / (LT_HW / 4), q = rem % (LT_HW / 4);
        src[i] = (uint32_t)(LT_RAW_A + ((ch * LT_HH + hy) * LT_HW + 4 * q) * 4);
        dst[i] = 2 * LT_A_BYTES + lt_off(hy * LT_HW + 4 * q, ch);
      }
    }
    int it = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
      for (int kc = 0; kc < kchunks; ++kc, ++it) {
        const int st = it % LT_STAGES, rs = it % LT_RAWS;
        mbar_wait(&bars->raw_full[rs], (it / LT_RAWS) & 1);
        WS_T(31);
        if (it >= LT_STAGES) mbar_wait(&bars->empty[st], ((it / LT_STAGES) - 1) & 1);
        WS_T(32);
        // explicit shared-space accesses (plain dereferences of the manually aligned window compile to generic LD / ST)
        const uint32_t rbuf = smem_u32(raw) + rs * LT_RAW;
        const uint32_t stage = smem_u32(smem) + st * LT_STAGE;
#pragma unroll
        for (int i = 0; i < PER; ++i) {
          if (dst[i] != 0xffffffffu) {
            const float4 v = lds_f4(rbuf + src[i]);
            const __nv_bfloat162 h01 = __floats2bfloat162_rn(v.x, v.y), h23 = __floats2bfloat162_rn(v.z, v.w);
            const float2 f01 = __bfloat1622float2(h01), f23 = __bfloat1622float2(h23);
            const __nv_bfloat162 l01 = __floats2bfloat162_rn(v.x - f01.x, v.y - f01.y);
            const __nv_bfloat162 l23 = __floats2bfloat162_rn(v.z - f23.x, v.w - f23.y);
            const uint32_t lo_off = dst[i] < 2 * LT_A_BYTES ? LT_A_BYTES : LT_B_BYTES;   // [A hi][A lo][B hi][B lo]
            sts_u2(stage + dst[i], *reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
            sts_u2(stage + dst[i] + lo_off, *reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
          }
        }
        fence_proxy_async();          // generic-proxy writes -> visible to the tensor core's async-proxy reads
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&bars->full[st]);
          mbar_arrive(&bars->raw_empty[rs]);      // this warp's reads of the raw buffer are done
        }
        WS_T(30);
      }
    }
  } else if (warp == LT_W_TMA) {
    // ------------------------------------------------------------------ TMA producer of the raw fp32 boxes
    int it = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
      const int b = t / (tiles_y * tiles_x), y0 = ((t / tiles_x) % tiles_y) * LT_TH, x0 = (t % tiles_x) * LT_TW;
      for (int kc = 0; kc < kchunks; ++kc, ++it) {
        const int rs = it % LT_RAWS;
        if (it >= LT_RAWS) mbar_wait(&bars->raw_empty[rs], ((it / LT_RAWS) - 1) & 1);
        if (elect_one()) {
          uint8_t* rbuf = raw + rs * LT_RAW;
          mbar_expect_tx(&bars->raw_full[rs], LT_RAW);
          tma_load_4d(rbuf, &tm1, &bars->raw_full[rs], x0, y0, kc * LT_KC, b);
          tma_load_4d(rbuf + LT_RAW_A, &tm2, &bars->raw_full[rs], x0 - LT_R, y0 - LT_R, kc * LT_KC, b);
        }
        __syncwarp();
        WS_T(40);
      }
    }
  } else if (warp == LT_W_MMA) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t ID0 = make_idesc(FMT_BF16, 128, 256, 1, 1), ID1 = make_idesc(FMT_BF16, 128, LT_N - 256, 1, 1);
    const uint32_t base = smem_u32(smem);
    int it = 0, local = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++local) {
      if (local > 0) mbar_wait(&bars->acc_empty, (local - 1) & 1);
      tc_fence_after();
      for (int kc = 0; kc < kchunks; ++kc, ++it) {
        const int st = it % LT_STAGES;
        mbar_wait(&bars->full[st], (it / LT_STAGES) & 1);
        tc_fence_after();
        WS_T(20);
        const uint32_t sa = base + st * LT_STAGE;
        const uint64_t a_hi = make_sdesc_sw128(sa, LT_BLK, 1024), a_lo = make_sdesc_sw128(sa + LT_A_BYTES, LT_BLK, 1024);
        const uint32_t sb = sa + 2 * LT_A_BYTES;
        const uint64_t b_hi0 = make_sdesc_sw128(sb, LT_BLK, 1024), b_hi1 = make_sdesc_sw128(sb + 4 * LT_BLK, LT_BLK, 1024);
        const uint64_t b_lo0 = make_sdesc_sw128(sb + LT_B_BYTES, LT_BLK, 1024), b_lo1 = make_sdesc_sw128(sb + LT_B_BYTES + 4 * LT_BLK, LT_BLK, 1024);
        if (elect_one()) {
          const uint32_t acc = kc > 0 ? 1u : 0u;
          mma_f16_ss(tmem, a_hi, b_hi0, ID0, acc);
          mma_f16_ss(tmem + 256, a_hi, b_hi1, ID1, acc);
          mma_f16_ss(tmem, a_hi, b_lo0, ID0, 1u);
          mma_f16_ss(tmem + 256, a_hi, b_lo1, ID1, 1u);
          mma_f16_ss(tmem, a_lo, b_hi0, ID0, 1u);
          mma_f16_ss(tmem + 256, a_lo, b_hi1, ID1, 1u);
          tc_commit(&bars->empty[st]);
          if (kc == kchunks - 1) tc_commit(&bars->acc_full);
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------------ band extraction (warp = tile row, lane = tile column)
    const int q = warp & 3;                        // TMEM lane quarter this warp may access = its tile row
    const uint32_t tbase = tmem + ((uint32_t)(q * 32) << 16);
    const uint32_t brow = smem_u32(bounce) + (q * 32 + lane) * LT_BOUNCE_PITCH * 4;   // this lane's bounce row (shared address)
    int local = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++local) {
      const int b = t / (tiles_y * tiles_x), y0 = ((t / tiles_x) % tiles_y) * LT_TH, x0 = (t % tiles_x) * LT_TW;
      mbar_wait(&bars->acc_full, local & 1);
      tc_fence_after();
      WS_T(10);
      float v[LT_P * LT_P];
      // halo row q + dy of the accumulator (40 columns) -> this lane's bounce row -> the 9 columns lane .. lane + 8.
      // The TMEM load of the next row is in flight while this one goes through shared memory.
      uint32_t h0[32], h1[8];
      tmem_ld32(tbase + q * LT_HW, h0);
      tmem_ld8(tbase + q * LT_HW + 32, h1);
#pragma unroll
      for (int dy = 0; dy < LT_P; ++dy) {
        tc_wait_ld();
#pragma unroll
        for (int c = 0; c < 32; c += 4) sts_u4(brow + 4 * c, h0[c], h0[c + 1], h0[c + 2], h0[c + 3]);
        sts_u4(brow + 128, h1[0], h1[1], h1[2], h1[3]);
        sts_u4(brow + 144, h1[4], h1[5], h1[6], h1[7]);
        if (dy + 1 < LT_P) {
          tmem_ld32(tbase + (q + dy + 1) * LT_HW, h0);
          tmem_ld8(tbase + (q + dy + 1) * LT_HW + 32, h1);
        }
        __syncwarp();                               // (a lane only reads its own row: ordering within the thread suffices,
#pragma unroll                                      //  the barrier keeps the compiler from reordering the shared accesses)
        for (int dx = 0; dx < LT_P; ++dx) v[dy * LT_P + dx] = lds_f1(brow + 4 * (lane + dx));
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->acc_empty);   // the accumulator may be overwritten by the next tile's MMAs
      WS_T(11);
      const int y = y0 + q, x = x0 + lane;
      float inv = 1.f;
      if (FUSE) {
        float ss = 0.f;
#pragma unroll
        for (int d = 0; d < LT_P * LT_P; ++d) {
          v[d] = fmaxf(v[d], 0.f);
          ss = fmaf(v[d], v[d], ss);
        }
        const float nrm = fmaxf(sqrtf(ss), 1e-12f);
        inv = 1.f / nrm;        // one division per pixel; 81 multiplications (<= 1 ulp from v / nrm)
        if (norm_out != nullptr && y < H && x < W) norm_out[(long)b * plane + (long)y * W + x] = nrm;
      }
      if (y < H && x < W) {
        float* o = out + (long)b * LT_P * LT_P * plane + (long)y * W + x;
#pragma unroll
        for (int d = 0; d < LT_P * LT_P; ++d) o[(long)d * plane] = FUSE ? v[d] * inv : v[d];
      }
      WS_T(12);
    }
  }
  tc_fence_before();
  __syncthreads();
  WS_T(2);
  WS_T_FLUSH();
  if (warp == LT_W_MMA) tmem_dealloc<512>(tmem);
}

// 9 x 9 patch, kernel 1, stride 1, no padding / dilation: in1, in2 f32 [B, C, H, W] -> out f32 [B, 81, H, W]
int local_corr_tc_launch(const float* in1, const float* in2, float* out, float* norm_out, int B, int C, int H, int W, bool fuse,
                         cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    RF_CUDA(cudaFuncSetAttribute(local_corr_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, LT_SMEM));
    RF_CUDA(cudaFuncSetAttribute(local_corr_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, LT_SMEM));
    attr = true;
  }
  const int tiles_y = (H + LT_TH - 1) / LT_TH, tiles_x = (W + LT_TW - 1) / LT_TW;
  const long ntiles = (long)B * tiles_y * tiles_x;
  const int grid = (int)(ntiles < kNumSMs ? ntiles : kNumSMs);
  // raw fp32 boxes [x, y, channel, image]; coordinates outside the image are zero-filled by TMA
  CUtensorMap tm1, tm2;
  const uint64_t dims[4] = {(uint64_t)W, (uint64_t)H, (uint64_t)C, (uint64_t)B};
  const uint64_t strides[3] = {(uint64_t)W * 4, (uint64_t)H * W * 4, (uint64_t)C * H * W * 4};
  const uint32_t box1[4] = {LT_TW, LT_TH, LT_KC, 1}, box2[4] = {LT_HW, LT_HH, LT_KC, 1};
  int rc = make_tmap_nd(&tm1, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, in1, 4, dims, strides, box1, CU_TENSOR_MAP_SWIZZLE_NONE);
  if (rc != RF_OK) return rc;
  rc = make_tmap_nd(&tm2, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, in2, 4, dims, strides, box2, CU_TENSOR_MAP_SWIZZLE_NONE);
  if (rc != RF_OK) return rc;
  if (fuse)
    local_corr_tc_kernel<true><<<grid, LT_THREADS, LT_SMEM, st>>>(tm1, tm2, out, norm_out, B, C, H, W, tiles_y, tiles_x);
  else
    local_corr_tc_kernel<false><<<grid, LT_THREADS, LT_SMEM, st>>>(tm1, tm2, out, norm_out, B, C, H, W, tiles_y, tiles_x);
  RF_CHECK_LAUNCH("local_corr_tc_kernel");
  return RF_OK;
}

bool local_corr_tc_ok(int C, int H, int W) { return C % LT_KC == 0 && C >= LT_KC && W % 4 == 0 && (long)H * W >= 64 * 64; }

}  // namespace rf
