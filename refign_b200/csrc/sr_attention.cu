// Fused spatial-reduction attention core for the MiT encoder on sm_100a (tcgen05 + TMEM + TMA):
//   O[b, n, h, :] = softmax_m( scale * <Q[b, n, h, :], K[b, m, h, :]> ) V[b, m, h, :],  head_dim = 64
// Restates the attention core of /root/reference/models/backbones/mix_transformer.py:150-160
//   attn = (q @ k.transpose(-2, -1)) * self.scale; attn = attn.softmax(dim=-1); x = (attn @ v)
// without materialising the [B, heads, N, N_kv] matrix (268 MB per image per block in fp32 at
// 1024x1024 stage 1) and without the head permute / contiguous copies: Q is read in place from the
// q-projection output [B, N, heads*64] and K / V from the kv-projection output [B, M, 2*heads*64]
// (k = channels [0, C), v = channels [C, 2C), as the reference's reshape(B,-1,2,h,d) implies).
//
// One CTA = one 128-query tile of one (batch, head); 128 threads, thread t owns query row t (TMEM lane t).
//   TMA      : Q tile [128 x 64] once; K / V chunks [128 x 64] double-buffered, 128-byte swizzle
//   tcgen05  : S = Q K^T      (M128 N128 K16 x4, both operands K-major from shared memory, fp32 in TMEM)
//              O += P V       (M128 N64  K16 x8, A = P as packed bf16 in TMEM, B = V MN-major from smem)
//   softmax  : online (running max / sum per row) in registers, exp2 with the scale folded in;
//              O is rescaled in TMEM when the running max moves
// TMEM columns: S [0,128) | O [128,192) | P [192,256)  -> 256 columns, two CTAs per SM.
// The forward also emits LSE[b,h,n] = log sum_m exp(scale * s) for the backward.
//
// Bound: MUFU (exp2) -- 256 tensor flops per exponential at head_dim 64; reported against the bf16
// tensor peak, see DESIGN.md.
#include <cuda_bf16.h>

#include "rf_common.cuh"
#include "rf_sm100.cuh"

namespace rf {
using namespace sm100;

constexpr int AT_BM = 128;   // queries per CTA
constexpr int AT_BN = 128;   // keys per chunk
constexpr int AT_D = 64;     // head dim
constexpr int AT_TILE_BYTES = AT_BM * AT_D * 2;  // 16 KiB, one [128 x 64] bf16 tile
constexpr int AT_SMEM = 5 * AT_TILE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;

struct __align__(8) AttnBars {
  uint64_t q, kv[2], s, o;
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(128, 2)
sr_attention_fwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
                        __nv_bfloat16* __restrict__ out, float* __restrict__ lse, int N, int M, int heads,
                        float scale_log2) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;
  uint8_t* sK = smem + AT_TILE_BYTES;            // 2 stages
  uint8_t* sV = smem + 3 * AT_TILE_BYTES;        // 2 stages
  AttnBars* bars = reinterpret_cast<AttnBars*>(smem + 5 * AT_TILE_BYTES);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int q0 = blockIdx.x * AT_BM, head = blockIdx.y, b = blockIdx.z;
  const int C = heads * AT_D;
  const int nchunks = (M + AT_BN - 1) / AT_BN;

  if (tid == 0) {
    mbar_init(&bars->q, 1);
    mbar_init(&bars->kv[0], 1);
    mbar_init(&bars->kv[1], 1);
    mbar_init(&bars->s, 1);
    mbar_init(&bars->o, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<256>(&bars->tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;
  const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
  const uint32_t tS = tmem + lane_off, tO = tmem + lane_off + 128, tP = tmem + lane_off + 192;

  if (tid == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_kv);
    mbar_expect_tx(&bars->q, AT_TILE_BYTES);
    tma_load_3d(sQ, &tm_q, &bars->q, head * AT_D, q0, b);
    for (int s = 0; s < 2 && s < nchunks; ++s) {
      mbar_expect_tx(&bars->kv[s], 2 * AT_TILE_BYTES);
      tma_load_3d(sK + s * AT_TILE_BYTES, &tm_kv, &bars->kv[s], head * AT_D, s * AT_BN, b);
      tma_load_3d(sV + s * AT_TILE_BYTES, &tm_kv, &bars->kv[s], C + head * AT_D, s * AT_BN, b);
    }
  }
  __syncwarp();

  constexpr uint32_t IDESC_S = make_idesc(FMT_BF16, 128, 128, 0, 0);
  constexpr uint32_t IDESC_O = make_idesc(FMT_BF16, 128, 64, 0, 1);
  const uint64_t descQ = make_sdesc_sw128(smem_u32(sQ), 16, 1024);

  float m_run = -INFINITY, l_run = 0.f;

  for (int j = 0; j < nchunks; ++j) {
    const int st = j & 1;
    if (tid == 0) {
      if (j == 0) mbar_wait(&bars->q, 0);
      mbar_wait(&bars->kv[st], (j >> 1) & 1);
      tc_fence_after();
      const uint64_t descK = make_sdesc_sw128(smem_u32(sK + st * AT_TILE_BYTES), 16, 1024);
#pragma unroll
      for (int k = 0; k < AT_D / 16; ++k)
        mma_f16_ss(tmem, descQ + (uint64_t)(k * 2), descK + (uint64_t)(k * 2), IDESC_S, k > 0 ? 1u : 0u);
      tc_commit(&bars->s);
    }
    __syncwarp();
    mbar_wait(&bars->s, j & 1);
    tc_fence_after();

    const int kvalid = M - j * AT_BN;  // columns >= kvalid are padding (zero-filled by TMA)
    // pass 1: row maximum of the raw scores
    float mx = m_run;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint32_t v[32];
      tmem_ld32(tS + c * 32, v);
      tc_wait_ld();
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (c * 32 + i < kvalid) mx = fmaxf(mx, __uint_as_float(v[i]));
    }
    const float alpha = exp2f((m_run - mx) * scale_log2);  // 0 on the first chunk (m_run = -inf)
    const float mneg = -mx * scale_log2;
    l_run *= alpha;
    if (j > 0) {
      // O += P V of the previous chunk must have landed before O is rescaled and P overwritten
      mbar_wait(&bars->o, (j - 1) & 1);
      tc_fence_after();
      if (tid == 0 && j + 1 < nchunks) {  // its K / V stage is free again: prefetch chunk j+1
        const int s2 = (j + 1) & 1;
        mbar_expect_tx(&bars->kv[s2], 2 * AT_TILE_BYTES);
        tma_load_3d(sK + s2 * AT_TILE_BYTES, &tm_kv, &bars->kv[s2], head * AT_D, (j + 1) * AT_BN, b);
        tma_load_3d(sV + s2 * AT_TILE_BYTES, &tm_kv, &bars->kv[s2], C + head * AT_D, (j + 1) * AT_BN, b);
      }
      __syncwarp();
      if (__any_sync(0xffffffffu, alpha != 1.f)) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t v[32];
          tmem_ld32(tO + c * 32, v);
          tc_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
          tmem_st32(tO + c * 32, v);
        }
      }
    }
    // pass 2: P = exp2(scale * s - scale * max), row sum, bf16 pack into TMEM
    float rs = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint32_t v[32];
      tmem_ld32(tS + c * 32, v);
      tc_wait_ld();
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        float p0 = exp2f(fmaf(__uint_as_float(v[2 * i]), scale_log2, mneg));
        float p1 = exp2f(fmaf(__uint_as_float(v[2 * i + 1]), scale_log2, mneg));
        if (c * 32 + 2 * i >= kvalid) p0 = 0.f;
        if (c * 32 + 2 * i + 1 >= kvalid) p1 = 0.f;
        rs += p0 + p1;
        const __nv_bfloat162 h = __floats2bfloat162_rn(p0, p1);
        pk[i] = *reinterpret_cast<const uint32_t*>(&h);
      }
      tmem_st16(tP + c * 16, pk);
    }
    l_run += rs;
    m_run = mx;
    tc_wait_st();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint64_t descV = make_sdesc_sw128(smem_u32(sV + st * AT_TILE_BYTES), 8192, 1024);
#pragma unroll
      for (int k = 0; k < AT_BN / 16; ++k)
        mma_f16_ts(tmem + 128, tmem + 192 + k * 8, descV + (uint64_t)(k * (16 * 128 / 16)), IDESC_O,
                   (j > 0 || k > 0) ? 1u : 0u);
      tc_commit(&bars->o);
    }
    __syncwarp();
  }

  mbar_wait(&bars->o, (nchunks - 1) & 1);
  tc_fence_after();
  const int row = q0 + tid;
  const float inv_l = 1.f / l_run;
  uint32_t packed[32];
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    uint32_t v[32];
    tmem_ld32(tO + c * 32, v);
    tc_wait_ld();
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const __nv_bfloat162 h =
          __floats2bfloat162_rn(__uint_as_float(v[2 * i]) * inv_l, __uint_as_float(v[2 * i + 1]) * inv_l);
      packed[c * 16 + i] = *reinterpret_cast<const uint32_t*>(&h);
    }
  }
  if (row < N) {
    uint4* dst = reinterpret_cast<uint4*>(out + ((long)b * N + row) * C + head * AT_D);
#pragma unroll
    for (int i = 0; i < 8; ++i) dst[i] = make_uint4(packed[4 * i], packed[4 * i + 1], packed[4 * i + 2], packed[4 * i + 3]);
    if (lse != nullptr)
      lse[((long)b * heads + head) * N + row] = m_run * scale_log2 * 0.69314718055994531f + __logf(l_run);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tmem);
}

}  // namespace rf

using namespace rf;

extern "C" int rf_sr_attention_fwd(const void* q, const void* kv, void* out, float* lse, int B, int N, int M,
                                   int heads, float scale, void* stream) {
  RF_REQUIRE(q && kv && out, "rf_sr_attention_fwd: null pointer");
  RF_REQUIRE(B > 0 && N > 0 && M > 0 && heads > 0 && B <= 65535 && heads <= 65535, "rf_sr_attention_fwd: bad shape");
  RF_REQUIRE(((uintptr_t)out & 15) == 0, "rf_sr_attention_fwd: out must be 16-byte aligned");
  const int C = heads * AT_D;
  CUtensorMap tq, tkv;
  int rc = make_tmap_3d(&tq, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, q, (uint64_t)C, (uint64_t)N, (uint64_t)B,
                        (uint64_t)C * 2, (uint64_t)N * C * 2, AT_D, AT_BM);
  if (rc != RF_OK) return rc;
  rc = make_tmap_3d(&tkv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, kv, (uint64_t)2 * C, (uint64_t)M, (uint64_t)B,
                    (uint64_t)2 * C * 2, (uint64_t)M * 2 * C * 2, AT_D, AT_BN);
  if (rc != RF_OK) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    RF_CUDA(cudaFuncSetAttribute(sr_attention_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM));
    attr_set = true;
  }
  dim3 grid((unsigned)((N + AT_BM - 1) / AT_BM), (unsigned)heads, (unsigned)B);
  const float scale_log2 = scale * 1.44269504088896341f;
  sr_attention_fwd_kernel<<<grid, 128, AT_SMEM, (cudaStream_t)stream>>>(tq, tkv, (__nv_bfloat16*)out, lse, N, M, heads,
                                                                          scale_log2);
  RF_CHECK_LAUNCH("sr_attention_fwd_kernel");
  return RF_OK;
}
