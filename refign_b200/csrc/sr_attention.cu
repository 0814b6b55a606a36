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
#include <stdlib.h>

#include "rf_common.cuh"
#include "rf_sm100.cuh"
#include "rf_trace.cuh"

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
    const float alpha = fast_exp2((m_run - mx) * scale_log2);  // 0 on the first chunk (m_run = -inf)
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
        float p0 = fast_exp2(fmaf(__uint_as_float(v[2 * i]), scale_log2, mneg));
        float p1 = fast_exp2(fmaf(__uint_as_float(v[2 * i + 1]), scale_log2, mneg));
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

// Row maximum of this lane's 128 fp32 scores in TMEM (four independent partial maxima for ILP).
template <bool MASK>
__device__ __forceinline__ float attn_row_max(uint32_t tS, int kvalid) {
  float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint32_t v[32];
    tmem_ld32(tS + c * 32, v);
    tc_wait_ld();
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      if (!MASK || c * 32 + i + 0 < kvalid) m0 = fmaxf(m0, __uint_as_float(v[i + 0]));
      if (!MASK || c * 32 + i + 1 < kvalid) m1 = fmaxf(m1, __uint_as_float(v[i + 1]));
      if (!MASK || c * 32 + i + 2 < kvalid) m2 = fmaxf(m2, __uint_as_float(v[i + 2]));
      if (!MASK || c * 32 + i + 3 < kvalid) m3 = fmaxf(m3, __uint_as_float(v[i + 3]));
    }
  }
  return fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
}

// P = exp2(s * scale_log2 + mneg) for this lane's 128 scores, packed to bf16 into TMEM; returns the row sum.
template <bool MASK>
__device__ __forceinline__ float attn_exp_pack(uint32_t tS, uint32_t tP, float scale_log2, float mneg, int kvalid) {
  float r0 = 0.f, r1 = 0.f, r2 = 0.f, r3 = 0.f;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint32_t v[32];
    tmem_ld32(tS + c * 32, v);
    tc_wait_ld();
    uint32_t pk[16];
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
      float p0 = fast_exp2(fmaf(__uint_as_float(v[2 * i + 0]), scale_log2, mneg));
      float p1 = fast_exp2(fmaf(__uint_as_float(v[2 * i + 1]), scale_log2, mneg));
      float p2 = fast_exp2(fmaf(__uint_as_float(v[2 * i + 2]), scale_log2, mneg));
      float p3 = fast_exp2(fmaf(__uint_as_float(v[2 * i + 3]), scale_log2, mneg));
      if (MASK) {
        if (c * 32 + 2 * i + 0 >= kvalid) p0 = 0.f;
        if (c * 32 + 2 * i + 1 >= kvalid) p1 = 0.f;
        if (c * 32 + 2 * i + 2 >= kvalid) p2 = 0.f;
        if (c * 32 + 2 * i + 3 >= kvalid) p3 = 0.f;
      }
      r0 += p0;
      r1 += p1;
      r2 += p2;
      r3 += p3;
      const __nv_bfloat162 h0 = __floats2bfloat162_rn(p0, p1), h1 = __floats2bfloat162_rn(p2, p3);
      pk[i] = *reinterpret_cast<const uint32_t*>(&h0);
      pk[i + 1] = *reinterpret_cast<const uint32_t*>(&h1);
    }
    tmem_st16(tP + c * 16, pk);
  }
  return (r0 + r1) + (r2 + r3);
}

// Single pass over this lane's 128 scores: TMEM read bandwidth (64 B/clk/SM) is the scarcest resource
// of a head_dim-64 softmax, so S is read ONCE into registers; row maximum (partial maxima for ILP),
// lazy-rescale decision by the caller through `decide`, then P = exp2(s * scale_log2 - m * scale_log2)
// packed to bf16 into TMEM.  Returns the row sum; `mx_out` receives the chunk maximum.
struct AttnRow {
  uint32_t v[4][32];
};
template <bool MASK>
__device__ __forceinline__ float attn_load_max(uint32_t tS, int kvalid, AttnRow& row) {
#pragma unroll
  for (int c = 0; c < 4; ++c) tmem_ld32(tS + c * 32, row.v[c]);
  tc_wait_ld();
  float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      if (!MASK || c * 32 + i + 0 < kvalid) m0 = fmaxf(m0, __uint_as_float(row.v[c][i + 0]));
      if (!MASK || c * 32 + i + 1 < kvalid) m1 = fmaxf(m1, __uint_as_float(row.v[c][i + 1]));
      if (!MASK || c * 32 + i + 2 < kvalid) m2 = fmaxf(m2, __uint_as_float(row.v[c][i + 2]));
      if (!MASK || c * 32 + i + 3 < kvalid) m3 = fmaxf(m3, __uint_as_float(row.v[c][i + 3]));
    }
  return fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
}
template <bool MASK>
__device__ __forceinline__ float attn_exp_pack_regs(const AttnRow& row, uint32_t tP, float scale_log2, float mneg,
                                                    int kvalid) {
  // packed fp32 pipe (FFMA2 / FADD2): 5 issue slots per PAIR of scores around its two MUFU.EX2
  const float2 sl = make_float2(scale_log2, scale_log2), mn = make_float2(mneg, mneg);
  float2 r01 = make_float2(0.f, 0.f), r23 = make_float2(0.f, 0.f);
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint32_t pk[16];
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
      const float2 a = __ffma2_rn(make_float2(__uint_as_float(row.v[c][2 * i + 0]), __uint_as_float(row.v[c][2 * i + 1])), sl, mn);
      const float2 b = __ffma2_rn(make_float2(__uint_as_float(row.v[c][2 * i + 2]), __uint_as_float(row.v[c][2 * i + 3])), sl, mn);
      float2 p01 = make_float2(fast_exp2(a.x), fast_exp2(a.y));
      float2 p23 = make_float2(fast_exp2(b.x), fast_exp2(b.y));
      if (MASK) {
        if (c * 32 + 2 * i + 0 >= kvalid) p01.x = 0.f;
        if (c * 32 + 2 * i + 1 >= kvalid) p01.y = 0.f;
        if (c * 32 + 2 * i + 2 >= kvalid) p23.x = 0.f;
        if (c * 32 + 2 * i + 3 >= kvalid) p23.y = 0.f;
      }
      r01 = __fadd2_rn(r01, p01);
      r23 = __fadd2_rn(r23, p23);
      const __nv_bfloat162 h0 = __floats2bfloat162_rn(p01.x, p01.y), h1 = __floats2bfloat162_rn(p23.x, p23.y);
      pk[i] = *reinterpret_cast<const uint32_t*>(&h0);
      pk[i + 1] = *reinterpret_cast<const uint32_t*>(&h1);
    }
    tmem_st16(tP + c * 16, pk);
  }
  return (r01.x + r01.y) + (r23.x + r23.y);
}

// ---------------------------------------------------------------------------------------------------
// Warp-specialised "ping-pong" forward: one CTA per SM works on TWO 128-query tiles at once.
//   warps 0-3  softmax warpgroup of tile A     warps 4-7  softmax warpgroup of tile B
//   warp 8     MMA issuer (one elected lane)    warp 9     TMA producer (one elected lane)
// The MMA warp interleaves the tiles -- PV_A(j), S_A(j+1), PV_B(j), S_B(j+1), ... -- so the tensor core
// works on one tile while the other tile's warpgroup runs its softmax, and both tiles share every K / V
// chunk that TMA brings in (3-stage ring).  The O accumulator is rescaled lazily: the running maximum a
// row normalises by is only moved when the new maximum exceeds it by 2^8 (exact: the final 1/l uses the
// same stale maximum), which removes almost all TMEM round trips of O.
// TMEM (all 512 columns): tile t at column 256 t: S [0,128) | O [128,192) | P [192,256).
// mbarriers: q_full, kv_full[3] (TMA tx), kv_free[3] (tcgen05.commit), s_full[2] (commit),
//            p_full[2] (128 thread arrivals), o_full[2] (commit).
constexpr int AT2_STAGES = 3;
constexpr int AT2_THREADS = 320;
constexpr int AT2_SMEM = (2 + 2 * AT2_STAGES) * AT_TILE_BYTES + 1024 + 256;

struct __align__(8) Attn2Bars {
  uint64_t q_full, kv_full[AT2_STAGES], kv_free[AT2_STAGES], s_full[2], p_full[2], o_full[2], s_free[2];
  uint32_t tmem_base;
};

// UNI: the TMA / MMA warps run with all lanes on warp-uniform values and elect one lane per issue (rf_sm100.cuh
// elect_one(): descriptors in uniform registers, one MMA = UIADD3.64 pair + UTCHMMA); UNI = false keeps the
// `if (lane == 0)` form (~17 SASS instructions per MMA) for A/B measurement (RF_UNIFORM_ISSUE=0).
template <bool UNI>
__global__ void __launch_bounds__(AT2_THREADS, 1)
sr_attention_fwd_pp_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
                           __nv_bfloat16* __restrict__ out, float* __restrict__ lse, int N, int M, int heads,
                           float scale_log2) {
  WS_T_INIT();
  WS_T(0);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;                                   // 2 tiles
  uint8_t* ring = smem + 2 * AT_TILE_BYTES;             // stage s: K at ring + 2 s T, V right after
  Attn2Bars* bars = reinterpret_cast<Attn2Bars*>(smem + (2 + 2 * AT2_STAGES) * AT_TILE_BYTES);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * 2 * AT_BM, head = blockIdx.y, b = blockIdx.z;
  const int C = heads * AT_D;
  const int nchunks = (M + AT_BN - 1) / AT_BN;

  if (tid == 0) {
    mbar_init(&bars->q_full, 1);
    for (int s = 0; s < AT2_STAGES; ++s) {
      mbar_init(&bars->kv_full[s], 1);
      mbar_init(&bars->kv_free[s], 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&bars->s_full[t], 1);
      mbar_init(&bars->p_full[t], 4);
      mbar_init(&bars->o_full[t], 1);
      mbar_init(&bars->s_free[t], 4);
    }
    fence_barrier_init();
  }
  if (warp == 8) tmem_alloc<512>(&bars->tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = UNI ? uniform_u32(bars->tmem_base) : bars->tmem_base;

  constexpr uint32_t IDESC_S = make_idesc(FMT_BF16, 128, 128, 0, 0);
  constexpr uint32_t IDESC_O = make_idesc(FMT_BF16, 128, 64, 0, 1);

  if (warp == 9) {
    // ------------------------------------------------------------------ TMA producer
    if (UNI || lane == 0) {
      if (!UNI || elect_one()) {
        tma_prefetch_desc(&tm_q);
        tma_prefetch_desc(&tm_kv);
        mbar_expect_tx(&bars->q_full, 2 * AT_TILE_BYTES);
        tma_load_3d(sQ, &tm_q, &bars->q_full, head * AT_D, q0, b);
        tma_load_3d(sQ + AT_TILE_BYTES, &tm_q, &bars->q_full, head * AT_D, q0 + AT_BM, b);
      }
      for (int j = 0; j < nchunks; ++j) {
        const int st = j % AT2_STAGES;
        if (j >= AT2_STAGES) mbar_wait(&bars->kv_free[st], ((j / AT2_STAGES) - 1) & 1);
        uint8_t* sK = ring + st * 2 * AT_TILE_BYTES;
        if (!UNI || elect_one()) {
          mbar_expect_tx(&bars->kv_full[st], 2 * AT_TILE_BYTES);
          tma_load_3d(sK, &tm_kv, &bars->kv_full[st], head * AT_D, j * AT_BN, b);
          tma_load_3d(sK + AT_TILE_BYTES, &tm_kv, &bars->kv_full[st], C + head * AT_D, j * AT_BN, b);
        }
        if (UNI) __syncwarp();
      }
    }
  } else if (warp == 8) {
    // ------------------------------------------------------------------ MMA issuer
    if (UNI || lane == 0) {
      const uint64_t descQ[2] = {make_sdesc_sw128(smem_u32(sQ), 16, 1024),
                                 make_sdesc_sw128(smem_u32(sQ + AT_TILE_BYTES), 16, 1024)};
      // ring-stage descriptors built once: stage st adds st * (stage bytes >> 4) to the 14-bit start-address field
      constexpr uint64_t STEP = (uint64_t)((2 * AT_TILE_BYTES) >> 4);
      const uint64_t descK0 = make_sdesc_sw128(smem_u32(ring), 16, 1024);
      const uint64_t descV0 = make_sdesc_sw128(smem_u32(ring + AT_TILE_BYTES), 8192, 1024);
      auto issue_s = [&](int t, int j) {
        const uint64_t descK = descK0 + (uint64_t)(j % AT2_STAGES) * STEP;
        if (!UNI || elect_one()) {
#pragma unroll
          for (int k = 0; k < AT_D / 16; ++k)
            mma_f16_ss(tmem + t * 256, descQ[t] + (uint64_t)(k * 2), descK + (uint64_t)(k * 2), IDESC_S, k > 0 ? 1u : 0u);
          tc_commit(&bars->s_full[t]);
        }
        if (UNI) __syncwarp();
      };
      mbar_wait(&bars->q_full, 0);
      mbar_wait(&bars->kv_full[0], 0);
      tc_fence_after();
      issue_s(0, 0);
      issue_s(1, 0);
      for (int j = 0; j < nchunks; ++j) {
        const int st = j % AT2_STAGES;
        const uint64_t descV = descV0 + (uint64_t)st * STEP;
        if (j + 1 < nchunks) {
          mbar_wait(&bars->kv_full[(j + 1) % AT2_STAGES], ((j + 1) / AT2_STAGES) & 1);
          // the next scores are queued as soon as warpgroup t holds S(j) in registers -- not after it has written P(j):
          // S(j+1) is then long finished when the warpgroup comes back for it (the commit -> wake-up -> issue -> MMA
          // round trip, ~500 cycles, leaves the per-tile critical path)
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            mbar_wait(&bars->s_free[t], j & 1);
            tc_fence_after();
            WS_T(20 + t);
            issue_s(t, j + 1);
          }
        }
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          mbar_wait(&bars->p_full[t], j & 1);   // warpgroup t has written P(j)
          tc_fence_after();
          WS_T(22 + t);
          if (!UNI || elect_one()) {
#pragma unroll
            for (int k = 0; k < AT_BN / 16; ++k)
              mma_f16_ts(tmem + t * 256 + 128, tmem + t * 256 + 192 + k * 8, descV + (uint64_t)(k * 128), IDESC_O,
                         (j > 0 || k > 0) ? 1u : 0u);
            tc_commit(&bars->o_full[t]);
          }
          if (UNI) __syncwarp();
        }
        if (!UNI || elect_one()) tc_commit(&bars->kv_free[st]);   // K_j / V_j fully consumed by both tiles (S(j) was committed earlier)
        if (UNI) __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax warpgroups
    const int t = warp >> 2;                         // tile handled by this warpgroup
    const int r = tid & 127;                         // row inside the tile == TMEM lane
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t tS = tmem + t * 256 + lane_off, tO = tS + 128, tP = tS + 192;
    float m_run = -INFINITY, l_run = 0.f;            // m_run: the (possibly stale) maximum rows are normalised by
    // The exponential phases of the two warpgroups ALTERNATE (named barriers 1 + t, the FlashAttention-3 "ping-pong"):
    // measured on B200, both warpgroups otherwise run in lockstep, share the 16 ex2/clk of the SM during their
    // simultaneous MUFU phases (2 x 1024 cycles) and leave the MUFU idle during their simultaneous load / max / wait
    // phases (~1 200 cycles) -- alternating lets one warpgroup's non-MUFU work hide under the other's exponentials.
    if (t == 1) asm volatile("bar.arrive 1, 256;" ::: "memory");   // warpgroup 0 goes first
    for (int j = 0; j < nchunks; ++j) {
      mbar_wait(&bars->s_full[t], j & 1);
      tc_fence_after();
      WS_T(10);
      const int kvalid = M - j * AT_BN;
      const bool full = kvalid >= AT_BN;               // only the last chunk of a ragged M needs masking
      // S is read ONCE: four tcgen05.ld in flight, one wait (every ld + wait round trip costs ~100 cycles of latency)
      AttnRow srow;
      const float mx = full ? attn_load_max<false>(tS, kvalid, srow) : attn_load_max<true>(tS, kvalid, srow);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->s_free[t]);   // S(j) is in registers: the MMA warp may overwrite it with S(j+1)
      WS_T(11);
      // lazy rescale: move the reference maximum only when it is exceeded by more than 2^8
      const bool move = (mx - m_run) * scale_log2 > 8.f;   // true on the first chunk (m_run = -inf)
      if (j > 0) {   // PV(j-1) must have consumed P(j-1) (and landed in O) before P / O are touched again
        mbar_wait(&bars->o_full[t], (j - 1) & 1);
        tc_fence_after();
      }
      WS_T(12);
      if (__any_sync(0xffffffffu, move)) {
        const float m_new = move ? mx : m_run;
        const float alpha = fast_exp2((m_run - m_new) * scale_log2);   // 0 on the first chunk, 1 for rows that stay
        m_run = m_new;
        l_run *= alpha;
        if (j > 0) {   // rare (lazy rescale): 16 columns at a time, the 128 score registers stay live
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {
            uint32_t v[16];
            tmem_ld16(tO + c * 16, v);
            tc_wait_ld();
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
            tmem_st16(tO + c * 16, v);
          }
        }
      }
      const float mneg = -m_run * scale_log2;
      if (t == 0) asm volatile("bar.sync 1, 256;" ::: "memory"); else asm volatile("bar.sync 2, 256;" ::: "memory");
      WS_T(13);
      const float rs = full ? attn_exp_pack_regs<false>(srow, tP, scale_log2, mneg, kvalid)
                            : attn_exp_pack_regs<true>(srow, tP, scale_log2, mneg, kvalid);
      if (t == 0) asm volatile("bar.arrive 2, 256;" ::: "memory"); else asm volatile("bar.arrive 1, 256;" ::: "memory");
      l_run += rs;
      tc_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->p_full[t]);   // one arrival per warp (128 per-thread arrivals serialise in the barrier unit)
      WS_T(14);
    }
    mbar_wait(&bars->o_full[t], (nchunks - 1) & 1);
    tc_fence_after();
    const int row = q0 + t * AT_BM + r;
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
      for (int i = 0; i < 8; ++i)
        dst[i] = make_uint4(packed[4 * i], packed[4 * i + 1], packed[4 * i + 2], packed[4 * i + 3]);
      if (lse != nullptr)
        lse[((long)b * heads + head) * N + row] = m_run * scale_log2 * 0.69314718055994531f + __logf(l_run);
    }
  }
  tc_fence_before();
  __syncthreads();
  WS_T(2);
  WS_T_FLUSH();
  if (warp == 8) tmem_dealloc<512>(tmem);
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
    RF_CUDA(cudaFuncSetAttribute(sr_attention_fwd_pp_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT2_SMEM));
    RF_CUDA(cudaFuncSetAttribute(sr_attention_fwd_pp_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT2_SMEM));
    attr_set = true;
  }
  const float scale_log2 = scale * 1.44269504088896341f;
  static const int variant = getenv("RF_ATTN_FWD") ? atoi(getenv("RF_ATTN_FWD")) : 2;
  if (variant == 2 && N > AT_BM) {   // two query tiles per CTA, warp-specialised
    dim3 grid((unsigned)((N + 2 * AT_BM - 1) / (2 * AT_BM)), (unsigned)heads, (unsigned)B);
    static const bool uni = [] { const char* e = getenv("RF_UNIFORM_ISSUE"); return !(e && e[0] == '0'); }();
    auto kern = uni ? sr_attention_fwd_pp_kernel<true> : sr_attention_fwd_pp_kernel<false>;
    kern<<<grid, AT2_THREADS, AT2_SMEM, (cudaStream_t)stream>>>(tq, tkv, (__nv_bfloat16*)out, lse, N, M, heads, scale_log2);
    RF_CHECK_LAUNCH("sr_attention_fwd_pp_kernel");
    return RF_OK;
  }
  dim3 grid((unsigned)((N + AT_BM - 1) / AT_BM), (unsigned)heads, (unsigned)B);
  sr_attention_fwd_kernel<<<grid, 128, AT_SMEM, (cudaStream_t)stream>>>(tq, tkv, (__nv_bfloat16*)out, lse, N, M, heads,
                                                                          scale_log2);
  RF_CHECK_LAUNCH("sr_attention_fwd_kernel");
  return RF_OK;
}
