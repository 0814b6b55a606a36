// Warp-specialised backward of the fused SR-attention core on sm_100a (tcgen05 + TMEM + TMA), ONE launch.
// Autograd of /root/reference/models/backbones/mix_transformer.py:150-160 without the saved [B, heads, N, N_kv]
// probability matrix (P is recomputed from Q, K and the forward's log-sum-exp):
//   P  = exp(scale * Q K^T - LSE)            dP = dO V^T            D = rowsum(dO * O)
//   dS = scale * P * (dP - D)                dQ = dS K              dK = dS^T Q          dV = P^T dO
//
// Why this shape (measured on B200, profiles/r02_ubench_sm100.json): a tcgen05.commit -> mbarrier -> issue round
// trip costs ~400 cycles, so the previous single-role 128-thread kernels (MMA -> softmax -> MMA strictly in sequence)
// sat at 20 % tensor-pipe; TMEM reads are NOT a limit (2-4 KB/clk/SM); N = 64 MMAs with both operands in shared
// memory are shared-memory-bound (58.6 clk instead of 32), N = 128 ones are not (75 vs 64) and the TMEM-A form
// of an N = 64 MMA costs 42.  Hence 128 x 128 score tiles, every [128 x 64] += product with its A operand in TMEM,
// and dedicated TMA / MMA warps that keep the tensor pipe queued while two softmax warpgroups work.
//
// One grid, two CTA roles (block index order = longest units first, the hardware block scheduler is the list
// scheduler; all CTAs are 320 threads, 1 per SM, 512 TMEM columns):
//   role KV (blocks [0, nKV)) : 128 keys of one (b, head) x a contiguous range of 128-query tiles.  Lane = key.
//        S^T = K Q^T, dP^T = V dO^T (M128 N128 K16 x4 each) -> softmax warpgroup g owns query columns [64g, 64g+64):
//        loads its S^T / dP^T columns into registers, releases them at once (the MMA warp queues the NEXT tile's
//        S^T / dP^T while the exponentials run), writes P^T and dS^T as bf16 TMEM A operands ->
//        dV += P^T dO, dK += dS^T Q (M128 N64 K16 x8, B = the Q / dO tile read MN-major from the same bytes).
//        Results leave with red.global.add.v4.f32 into the small fp32 [B, M, 2C] buffer.
//   role Q  (blocks [nKV, nKV+nQ)): 128 queries of one (b, head), all keys in chunks of 128.  Lane = query.
//        S = Q K^T, dP = dO V^T -> warpgroup g owns key columns [64g, 64g+64) -> dS (bf16, double-buffered TMEM
//        A operand) -> dQ += dS K (M128 N64 K16 x8, B = the K chunk MN-major).
// `scale` is applied once per output element in the epilogues (dK, dQ), not per score.
// No masking is needed for ragged N / M: TMA zero-fills rows outside the tensors, a zero Q / dO row or K / V row
// contributes exact zeros to every product, the padded D / LSE workspace entries are zero, and out-of-range
// rows are simply not stored.
//
// TMEM columns   role KV: S^T [0,128) | dP^T [128,256) | P^T [256,320) | dS^T [320,384) | dK [384,448) | dV [448,512)
//                role Q : S   [0,128) | dP   [128,256) | dS0 [256,320) | dS1  [320,384) | dQ [384,448)
#include <cuda_bf16.h>
#include <stdlib.h>

#include "rf_common.cuh"
#include "rf_sm100.cuh"
#include "rf_trace.cuh"

namespace rf {
using namespace sm100;

constexpr int WS_TILE = 128 * 64 * 2;                 // one [128 x 64] bf16 tile, 16 KiB
constexpr int WS_VEC = 128 * 4;                       // 128 floats
constexpr int WS_STAGES = 3;
constexpr int WS_STAGE_BYTES = 2 * WS_TILE + 2 * WS_VEC;   // role KV: Q | dO | lse2 | D   (role Q uses the first 2 tiles: K | V)
constexpr int WS_THREADS = 320;
constexpr int WS_SMEM = 2 * WS_TILE + WS_STAGES * WS_STAGE_BYTES + 1024 /*alignment*/ + 256 /*barriers*/;

struct __align__(8) WsBars {
  uint64_t fixed_full;                         // role KV: K, V block; role Q: Q, dO tile
  uint64_t ring_full[WS_STAGES], ring_free[WS_STAGES];
  uint64_t s_full;                             // S and dP tiles complete (tcgen05.commit)
  uint64_t s_free;                             // ... loaded into registers by all 8 softmax warps (one elected arrive per warp)
  uint64_t ds_full[2], ds_free[2];             // bf16 A operands (dS; role KV: + P^T) written / consumed (role Q: double-buffered; role KV: index 0)
  uint64_t fin;
  uint32_t tmem_base;
};

struct WsParams {
  const float* dvec;      // [B*heads][Npad]  -D = -rowsum(dO * O)
  const float* lse2;      // [B*heads][Npad]  -LSE * log2(e)
  __nv_bfloat16* dq;      // [B, N, C]
  float* dkv;             // [B, M, 2C] fp32, zeroed by the caller
  int N, M, heads, Npad;
  int n_kv_units, kv_blocks, kv_splits, tiles_per_split, q_tiles;
  float scale, scale_log2;
};

__device__ __forceinline__ uint32_t ws_pack_bf16(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__device__ __forceinline__ float4 lds128(uint32_t saddr) {   // explicit shared-space load (a generic LD costs a long-scoreboard round trip)
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
  return v;
}
// dS pair and (optionally) P pair of two scores, packed fp32 pipe (FFMA2 / FADD2 / FMUL2): 3 issue slots + 2 MUFU per pair
//   p = 2^(s * scale_log2 - lse2)      ds = p * (dp - D)        (nl = -lse2, nd = -D)
__device__ __forceinline__ float2 ws_prob2(uint32_t s0, uint32_t s1, float2 sl, float2 nl) {
  const float2 t = __ffma2_rn(make_float2(__uint_as_float(s0), __uint_as_float(s1)), sl, nl);
  return make_float2(fast_exp2(t.x), fast_exp2(t.y));
}
__device__ __forceinline__ uint32_t ws_ds2(float2 p, uint32_t g0, uint32_t g1, float2 nd) {
  const float2 d = __fmul2_rn(p, __fadd2_rn(make_float2(__uint_as_float(g0), __uint_as_float(g1)), nd));
  return ws_pack_bf16(d.x, d.y);
}

constexpr uint32_t WS_IDESC_S = make_idesc(FMT_BF16, 128, 128, 0, 0);   // both operands K-major
constexpr uint32_t WS_IDESC_ACC = make_idesc(FMT_BF16, 128, 64, 0, 1);  // A from TMEM, B MN-major

// ---------------------------------------------------------------------------------------------------- role KV
__device__ __forceinline__ void ws_role_kv(const CUtensorMap& tm_q, const CUtensorMap& tm_do, const CUtensorMap& tm_kv,
                                           const WsParams& p, uint8_t* smem, WsBars* bars, uint32_t tmem, int unit) {
  const int tid = threadIdx.x, warp = tid >> 5;
  const int split = unit % p.kv_splits;
  const int kb = (unit / p.kv_splits) % p.kv_blocks;
  const int bh = unit / (p.kv_splits * p.kv_blocks);
  const int b = bh / p.heads, head = bh % p.heads;
  const int C = p.heads * 64;
  const int t_begin = split * p.tiles_per_split;
  const int t_end = min(p.q_tiles, t_begin + p.tiles_per_split);
  const int ntiles = t_end - t_begin;
  if (ntiles <= 0) return;   // uniform per CTA
  const int kv0 = kb * 128;
  uint8_t* sK = smem;
  uint8_t* sV = smem + WS_TILE;
  uint8_t* ring = smem + 2 * WS_TILE;

  if (warp == 9) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      mbar_expect_tx(&bars->fixed_full, 2 * WS_TILE);
      tma_load_3d(sK, &tm_kv, &bars->fixed_full, head * 64, kv0, b);
      tma_load_3d(sV, &tm_kv, &bars->fixed_full, C + head * 64, kv0, b);
    }
    __syncwarp();
    const float* lse_row = p.lse2 + (long)bh * p.Npad;
    const float* d_row = p.dvec + (long)bh * p.Npad;
    for (int i = 0; i < ntiles; ++i) {
      const int st = i % WS_STAGES;
      if (i >= WS_STAGES) mbar_wait(&bars->ring_free[st], ((i / WS_STAGES) - 1) & 1);
      uint8_t* stage = ring + st * WS_STAGE_BYTES;
      const int q0 = (t_begin + i) * 128;
      if (elect_one()) {
        mbar_expect_tx(&bars->ring_full[st], WS_STAGE_BYTES);
        tma_load_3d(stage, &tm_q, &bars->ring_full[st], head * 64, q0, b);
        tma_load_3d(stage + WS_TILE, &tm_do, &bars->ring_full[st], head * 64, q0, b);
        bulk_load_1d(stage + 2 * WS_TILE, lse_row + q0, WS_VEC, &bars->ring_full[st]);
        bulk_load_1d(stage + 2 * WS_TILE + WS_VEC, d_row + q0, WS_VEC, &bars->ring_full[st]);
      }
      __syncwarp();
      WS_T(30);
    }
  } else if (warp == 8) {
    // ------------------------------------------------------------------ MMA issuer
    // per tile i, in this order:  S^T(i+1) | dV(i) | dP^T(i+1) | dK(i)   -- the next S^T is queued as soon as the softmax
    // warps hold S^T(i) in registers, so it is long finished when they come back for it
    const uint64_t descK = make_sdesc_sw128(smem_u32(sK), 16, 1024);
    const uint64_t descV = make_sdesc_sw128(smem_u32(sV), 16, 1024);
    // ring-stage descriptors: built once; stage st adds st * (stage bytes >> 4) to the 14-bit start-address field
    constexpr uint64_t STEP = (uint64_t)(WS_STAGE_BYTES >> 4);
    const uint32_t ring_a = smem_u32(ring);
    const uint64_t dQk0 = make_sdesc_sw128(ring_a, 16, 1024), dOk0 = make_sdesc_sw128(ring_a + WS_TILE, 16, 1024);
    const uint64_t dQmn0 = make_sdesc_sw128(ring_a, 8192, 1024), dOmn0 = make_sdesc_sw128(ring_a + WS_TILE, 8192, 1024);
    auto issue_s = [&](uint64_t st) {
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          mma_f16_ss(tmem, descK + (uint64_t)(k * 2), dQk0 + st * STEP + (uint64_t)(k * 2), WS_IDESC_S, k > 0 ? 1u : 0u);
      }
      __syncwarp();
    };
    auto issue_dp = [&](uint64_t st) {
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          mma_f16_ss(tmem + 128, descV + (uint64_t)(k * 2), dOk0 + st * STEP + (uint64_t)(k * 2), WS_IDESC_S, k > 0 ? 1u : 0u);
        tc_commit(&bars->s_full);   // S and dP of this tile are complete (in-order tensor pipe)
      }
      __syncwarp();
    };
    mbar_wait(&bars->fixed_full, 0);
    mbar_wait(&bars->ring_full[0], 0);
    tc_fence_after();
    issue_s(0);
    issue_dp(0);
    int st = 0;
    for (int i = 0; i < ntiles; ++i) {
      const int st1 = st + 1 == WS_STAGES ? 0 : st + 1;
      const bool more = i + 1 < ntiles;
      if (more) {
        mbar_wait(&bars->ring_full[st1], ((i + 1) / WS_STAGES) & 1);
        mbar_wait(&bars->s_free, i & 1);        // S^T(i) and dP^T(i) are in registers
        tc_fence_after();
        issue_s(st1);
        issue_dp(st1);
      }
      mbar_wait(&bars->ds_full[0], i & 1);      // P^T(i), dS^T(i) written
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
          mma_f16_ts(tmem + 448, tmem + 256 + k * 8, dOmn0 + st * STEP + (uint64_t)(k * 128), WS_IDESC_ACC, (i > 0 || k > 0) ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          mma_f16_ts(tmem + 384, tmem + 320 + k * 8, dQmn0 + st * STEP + (uint64_t)(k * 128), WS_IDESC_ACC, (i > 0 || k > 0) ? 1u : 0u);
        tc_commit(&bars->ds_free[0]);
        tc_commit(&bars->ring_free[st]);
        if (!more) tc_commit(&bars->fin);
      }
      __syncwarp();
      WS_T(25);
      st = st1;
    }
  } else {
    // ------------------------------------------------------------------ softmax warpgroups (lane = key)
    const int g = warp >> 2;                       // query columns [64 g, 64 g + 64)
    const bool lane0 = (tid & 31) == 0;
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t tS = tmem + lane_off + 64 * g, tdP = tmem + lane_off + 128 + 64 * g;
    const uint32_t tP = tmem + lane_off + 256 + 32 * g, tdS = tmem + lane_off + 320 + 32 * g;
    const float2 sl = make_float2(p.scale_log2, p.scale_log2);
    for (int i = 0; i < ntiles; ++i) {
      const int st = i % WS_STAGES;
      const uint32_t vec = smem_u32(ring + st * WS_STAGE_BYTES + 2 * WS_TILE) + 256 * g;   // -lse2 | -D of this warpgroup's 64 queries
      // the -lse2 / -D vectors of this stage landed before the MMA warp issued S^T(i) (it waited on ring_full), and
      // s_full is observed after that: no separate wait (every mbarrier test costs ~100 cycles of latency here)
      mbar_wait(&bars->s_full, i & 1);
      tc_fence_after();
      WS_T(10);
      uint32_t sv[2][32], dv[2][32];
      tmem_ld32(tS, sv[0]);
      tmem_ld32(tS + 32, sv[1]);
      tmem_ld32(tdP, dv[0]);
      tmem_ld32(tdP + 32, dv[1]);
      tc_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane0) mbar_arrive(&bars->s_free);   // one arrival per warp: 256 per-thread arrivals serialise in the barrier unit (~380 cycles)
      WS_T(11);
      // ONE compute phase: the exponentials (MUFU) and the dS arithmetic (FMA pipe) interleave
      uint32_t pp[32], pd[32];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 l = lds128(vec + (h * 8 + q) * 16), dd = lds128(vec + WS_VEC + (h * 8 + q) * 16);
          const float2 p01 = ws_prob2(sv[h][4 * q + 0], sv[h][4 * q + 1], sl, make_float2(l.x, l.y));
          const float2 p23 = ws_prob2(sv[h][4 * q + 2], sv[h][4 * q + 3], sl, make_float2(l.z, l.w));
          pp[h * 16 + 2 * q] = ws_pack_bf16(p01.x, p01.y);
          pp[h * 16 + 2 * q + 1] = ws_pack_bf16(p23.x, p23.y);
          pd[h * 16 + 2 * q] = ws_ds2(p01, dv[h][4 * q + 0], dv[h][4 * q + 1], make_float2(dd.x, dd.y));
          pd[h * 16 + 2 * q + 1] = ws_ds2(p23, dv[h][4 * q + 2], dv[h][4 * q + 3], make_float2(dd.z, dd.w));
        }
      }
      if (i >= 1) {   // dV / dK MMAs of tile i-1 have consumed P^T / dS^T
        mbar_wait(&bars->ds_free[0], (i - 1) & 1);
        tc_fence_after();
      }
      tmem_st32(tP, pp);
      tmem_st32(tdS, pd);
      tc_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane0) mbar_arrive(&bars->ds_full[0]);
      WS_T(14);
    }
    // epilogue: warpgroup 0 -> dK (scaled), warpgroup 1 -> dV
    mbar_wait(&bars->fin, 0);
    tc_fence_after();
    WS_T(15);
    const int kv = kv0 + (tid & 127);
    const uint32_t tacc = tmem + lane_off + (g == 0 ? 384 : 448);
    const float mul = g == 0 ? p.scale : 1.f;
    float* dst = p.dkv + ((long)b * p.M + kv) * 2 * C + (g == 0 ? 0 : C) + head * 64;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t v[32];
      tmem_ld32(tacc + c * 32, v);
      tc_wait_ld();
      if (kv < p.M) {
        if (p.kv_splits == 1) {   // this unit saw every query: plain stores, no atomics
#pragma unroll
          for (int e = 0; e < 32; e += 4)
            *reinterpret_cast<float4*>(dst + c * 32 + e) =
                make_float4(__uint_as_float(v[e]) * mul, __uint_as_float(v[e + 1]) * mul, __uint_as_float(v[e + 2]) * mul,
                            __uint_as_float(v[e + 3]) * mul);
        } else {
#pragma unroll
          for (int e = 0; e < 32; e += 4)
            red_add_v4(dst + c * 32 + e, __uint_as_float(v[e]) * mul, __uint_as_float(v[e + 1]) * mul,
                       __uint_as_float(v[e + 2]) * mul, __uint_as_float(v[e + 3]) * mul);
        }
      }
    }
    WS_T(16);
  }
}

// ---------------------------------------------------------------------------------------------------- role Q
__device__ __forceinline__ void ws_role_q(const CUtensorMap& tm_q, const CUtensorMap& tm_do, const CUtensorMap& tm_kv,
                                          const WsParams& p, uint8_t* smem, WsBars* bars, uint32_t tmem, int unit) {
  const int tid = threadIdx.x, warp = tid >> 5;
  const int qt = unit % p.q_tiles;
  const int bh = unit / p.q_tiles;
  const int b = bh / p.heads, head = bh % p.heads;
  const int C = p.heads * 64;
  const int q0 = qt * 128;
  const int nchunks = (p.M + 127) / 128;
  uint8_t* sQ = smem;
  uint8_t* sdO = smem + WS_TILE;
  uint8_t* ring = smem + 2 * WS_TILE;

  if (warp == 9) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      mbar_expect_tx(&bars->fixed_full, 2 * WS_TILE);
      tma_load_3d(sQ, &tm_q, &bars->fixed_full, head * 64, q0, b);
      tma_load_3d(sdO, &tm_do, &bars->fixed_full, head * 64, q0, b);
    }
    __syncwarp();
    for (int j = 0; j < nchunks; ++j) {
      const int st = j % WS_STAGES;
      if (j >= WS_STAGES) mbar_wait(&bars->ring_free[st], ((j / WS_STAGES) - 1) & 1);
      uint8_t* stage = ring + st * WS_STAGE_BYTES;
      if (elect_one()) {
        mbar_expect_tx(&bars->ring_full[st], 2 * WS_TILE);
        tma_load_3d(stage, &tm_kv, &bars->ring_full[st], head * 64, j * 128, b);
        tma_load_3d(stage + WS_TILE, &tm_kv, &bars->ring_full[st], C + head * 64, j * 128, b);
      }
      __syncwarp();
    }
  } else if (warp == 8) {
    // ------------------------------------------------------------------ MMA issuer
    // per chunk j, in this order:  S(j+1) | dP(j+1) | dQ(j)
    const uint64_t descQ = make_sdesc_sw128(smem_u32(sQ), 16, 1024);
    const uint64_t descdO = make_sdesc_sw128(smem_u32(sdO), 16, 1024);
    constexpr uint64_t STEP = (uint64_t)(WS_STAGE_BYTES >> 4);
    const uint32_t ring_a = smem_u32(ring);
    const uint64_t dKk0 = make_sdesc_sw128(ring_a, 16, 1024), dVk0 = make_sdesc_sw128(ring_a + WS_TILE, 16, 1024);
    const uint64_t dKmn0 = make_sdesc_sw128(ring_a, 8192, 1024);
    auto issue_s = [&](uint64_t st) {
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          mma_f16_ss(tmem, descQ + (uint64_t)(k * 2), dKk0 + st * STEP + (uint64_t)(k * 2), WS_IDESC_S, k > 0 ? 1u : 0u);
      }
      __syncwarp();
    };
    auto issue_dp = [&](uint64_t st) {
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          mma_f16_ss(tmem + 128, descdO + (uint64_t)(k * 2), dVk0 + st * STEP + (uint64_t)(k * 2), WS_IDESC_S, k > 0 ? 1u : 0u);
        tc_commit(&bars->s_full);   // S and dP of this tile are complete (in-order tensor pipe)
      }
      __syncwarp();
    };
    mbar_wait(&bars->fixed_full, 0);
    mbar_wait(&bars->ring_full[0], 0);
    tc_fence_after();
    issue_s(0);
    issue_dp(0);
    int st = 0;
    for (int j = 0; j < nchunks; ++j) {
      const int st1 = st + 1 == WS_STAGES ? 0 : st + 1;
      const bool more = j + 1 < nchunks;
      if (more) {
        mbar_wait(&bars->ring_full[st1], ((j + 1) / WS_STAGES) & 1);
        mbar_wait(&bars->s_free, j & 1);
        tc_fence_after();
        issue_s(st1);
        issue_dp(st1);
      }
      mbar_wait(&bars->ds_full[j & 1], (j >> 1) & 1);   // dS(j) written
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
          mma_f16_ts(tmem + 384, tmem + 256 + (j & 1) * 64 + k * 8, dKmn0 + st * STEP + (uint64_t)(k * 128), WS_IDESC_ACC,
                     (j > 0 || k > 0) ? 1u : 0u);
        tc_commit(&bars->ds_free[j & 1]);
        tc_commit(&bars->ring_free[st]);
        if (!more) tc_commit(&bars->fin);
      }
      __syncwarp();
      WS_T(25);
      st = st1;
    }
  } else {
    // ------------------------------------------------------------------ softmax warpgroups (lane = query)
    const int g = warp >> 2;                       // key columns [64 g, 64 g + 64) of every chunk
    const bool lane0 = (tid & 31) == 0;
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t tS = tmem + lane_off + 64 * g, tdP = tmem + lane_off + 128 + 64 * g;
    const int row = q0 + (tid & 127);              // < Npad: the workspace rows are padded to whole tiles
    const float ndrow = __ldg(p.dvec + (long)bh * p.Npad + row);
    const float nlse = __ldg(p.lse2 + (long)bh * p.Npad + row);
    const float2 sl = make_float2(p.scale_log2, p.scale_log2), nl = make_float2(nlse, nlse), nd = make_float2(ndrow, ndrow);
    for (int j = 0; j < nchunks; ++j) {
      mbar_wait(&bars->s_full, j & 1);
      tc_fence_after();
      WS_T(10);
      uint32_t sv[2][32], dv[2][32];
      tmem_ld32(tS, sv[0]);
      tmem_ld32(tS + 32, sv[1]);
      tmem_ld32(tdP, dv[0]);
      tmem_ld32(tdP + 32, dv[1]);
      tc_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane0) mbar_arrive(&bars->s_free);
      WS_T(11);
      uint32_t pd[32];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
#pragma unroll
        for (int e = 0; e < 16; ++e)
          pd[h * 16 + e] = ws_ds2(ws_prob2(sv[h][2 * e], sv[h][2 * e + 1], sl, nl), dv[h][2 * e], dv[h][2 * e + 1], nd);
      }
      if (j >= 2) {   // dQ MMAs of chunk j-2 have consumed this dS buffer
        mbar_wait(&bars->ds_free[j & 1], ((j >> 1) - 1) & 1);
        tc_fence_after();
      }
      tmem_st32(tmem + lane_off + 256 + (j & 1) * 64 + 32 * g, pd);
      tc_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane0) mbar_arrive(&bars->ds_full[j & 1]);
      WS_T(14);
    }
    // epilogue: warpgroup g stores dQ columns [32 g, 32 g + 32) of its row
    mbar_wait(&bars->fin, 0);
    tc_fence_after();
    WS_T(15);
    uint32_t v[32];
    tmem_ld32(tmem + lane_off + 384 + 32 * g, v);
    tc_wait_ld();
    if (row < p.N) {
      uint4* dst = reinterpret_cast<uint4*>(p.dq + ((long)b * p.N + row) * C + head * 64 + 32 * g);
#pragma unroll
      for (int e = 0; e < 4; ++e)
        dst[e] = make_uint4(ws_pack_bf16(__uint_as_float(v[8 * e + 0]) * p.scale, __uint_as_float(v[8 * e + 1]) * p.scale),
                            ws_pack_bf16(__uint_as_float(v[8 * e + 2]) * p.scale, __uint_as_float(v[8 * e + 3]) * p.scale),
                            ws_pack_bf16(__uint_as_float(v[8 * e + 4]) * p.scale, __uint_as_float(v[8 * e + 5]) * p.scale),
                            ws_pack_bf16(__uint_as_float(v[8 * e + 6]) * p.scale, __uint_as_float(v[8 * e + 7]) * p.scale));
    }
  }
}

__global__ void __launch_bounds__(WS_THREADS, 1)
sr_attention_bwd_ws_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_do,
                           const __grid_constant__ CUtensorMap tm_kv, const WsParams p) {
  WS_T_INIT();
  WS_T(0);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  WsBars* bars = reinterpret_cast<WsBars*>(smem + 2 * WS_TILE + WS_STAGES * WS_STAGE_BYTES);
  const int warp = threadIdx.x >> 5;
  if (warp == 9) {
    if (elect_one()) {
      tma_prefetch_desc(&tm_q);
      tma_prefetch_desc(&tm_do);
      tma_prefetch_desc(&tm_kv);
      mbar_init(&bars->fixed_full, 1);
      for (int s = 0; s < WS_STAGES; ++s) {
        mbar_init(&bars->ring_full[s], 1);
        mbar_init(&bars->ring_free[s], 1);
      }
      mbar_init(&bars->s_full, 1);
      mbar_init(&bars->s_free, 8);
      for (int s = 0; s < 2; ++s) {
        mbar_init(&bars->ds_full[s], 8);
        mbar_init(&bars->ds_free[s], 1);
      }
      mbar_init(&bars->fin, 1);
      fence_barrier_init();
    }
    __syncwarp();
  }
  if (warp == 8) tmem_alloc<512>(&bars->tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = uniform_u32(bars->tmem_base);
  const int unit = blockIdx.x;
  WS_T(1);
  if (unit < p.n_kv_units)
    ws_role_kv(tm_q, tm_do, tm_kv, p, smem, bars, tmem, unit);
  else
    ws_role_q(tm_q, tm_do, tm_kv, p, smem, bars, tmem, unit - p.n_kv_units);
  tc_fence_before();
  __syncthreads();
  WS_T(2);
  WS_T_FLUSH();
  if (warp == 8) tmem_dealloc<512>(tmem);
}

// -D[bh][n] = -sum_d dO * O and -LSE * log2(e), rows padded to whole 128-query tiles with zeros.  Eight lanes per
// (row, head) unit: one warp instruction reads 512 contiguous bytes of O / dO (four units).
__global__ void __launch_bounds__(256)
sr_attention_bwd_prep_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ dout,
                             const float* __restrict__ lse, float* __restrict__ dvec, float* __restrict__ lse2, int N,
                             int Npad, int heads, long units, long pad_entries) {
  const long t = blockIdx.x * (long)blockDim.x + threadIdx.x;
  const long u = t >> 3;                 // ((b * N + row) * heads + head): 128-byte chunks in memory order
  const int l8 = (int)(t & 7);
  float d = 0.f;
  if (u < units) {
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(o) + u * 8 + l8);
    const uint4 g = __ldg(reinterpret_cast<const uint4*>(dout) + u * 8 + l8);
    const __nv_bfloat162* ha = reinterpret_cast<const __nv_bfloat162*>(&a);
    const __nv_bfloat162* hg = reinterpret_cast<const __nv_bfloat162*>(&g);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 fa = __bfloat1622float2(ha[k]), fg = __bfloat1622float2(hg[k]);
      d = fmaf(fa.x, fg.x, d);
      d = fmaf(fa.y, fg.y, d);
    }
  }
  d += __shfl_xor_sync(0xffffffffu, d, 1);
  d += __shfl_xor_sync(0xffffffffu, d, 2);
  d += __shfl_xor_sync(0xffffffffu, d, 4);
  if (u < units && l8 == 0) {
    const int head = (int)(u % heads);
    const long brow = u / heads;
    const long b = brow / N;
    const int row = (int)(brow % N);
    const long bh = b * heads + head;
    dvec[bh * Npad + row] = -d;     // both vectors are stored negated: the kernels add them (FFMA2 / FADD2 operands)
    lse2[bh * Npad + row] = -__ldg(lse + bh * N + row) * 1.44269504088896341f;
  }
  if (t < pad_entries) {            // zero the padding rows [N, Npad) of every (b, head)
    const int npadrows = Npad - N;
    const long bh = t / npadrows;
    const int row = N + (int)(t % npadrows);
    dvec[bh * Npad + row] = 0.f;
    lse2[bh * Npad + row] = 0.f;
  }
}

}  // namespace rf

using namespace rf;

// workspace: D and LSE*log2(e), each [B*heads][Npad] fp32
int64_t rf_sr_attention_bwd_ws_workspace_bytes(int B, int N, int heads) {
  const int64_t npad = ((int64_t)N + 127) / 128 * 128;
  return 2 * (int64_t)sizeof(float) * B * heads * npad;
}

int rf_sr_attention_bwd_ws(const void* q, const void* kv, const void* out, const void* grad_out, const float* lse,
                           void* grad_q, float* grad_kv_f32, void* workspace, int B, int N, int M, int heads, float scale,
                           cudaStream_t st) {
  const int C = heads * 64;
  const int npad = (N + 127) / 128 * 128;
  float* dvec = (float*)workspace;
  float* lse2 = dvec + (size_t)B * heads * npad;
  CUtensorMap tq, tdo, tkv;
  int rc = make_tmap_3d(&tq, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, q, (uint64_t)C, (uint64_t)N, (uint64_t)B, (uint64_t)C * 2,
                        (uint64_t)N * C * 2, 64, 128);
  if (rc != RF_OK) return rc;
  rc = make_tmap_3d(&tdo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, grad_out, (uint64_t)C, (uint64_t)N, (uint64_t)B,
                    (uint64_t)C * 2, (uint64_t)N * C * 2, 64, 128);
  if (rc != RF_OK) return rc;
  rc = make_tmap_3d(&tkv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, kv, (uint64_t)2 * C, (uint64_t)M, (uint64_t)B,
                    (uint64_t)2 * C * 2, (uint64_t)M * 2 * C * 2, 64, 128);
  if (rc != RF_OK) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    RF_CUDA(cudaFuncSetAttribute(sr_attention_bwd_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WS_SMEM));
    attr_set = true;
  }
  {
    const long units = (long)B * N * heads, pad_entries = (long)B * heads * (npad - N);
    const long threads = units * 8 > pad_entries ? units * 8 : pad_entries;
    sr_attention_bwd_prep_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(
        (const __nv_bfloat16*)out, (const __nv_bfloat16*)grad_out, lse, dvec, lse2, N, npad, heads, units, pad_entries);
    RF_CHECK_LAUNCH("sr_attention_bwd_prep_kernel");
  }
  WsParams p;
  p.dvec = dvec;
  p.lse2 = lse2;
  p.dq = (__nv_bfloat16*)grad_q;
  p.dkv = grad_kv_f32;
  p.N = N;
  p.M = M;
  p.heads = heads;
  p.Npad = npad;
  p.q_tiles = npad / 128;
  p.kv_blocks = (M + 127) / 128;
  // key-stationary units: (b, head, key block) x query splits, issued before the (shorter) query-stationary units
  // (block index order = LPT list scheduling by the hardware).  A split costs a prologue and a 64 KiB red.global
  // epilogue (~4 000 cycles), so use the FEWEST splits that keep one key-stationary unit below ~70 % of the
  // estimated makespan (both roles' tile steps summed over 148 SMs); with one split the unit stores its result plainly.
  const long bhk = (long)B * heads * p.kv_blocks;
  const long steps_kv = bhk * p.q_tiles, steps_q = (long)B * heads * p.q_tiles * p.kv_blocks;
  static const int force_splits = [] { const char* e = getenv("RF_ATTN_BWD_KV_SPLITS"); return e ? atoi(e) : 0; }();
  const double makespan = (1.5 * steps_kv + 1.0 * steps_q) / kNumSMs + 4.0;   // in Q-role tile steps
  long tps = (long)(0.7 * makespan / 1.5);
  if (tps < 4) tps = 4;
  if (tps > p.q_tiles) tps = p.q_tiles;
  p.kv_splits = (int)((p.q_tiles + tps - 1) / tps);
  if (force_splits > 0) p.kv_splits = force_splits > p.q_tiles ? p.q_tiles : force_splits;
  p.tiles_per_split = (int)((p.q_tiles + p.kv_splits - 1) / p.kv_splits);
  p.n_kv_units = (int)(bhk * p.kv_splits);
  if (p.kv_splits > 1) RF_CUDA(cudaMemsetAsync(grad_kv_f32, 0, sizeof(float) * (size_t)B * M * 2 * C, st));
  const long n_q_units = (long)B * heads * p.q_tiles;
  RF_REQUIRE(p.n_kv_units + n_q_units < (1l << 31), "rf_sr_attention_bwd: grid too large");
  p.scale = scale;
  p.scale_log2 = scale * 1.44269504088896341f;
  sr_attention_bwd_ws_kernel<<<(unsigned)(p.n_kv_units + n_q_units), WS_THREADS, WS_SMEM, st>>>(tq, tdo, tkv, p);
  RF_CHECK_LAUNCH("sr_attention_bwd_ws_kernel");
  return RF_OK;
}
