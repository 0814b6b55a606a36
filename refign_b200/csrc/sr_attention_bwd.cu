// Backward of the fused SR-attention core (sr_attention.cu) on sm_100a: tcgen05 + TMEM + TMA.
// Autograd of /root/reference/models/backbones/mix_transformer.py:150-160 without the saved
// [B, heads, N, N_kv] probability matrix: P is recomputed from Q, K and the forward's log-sum-exp.
//   P  = exp(scale * Q K^T - LSE)            dP = dO V^T            D = rowsum(dO * O)
//   dS = scale * P * (dP - D)                dQ = dS K              dK = dS^T Q          dV = P^T dO
// Two kernels, no global atomics on the large tensors:
//   A  sr_attention_bwd_dq_kernel : CTA = 128 queries of one (b, head); streams K / V chunks of 64 keys;
//      S, dP, dQ accumulate in TMEM (S 64 | dP 64 | dS 32 | dQ 64 columns), dS is fed back to the
//      tensor core as the TMEM A operand.  Also writes D[b,h,n].
//   B  sr_attention_bwd_dkv_kernel: CTA = 128 keys of one (b, head) x a split of the query range; streams
//      Q / dO tiles of 64 queries; S^T = K Q^T and dP^T = V dO^T put the key on the TMEM lane, so
//      P^T and dS^T are directly the A operands of dV += P^T dO and dK += dS^T Q (the Q / dO tiles are
//      read a second time by the tensor core as MN-major B operands -- same shared-memory bytes).
//      TMEM: S^T|P^T 64 | dP^T|dS^T 64 | dK 64 | dV 64 columns.  The (few) query splits are combined
//      with fp32 red.global on the small [B, M, 2C] result.
// Both kernels: 128 threads, thread t owns TMEM lane t, 4-stage TMA ring, 2 CTAs per SM.
#include <cuda_bf16.h>

#include <type_traits>

#include "rf_common.cuh"
#include <stdlib.h>

#include "rf_sm100.cuh"

namespace rf {
using namespace sm100;

constexpr int AB_D = 64;
constexpr int AB_T128 = 128 * AB_D * 2;  // 16 KiB tile of 128 rows
constexpr int AB_T64 = 64 * AB_D * 2;    // 8 KiB tile of 64 rows
constexpr int AB_STAGES = 4;
constexpr int AB_SMEM = 2 * AB_T128 + AB_STAGES * 2 * AB_T64 + 1024 + 1024 + 256;

struct __align__(8) BwdBars {
  uint64_t fixed, ring[AB_STAGES], s, fin;
  uint32_t tmem_base;
};

// Issue discipline of the tcgen05.mma / TMA instructions (template parameter UNI of both kernels):
//   UNI = true : warp 0 runs the issue code with ALL lanes on warp-uniform values and only the issue itself is
//                gated by elect_one() (rf_sm100.cuh): descriptors stay in uniform registers, one MMA = UIADD3.64
//                pair + UTCHMMA.  These kernels issue M128 N64 K16 MMAs (32 tensor-pipe cycles each): with the
//   UNI = false: `if (tid == 0)` form (kept for A/B measurement, RF_UNIFORM_ISSUE=0) ptxas needs ~17 SASS
//                instructions (ELECT, 5 x R2UR, BRA.U.ANY loop) per MMA, several times its execution time.
template <bool UNI>
__device__ __forceinline__ bool issue_warp(int tid) { return UNI ? (tid < 32) : (tid == 0); }
template <bool UNI>
__device__ __forceinline__ bool issue_lane() { return UNI ? elect_one() : true; }

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// ------------------------------------------------------------------------------------------ kernel A: dQ (+ D)
template <bool UNI>
__global__ void __launch_bounds__(128, 2)
sr_attention_bwd_dq_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_do,
                           const __grid_constant__ CUtensorMap tm_kv, const __nv_bfloat16* __restrict__ o,
                           const __nv_bfloat16* __restrict__ dout, const float* __restrict__ lse,
                           float* __restrict__ dvec, __nv_bfloat16* __restrict__ dq, int N, int M, int heads,
                           float scale, float scale_log2) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;
  uint8_t* sdO = smem + AB_T128;
  uint8_t* ring = smem + 2 * AB_T128;  // stage s: K at ring + s*2*T64, V right after
  BwdBars* bars = reinterpret_cast<BwdBars*>(smem + 2 * AB_T128 + AB_STAGES * 2 * AB_T64 + 1024);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int q0 = blockIdx.x * 128, head = blockIdx.y, b = blockIdx.z;
  const int C = heads * AB_D;
  const int nchunks = (M + 63) / 64;

  if (tid == 0) {
    mbar_init(&bars->fixed, 1);
    for (int s = 0; s < AB_STAGES; ++s) mbar_init(&bars->ring[s], 1);
    mbar_init(&bars->s, 1);
    mbar_init(&bars->fin, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<256>(&bars->tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = UNI ? uniform_u32(bars->tmem_base) : bars->tmem_base;
  const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
  // columns: S [0,64) | dP [64,128) | dS [128,160) | dQ [160,224)
  const uint32_t tS = tmem + lane_off, tdP = tmem + lane_off + 64, tdS = tmem + lane_off + 128,
                 tdQ = tmem + lane_off + 160;

  if (issue_warp<UNI>(tid) && issue_lane<UNI>()) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_do);
    tma_prefetch_desc(&tm_kv);
    mbar_expect_tx(&bars->fixed, 2 * AB_T128);
    tma_load_3d(sQ, &tm_q, &bars->fixed, head * AB_D, q0, b);
    tma_load_3d(sdO, &tm_do, &bars->fixed, head * AB_D, q0, b);
    for (int s = 0; s < AB_STAGES && s < nchunks; ++s) {
      mbar_expect_tx(&bars->ring[s], 2 * AB_T64);
      tma_load_3d(ring + s * 2 * AB_T64, &tm_kv, &bars->ring[s], head * AB_D, s * 64, b);
      tma_load_3d(ring + s * 2 * AB_T64 + AB_T64, &tm_kv, &bars->ring[s], C + head * AB_D, s * 64, b);
    }
  }
  __syncwarp();

  // per-row constants: D = <dO, O>, LSE in log2 units
  const int row = q0 + tid;
  float Drow = 0.f, lse2 = 0.f;
  if (row < N) {
    const uint4* po = reinterpret_cast<const uint4*>(o + ((long)b * N + row) * C + head * AB_D);
    const uint4* pg = reinterpret_cast<const uint4*>(dout + ((long)b * N + row) * C + head * AB_D);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint4 a = __ldg(po + i), g = __ldg(pg + i);
      const __nv_bfloat162* ha = reinterpret_cast<const __nv_bfloat162*>(&a);
      const __nv_bfloat162* hg = reinterpret_cast<const __nv_bfloat162*>(&g);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 fa = __bfloat1622float2(ha[k]), fg = __bfloat1622float2(hg[k]);
        Drow = fmaf(fa.x, fg.x, Drow);
        Drow = fmaf(fa.y, fg.y, Drow);
      }
    }
    lse2 = __ldg(lse + ((long)b * heads + head) * N + row) * 1.44269504088896341f;
  }

  constexpr uint32_t IDESC_KK = make_idesc(FMT_BF16, 128, 64, 0, 0);  // both operands K-major
  constexpr uint32_t IDESC_MN = make_idesc(FMT_BF16, 128, 64, 0, 1);  // B MN-major
  const uint64_t descQ = make_sdesc_sw128(smem_u32(sQ), 16, 1024);
  const uint64_t descdO = make_sdesc_sw128(smem_u32(sdO), 16, 1024);

  for (int j = 0; j < nchunks; ++j) {
    const int st = j % AB_STAGES;
    uint8_t* sK = ring + st * 2 * AB_T64;
    uint8_t* sV = sK + AB_T64;
    if (issue_warp<UNI>(tid)) {
      if (j == 0) mbar_wait(&bars->fixed, 0);
      mbar_wait(&bars->ring[st], (j / AB_STAGES) & 1);
      tc_fence_after();
      const uint64_t descK = make_sdesc_sw128(smem_u32(sK), 16, 1024);
      const uint64_t descV = make_sdesc_sw128(smem_u32(sV), 16, 1024);
      if (issue_lane<UNI>()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) mma_f16_ss(tmem, descQ + (uint64_t)(k * 2), descK + (uint64_t)(k * 2), IDESC_KK, k > 0);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          mma_f16_ss(tmem + 64, descdO + (uint64_t)(k * 2), descV + (uint64_t)(k * 2), IDESC_KK, k > 0);
        tc_commit(&bars->s);
      }
    }
    __syncwarp();
    mbar_wait(&bars->s, j & 1);
    tc_fence_after();
    // S_j and dP_j are complete, hence (in-order tensor pipe) so is dQ_{j-1}: its K / V stage is free
    if (issue_warp<UNI>(tid) && j >= 1 && j - 1 + AB_STAGES < nchunks && issue_lane<UNI>()) {
      const int jn = j - 1 + AB_STAGES, s2 = (j - 1) % AB_STAGES;
      mbar_expect_tx(&bars->ring[s2], 2 * AB_T64);
      tma_load_3d(ring + s2 * 2 * AB_T64, &tm_kv, &bars->ring[s2], head * AB_D, jn * 64, b);
      tma_load_3d(ring + s2 * 2 * AB_T64 + AB_T64, &tm_kv, &bars->ring[s2], C + head * AB_D, jn * 64, b);
    }
    __syncwarp();
    const int kvalid = M - j * 64;
    auto softmax_grad = [&](auto masked) {
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t sv[32], dv[32];
        tmem_ld32(tS + c * 32, sv);
        tmem_ld32(tdP + c * 32, dv);
        tc_wait_ld();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float p0 = fast_exp2(fmaf(__uint_as_float(sv[2 * i]), scale_log2, -lse2));
          float p1 = fast_exp2(fmaf(__uint_as_float(sv[2 * i + 1]), scale_log2, -lse2));
          if constexpr (decltype(masked)::value) {   // only the last chunk of a ragged M
            if (c * 32 + 2 * i >= kvalid) p0 = 0.f;
            if (c * 32 + 2 * i + 1 >= kvalid) p1 = 0.f;
          }
          const float d0 = p0 * (__uint_as_float(dv[2 * i]) - Drow) * scale;
          const float d1 = p1 * (__uint_as_float(dv[2 * i + 1]) - Drow) * scale;
          pk[i] = pack_bf16(d0, d1);
        }
        tmem_st16(tdS + c * 16, pk);
      }
    };
    if (kvalid >= 64)
      softmax_grad(std::false_type{});
    else
      softmax_grad(std::true_type{});
    tc_wait_st();
    tc_fence_before();
    __syncthreads();
    if (issue_warp<UNI>(tid)) {
      tc_fence_after();
      const uint64_t descKmn = make_sdesc_sw128(smem_u32(sK), 8192, 1024);
      if (issue_lane<UNI>()) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          mma_f16_ts(tmem + 160, tmem + 128 + k * 8, descKmn + (uint64_t)(k * 128), IDESC_MN, (j > 0 || k > 0) ? 1u : 0u);
        if (j == nchunks - 1) tc_commit(&bars->fin);
      }
    }
    __syncwarp();
  }

  mbar_wait(&bars->fin, 0);
  tc_fence_after();
  uint32_t packed[32];
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    uint32_t v[32];
    tmem_ld32(tdQ + c * 32, v);
    tc_wait_ld();
#pragma unroll
    for (int i = 0; i < 16; ++i) packed[c * 16 + i] = pack_bf16(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
  }
  if (row < N) {
    uint4* dst = reinterpret_cast<uint4*>(dq + ((long)b * N + row) * C + head * AB_D);
#pragma unroll
    for (int i = 0; i < 8; ++i) dst[i] = make_uint4(packed[4 * i], packed[4 * i + 1], packed[4 * i + 2], packed[4 * i + 3]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tmem);
}

// 32 columns (queries) of one key row: P^T = exp2(s * c - lse2[q]), dS^T = scale * P^T * (dP^T - D[q]), packed bf16
__device__ __forceinline__ void dkv_half(const uint32_t (&sv)[32], const uint32_t (&dv)[32], const float4* l4,
                                         const float4* d4, float scale, float scale_log2, uint32_t (&pp)[16],
                                         uint32_t (&pd)[16]) {
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    const float4 l = l4[g], dd = d4[g];
    const float p0 = fast_exp2(fmaf(__uint_as_float(sv[4 * g + 0]), scale_log2, -l.x));
    const float p1 = fast_exp2(fmaf(__uint_as_float(sv[4 * g + 1]), scale_log2, -l.y));
    const float p2 = fast_exp2(fmaf(__uint_as_float(sv[4 * g + 2]), scale_log2, -l.z));
    const float p3 = fast_exp2(fmaf(__uint_as_float(sv[4 * g + 3]), scale_log2, -l.w));
    pp[2 * g] = pack_bf16(p0, p1);
    pp[2 * g + 1] = pack_bf16(p2, p3);
    pd[2 * g] = pack_bf16(p0 * (__uint_as_float(dv[4 * g + 0]) - dd.x) * scale,
                          p1 * (__uint_as_float(dv[4 * g + 1]) - dd.y) * scale);
    pd[2 * g + 1] = pack_bf16(p2 * (__uint_as_float(dv[4 * g + 2]) - dd.z) * scale,
                              p3 * (__uint_as_float(dv[4 * g + 3]) - dd.w) * scale);
  }
}

// ------------------------------------------------------------------------------------------ kernel B: dK, dV
template <bool UNI>
__global__ void __launch_bounds__(128, 2)
sr_attention_bwd_dkv_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_do,
                            const __grid_constant__ CUtensorMap tm_kv, const float* __restrict__ lse,
                            const float* __restrict__ dvec, float* __restrict__ dkv_acc, int N, int M, int heads,
                            int tiles_per_split, float scale, float scale_log2) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sK = smem;
  uint8_t* sV = smem + AB_T128;
  uint8_t* ring = smem + 2 * AB_T128;  // stage s: Q tile at ring + s*2*T64, dO tile right after
  float* sLse = reinterpret_cast<float*>(smem + 2 * AB_T128 + AB_STAGES * 2 * AB_T64);  // [2][64]
  float* sD = sLse + 128;                                                                // [2][64]
  BwdBars* bars = reinterpret_cast<BwdBars*>(smem + 2 * AB_T128 + AB_STAGES * 2 * AB_T64 + 1024);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int split = blockIdx.x, kv0 = blockIdx.y * 128;
  const int b = blockIdx.z / heads, head = blockIdx.z % heads;
  const int C = heads * AB_D;
  const int ntiles_all = (N + 63) / 64;
  const int t_begin = split * tiles_per_split;
  const int t_end = min(ntiles_all, t_begin + tiles_per_split);
  const int ntiles = t_end - t_begin;
  if (ntiles <= 0) return;  // uniform per CTA

  if (tid == 0) {
    mbar_init(&bars->fixed, 1);
    for (int s = 0; s < AB_STAGES; ++s) mbar_init(&bars->ring[s], 1);
    mbar_init(&bars->s, 1);
    mbar_init(&bars->fin, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<256>(&bars->tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = UNI ? uniform_u32(bars->tmem_base) : bars->tmem_base;
  const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
  // columns: S^T / P^T [0,64) | dP^T / dS^T [64,128) | dK [128,192) | dV [192,256)
  const uint32_t tS = tmem + lane_off, tdP = tmem + lane_off + 64, tdK = tmem + lane_off + 128,
                 tdV = tmem + lane_off + 192;

  if (issue_warp<UNI>(tid) && issue_lane<UNI>()) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_do);
    tma_prefetch_desc(&tm_kv);
    mbar_expect_tx(&bars->fixed, 2 * AB_T128);
    tma_load_3d(sK, &tm_kv, &bars->fixed, head * AB_D, kv0, b);
    tma_load_3d(sV, &tm_kv, &bars->fixed, C + head * AB_D, kv0, b);
    for (int s = 0; s < AB_STAGES && s < ntiles; ++s) {
      mbar_expect_tx(&bars->ring[s], 2 * AB_T64);
      tma_load_3d(ring + s * 2 * AB_T64, &tm_q, &bars->ring[s], head * AB_D, (t_begin + s) * 64, b);
      tma_load_3d(ring + s * 2 * AB_T64 + AB_T64, &tm_do, &bars->ring[s], head * AB_D, (t_begin + s) * 64, b);
    }
  }
  __syncwarp();
  const float* lse_bh = lse + ((long)b * heads + head) * N;
  const float* d_bh = dvec + ((long)b * heads + head) * N;
  {  // per-query vectors of the first tile
    const int qi = t_begin * 64 + (tid & 63);
    const float v = (qi < N) ? (tid < 64 ? __ldg(lse_bh + qi) * 1.44269504088896341f : __ldg(d_bh + qi)) : 0.f;
    (tid < 64 ? sLse : sD)[tid & 63] = v;
  }
  __syncthreads();

  constexpr uint32_t IDESC_KK = make_idesc(FMT_BF16, 128, 64, 0, 0);
  constexpr uint32_t IDESC_MN = make_idesc(FMT_BF16, 128, 64, 0, 1);
  const uint64_t descK = make_sdesc_sw128(smem_u32(sK), 16, 1024);
  const uint64_t descV = make_sdesc_sw128(smem_u32(sV), 16, 1024);

  for (int i = 0; i < ntiles; ++i) {
    const int st = i % AB_STAGES, vb = i & 1;
    uint8_t* sQ = ring + st * 2 * AB_T64;
    uint8_t* sdO = sQ + AB_T64;
    if (issue_warp<UNI>(tid)) {
      if (i == 0) mbar_wait(&bars->fixed, 0);
      mbar_wait(&bars->ring[st], (i / AB_STAGES) & 1);
      tc_fence_after();
      const uint64_t descQ = make_sdesc_sw128(smem_u32(sQ), 16, 1024);
      const uint64_t descdO = make_sdesc_sw128(smem_u32(sdO), 16, 1024);
      if (issue_lane<UNI>()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) mma_f16_ss(tmem, descK + (uint64_t)(k * 2), descQ + (uint64_t)(k * 2), IDESC_KK, k > 0);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          mma_f16_ss(tmem + 64, descV + (uint64_t)(k * 2), descdO + (uint64_t)(k * 2), IDESC_KK, k > 0);
        tc_commit(&bars->s);
      }
    }
    __syncwarp();
    if (i + 1 < ntiles) {  // per-query vectors of the next tile (published by this iteration's __syncthreads)
      const int qi = (t_begin + i + 1) * 64 + (tid & 63);
      const float v = (qi < N) ? (tid < 64 ? __ldg(lse_bh + qi) * 1.44269504088896341f : __ldg(d_bh + qi)) : 0.f;
      (tid < 64 ? sLse : sD)[(vb ^ 1) * 64 + (tid & 63)] = v;
    }
    mbar_wait(&bars->s, i & 1);
    tc_fence_after();
    if (issue_warp<UNI>(tid) && i >= 1 && i - 1 + AB_STAGES < ntiles && issue_lane<UNI>()) {  // stage of tile i-1 is free (in-order tensor pipe)
      const int tn = t_begin + i - 1 + AB_STAGES, s2 = (i - 1) % AB_STAGES;
      mbar_expect_tx(&bars->ring[s2], 2 * AB_T64);
      tma_load_3d(ring + s2 * 2 * AB_T64, &tm_q, &bars->ring[s2], head * AB_D, tn * 64, b);
      tma_load_3d(ring + s2 * 2 * AB_T64 + AB_T64, &tm_do, &bars->ring[s2], head * AB_D, tn * 64, b);
    }
    __syncwarp();
    const float4* l4 = reinterpret_cast<const float4*>(sLse + vb * 64);
    const float4* d4 = reinterpret_cast<const float4*>(sD + vb * 64);
    {
      // P^T / dS^T alias the first 32 columns of S^T / dP^T: the whole row of this lane is read first
      uint32_t s0[32], s1[32], g0[32], g1[32];
      tmem_ld32(tS, s0);
      tmem_ld32(tS + 32, s1);
      tmem_ld32(tdP, g0);
      tmem_ld32(tdP + 32, g1);
      tc_wait_ld();
      uint32_t pp[16], pd[16];
      dkv_half(s0, g0, l4, d4, scale, scale_log2, pp, pd);
      tmem_st16(tS, pp);
      tmem_st16(tdP, pd);
      dkv_half(s1, g1, l4 + 8, d4 + 8, scale, scale_log2, pp, pd);
      tmem_st16(tS + 16, pp);
      tmem_st16(tdP + 16, pd);
    }
    tc_wait_st();
    tc_fence_before();
    __syncthreads();
    if (issue_warp<UNI>(tid)) {
      tc_fence_after();
      const uint64_t descdOmn = make_sdesc_sw128(smem_u32(sdO), 8192, 1024);
      const uint64_t descQmn = make_sdesc_sw128(smem_u32(sQ), 8192, 1024);
      if (issue_lane<UNI>()) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          mma_f16_ts(tmem + 192, tmem + k * 8, descdOmn + (uint64_t)(k * 128), IDESC_MN, (i > 0 || k > 0) ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          mma_f16_ts(tmem + 128, tmem + 64 + k * 8, descQmn + (uint64_t)(k * 128), IDESC_MN, (i > 0 || k > 0) ? 1u : 0u);
        if (i == ntiles - 1) tc_commit(&bars->fin);
      }
    }
    __syncwarp();
  }

  mbar_wait(&bars->fin, 0);
  tc_fence_after();
  const int kv = kv0 + tid;
  float* dst = dkv_acc + ((long)b * M + kv) * 2 * C + head * AB_D;
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    uint32_t vk[32], vv[32];
    tmem_ld32(tdK + c * 32, vk);
    tmem_ld32(tdV + c * 32, vv);
    tc_wait_ld();
    if (kv < M) {
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        atomicAdd(dst + c * 32 + i, __uint_as_float(vk[i]));
        atomicAdd(dst + C + c * 32 + i, __uint_as_float(vv[i]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tmem);
}

// D[b,h,n] = sum_d dO[b,n,h*64+d] * O[b,n,h*64+d] (same operation order as the dQ kernel's prologue), written
// before the two backward kernels are forked onto separate streams (the dK/dV kernel reads it).
__global__ void __launch_bounds__(256)
sr_attention_dvec_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ dout,
                         float* __restrict__ dvec, int N, int heads, long total) {
  const long idx = blockIdx.x * (long)blockDim.x + threadIdx.x;   // (b, head, row)
  if (idx >= total) return;
  const int row = (int)(idx % N);
  const long bh = idx / N;
  const int head = (int)(bh % heads);
  const long b = bh / heads;
  const int C = heads * AB_D;
  const uint4* po = reinterpret_cast<const uint4*>(o + (b * N + row) * C + head * AB_D);
  const uint4* pg = reinterpret_cast<const uint4*>(dout + (b * N + row) * C + head * AB_D);
  float d = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const uint4 a = __ldg(po + i), g = __ldg(pg + i);
    const __nv_bfloat162* ha = reinterpret_cast<const __nv_bfloat162*>(&a);
    const __nv_bfloat162* hg = reinterpret_cast<const __nv_bfloat162*>(&g);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 fa = __bfloat1622float2(ha[k]), fg = __bfloat1622float2(hg[k]);
      d = fmaf(fa.x, fg.x, d);
      d = fmaf(fa.y, fg.y, d);
    }
  }
  dvec[idx] = d;
}

// Side stream + fork / join events per device: the dQ kernel (320 CTAs on 296 resident slots at the dominant
// stage-3 shape = two rounds, the second nearly empty) and the dK/dV kernel (one wave by construction) are
// independent once D is known, so they run concurrently and fill each other's tail.  cudaStreamWaitEvent
// fork / join is legal inside a stream capture, so the CUDA-graph replay keeps the two branches parallel.
struct BwdFork {
  cudaStream_t side = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
};
static BwdFork* bwd_fork() {
  static BwdFork forks[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  BwdFork& f = forks[dev];
  if (f.side == nullptr) {
    if (cudaStreamCreateWithFlags(&f.side, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&f.fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&f.join, cudaEventDisableTiming) != cudaSuccess) return nullptr;
  }
  return &f;
}

}  // namespace rf

using namespace rf;

// warp-specialised single-launch backward (sr_attention_bwd_ws.cu) -- the default; RF_ATTN_BWD=old selects the two
// single-role kernels of this file (kept for A/B measurement)
int64_t rf_sr_attention_bwd_ws_workspace_bytes(int B, int N, int heads);
int rf_sr_attention_bwd_ws(const void* q, const void* kv, const void* out, const void* grad_out, const float* lse,
                           void* grad_q, float* grad_kv_f32, void* workspace, int B, int N, int M, int heads, float scale,
                           cudaStream_t st);
static bool use_ws_backward() {
  static const bool ws = [] { const char* e = getenv("RF_ATTN_BWD"); return !(e && e[0] == 'o'); }();
  return ws;
}

extern "C" int64_t rf_sr_attention_bwd_workspace_bytes(int B, int N, int M, int heads) {
  (void)M;
  const int64_t old_bytes = (int64_t)sizeof(float) * (int64_t)B * heads * N;  // D[b,h,n] = rowsum(dO * O)
  const int64_t ws_bytes = rf_sr_attention_bwd_ws_workspace_bytes(B, N, heads);
  return old_bytes > ws_bytes ? old_bytes : ws_bytes;
}

extern "C" int rf_sr_attention_bwd(const void* q, const void* kv, const void* out, const void* grad_out,
                                   const float* lse, void* grad_q, float* grad_kv_f32, void* workspace, int B, int N,
                                   int M, int heads, float scale, void* stream) {
  RF_REQUIRE(q && kv && out && grad_out && lse && grad_q && grad_kv_f32 && workspace, "rf_sr_attention_bwd: null pointer");
  RF_REQUIRE(B > 0 && N > 0 && M > 0 && heads > 0 && B <= 65535 && heads <= 65535 && (long)B * heads <= 65535,
             "rf_sr_attention_bwd: bad shape");
  const int C = heads * AB_D;
  cudaStream_t st = (cudaStream_t)stream;
  RF_REQUIRE(((uintptr_t)workspace & 15) == 0 && ((uintptr_t)grad_q & 15) == 0 && ((uintptr_t)grad_kv_f32 & 15) == 0,
             "rf_sr_attention_bwd: workspace / gradients must be 16-byte aligned");
  if (use_ws_backward())
    return rf_sr_attention_bwd_ws(q, kv, out, grad_out, lse, grad_q, grad_kv_f32, workspace, B, N, M, heads, scale, st);
  float* dvec = (float*)workspace;
  CUtensorMap tq128, tdo128, tkv64, tq64, tdo64, tkv128;
  int rc;
#define RF_TM(map, base, d0, d1, box1)                                                                         \
  rc = make_tmap_3d(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, (uint64_t)(d0), (uint64_t)(d1), (uint64_t)B, \
                    (uint64_t)(d0) * 2, (uint64_t)(d1) * (d0) * 2, AB_D, box1);                                \
  if (rc != RF_OK) return rc
  RF_TM(tq128, q, C, N, 128);
  RF_TM(tdo128, grad_out, C, N, 128);
  RF_TM(tkv64, kv, 2 * C, M, 64);
  RF_TM(tq64, q, C, N, 64);
  RF_TM(tdo64, grad_out, C, N, 64);
  RF_TM(tkv128, kv, 2 * C, M, 128);
#undef RF_TM
  static bool attr_set = false;
  if (!attr_set) {
    RF_CUDA(cudaFuncSetAttribute(sr_attention_bwd_dq_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM));
    RF_CUDA(cudaFuncSetAttribute(sr_attention_bwd_dkv_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM));
    RF_CUDA(cudaFuncSetAttribute(sr_attention_bwd_dq_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM));
    RF_CUDA(cudaFuncSetAttribute(sr_attention_bwd_dkv_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM));
    attr_set = true;
  }
  const float scale_log2 = scale * 1.44269504088896341f;
  {
    const long total = (long)B * heads * N;
    sr_attention_dvec_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
        (const __nv_bfloat16*)out, (const __nv_bfloat16*)grad_out, dvec, N, heads, total);
    RF_CHECK_LAUNCH("sr_attention_dvec_kernel");
  }
  RF_CUDA(cudaMemsetAsync(grad_kv_f32, 0, sizeof(float) * (size_t)B * M * 2 * C, st));
  static const bool uni = [] { const char* e = getenv("RF_UNIFORM_ISSUE"); return !(e && e[0] == '0'); }();
  static const bool serial = [] { const char* e = getenv("RF_ATTN_BWD_SERIAL"); return e && e[0] == '1'; }();
  BwdFork* fk = serial ? nullptr : bwd_fork();
  cudaStream_t st_kv = st;
  if (fk != nullptr) {   // fork: the dK/dV kernel goes to the side stream
    RF_CUDA(cudaEventRecord(fk->fork, st));
    RF_CUDA(cudaStreamWaitEvent(fk->side, fk->fork, 0));
    st_kv = fk->side;
  }
  {
    dim3 grid((unsigned)((N + 127) / 128), (unsigned)heads, (unsigned)B);
    auto kern = uni ? sr_attention_bwd_dq_kernel<true> : sr_attention_bwd_dq_kernel<false>;
    kern<<<grid, 128, AB_SMEM, st>>>(tq128, tdo128, tkv64, (const __nv_bfloat16*)out, (const __nv_bfloat16*)grad_out, lse,
                                     dvec, (__nv_bfloat16*)grad_q, N, M, heads, scale, scale_log2);
    RF_CHECK_LAUNCH("sr_attention_bwd_dq_kernel");
  }
  {
    const int kvchunks = (M + 127) / 128;
    const int ntiles = (N + 63) / 64;
    // query splits so that ONE wave of CTAs (2 per SM) covers the work -- rounding up would leave a few
    // CTAs for a second wave and nearly double the kernel time; at least 4 tiles per split
    long base = (long)B * heads * kvchunks;
    int splits = (int)((2l * kNumSMs) / base);
    if (splits > ntiles / 4) splits = ntiles / 4;
    if (splits < 1) splits = 1;
    const int tps = (ntiles + splits - 1) / splits;
    splits = (ntiles + tps - 1) / tps;
    dim3 grid((unsigned)splits, (unsigned)kvchunks, (unsigned)(B * heads));
    auto kern = uni ? sr_attention_bwd_dkv_kernel<true> : sr_attention_bwd_dkv_kernel<false>;
    kern<<<grid, 128, AB_SMEM, st_kv>>>(tq64, tdo64, tkv128, lse, dvec, grad_kv_f32, N, M, heads, tps, scale, scale_log2);
    RF_CHECK_LAUNCH("sr_attention_bwd_dkv_kernel");
  }
  if (fk != nullptr) {   // join
    RF_CUDA(cudaEventRecord(fk->join, fk->side));
    RF_CUDA(cudaStreamWaitEvent(st, fk->join, 0));
  }
  return RF_OK;
}
