// Persistent, warp-specialised tcgen05 (UMMA) TF32 global correlation for sm_100a -- the sweep-size path.
//
// Same mathematics and the same three recompute phases as global_corr_umma.cu (GlobalFeatureCorrelationLayer,
// /root/reference/models/modules.py:294-333,362-374: corr = src^T trg, mutual matching with the row / column
// maxima, ReLU, L2-norm over the source dimension), restructured so that the tensor pipe, TMA and the epilogue
// overlap inside ONE CTA per SM instead of relying on two co-resident single-tile CTAs:
//   * every CTA owns a CONTIGUOUS range of the (batch, s-tile, t-tile) tile sequence (t fastest), so the 128-row
//     source operand (C x 128 fp32 = 64 KB at C = 128) is loaded ONCE per s-tile and stays resident in shared
//     memory while only the target operand streams through a 6-stage ring of 16 KB K-blocks: operand traffic
//     per 128x128 tile drops from 128 KB to 64 KB (the old kernel was L2->SM bound: 32 flop per operand byte);
//   * warp 9 = TMA producer, warp 8 = MMA issuer, warps 0-7 = epilogue (TMEM lane quarter = warp % 4, column
//     half = warp / 4); four 128-column fp32 accumulators in TMEM (all 512 columns) decouple the MMA of tile
//     j+1..j+3 from the epilogue of tile j;
//   * phase 2 transposes each 32x32 block through a swizzled warp-private shared buffer, so a warp stores four
//     full 128-byte lines per instruction (the old kernel wrote 16-byte pieces of 32 different rows);
//   * the row maximum lives in a register for a whole s-tile; the column reductions go straight to global
//     memory with one coalesced 32-wide RED per warp and 32-column block.
// Operand layouts, descriptors and the instruction descriptor are those of global_corr_umma.cu (MN-major tf32,
// 128-byte swizzle with 32-byte atoms, boxes of 32 positions x 32 channels).
#include "rf_common.cuh"
#include "rf_sm100.cuh"

namespace rf {
using namespace sm100;

constexpr int GP_BK = 32;                        // channels per K block
constexpr int GP_MAXKB = 4;                      // C <= 128: the source tile stays resident
constexpr int GP_BOX_BYTES = GP_BK * 128;        // 32 positions x 32 channels
constexpr int GP_KB_BYTES = 4 * GP_BOX_BYTES;    // 128 positions x 32 channels = 16 KB
constexpr int GP_STAGES = 6;
constexpr int GP_ACC = 4;                        // TMEM accumulators (128 columns each)
constexpr int GP_EPI_WARPS = 8;
constexpr int GP_THREADS = (GP_EPI_WARPS + 2) * 32;
constexpr int GP_STAGE_BYTES = 32 * 128;         // one warp's 32x32 fp32 transpose buffer
constexpr int GP_SMEM_A = GP_MAXKB * GP_KB_BYTES;
constexpr int GP_SMEM_RING = GP_STAGES * GP_KB_BYTES;
constexpr int GP_SMEM_STAGE = GP_EPI_WARPS * GP_STAGE_BYTES;
constexpr int GP_SMEM = GP_SMEM_A + GP_SMEM_RING + GP_SMEM_STAGE + 1024 /* alignment */ + 5120 /* GpShared */;

struct __align__(16) GpShared {
  float fac[GP_EPI_WARPS][2][64];   // per epilogue warp: 1/(colmax+eps) and 1/norm of its 64 columns
  uint64_t a_full, a_empty, b_full[GP_STAGES], b_empty[GP_STAGES], acc_full[GP_ACC], acc_empty[GP_ACC];
  uint32_t tmem_base;
};
static_assert(sizeof(GpShared) <= 5120, "GpShared too large");

__device__ __forceinline__ void gp_atomic_max_float(float* addr, float v) {
  if (v >= 0.f)
    atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else
    atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

// Reduction ACROSS the 32 lanes of a warp of a 32-element per-lane array: on return lane l holds op_{lanes} w[l].
template <bool MAX>
__device__ __forceinline__ float gp_transpose_reduce(float (&w)[32], int lane) {
#pragma unroll
  for (int h = 16; h >= 1; h >>= 1) {
    const bool up = (lane & h) != 0;
#pragma unroll
    for (int i = 0; i < h; ++i) {
      const float send = up ? w[i] : w[i + h];
      const float keep = up ? w[i + h] : w[i];
      const float r = __shfl_xor_sync(0xffffffffu, send, h);
      w[i] = MAX ? fmaxf(keep, r) : keep + r;
    }
  }
  return w[0];
}

struct GpTile {
  long row;      // b * nST + s_tile
  int b, s_tile, t_tile;
};
__device__ __forceinline__ GpTile gp_decode(long i, int nST, int nTT) {
  GpTile t;
  t.row = i / nTT;
  t.t_tile = (int)(i - t.row * nTT);
  t.b = (int)(t.row / nST);
  t.s_tile = (int)(t.row - (long)t.b * nST);
  return t;
}

// one 32-column block of one accumulator row (thread = row): PHASE 0 / 1 reductions or the PHASE 2 store
template <int PHASE>
__device__ __forceinline__ void gp_block(const uint32_t (&v)[32], const float* __restrict__ fcb,
                                         const float* __restrict__ fcn, uint8_t* stage, int lane, bool row_ok,
                                         float ra, bool mm, bool nrm, long col0, long Nt, float& rmax,
                                         float* __restrict__ colred, float* __restrict__ out_rows0, long s_base,
                                         long Ns) {
  if (PHASE == 0) {
    float w[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const float x = __uint_as_float(v[i]);
      if (col0 + i < Nt) rmax = fmaxf(rmax, x);
      w[i] = row_ok ? x : -INFINITY;
    }
    const float cm = gp_transpose_reduce<true>(w, lane);   // lane l: max over this warp's 32 rows of column l
    if (col0 + lane < Nt) gp_atomic_max_float(colred + col0 + lane, cm);
    return;
  }
  float w[32];
  const float2 ra2 = make_float2(ra, ra);
#pragma unroll
  for (int i4 = 0; i4 < 8; ++i4) {
    const float4 cb = *reinterpret_cast<const float4*>(fcb + 4 * i4);   // warp-uniform address: broadcast
    float2 x01 = make_float2(__uint_as_float(v[4 * i4]), __uint_as_float(v[4 * i4 + 1]));
    float2 x23 = make_float2(__uint_as_float(v[4 * i4 + 2]), __uint_as_float(v[4 * i4 + 3]));
    if (mm) {   // c * ((c / (A + eps)) * (c / (Bm + eps)))
      x01 = __fmul2_rn(x01, __fmul2_rn(__fmul2_rn(x01, ra2), __fmul2_rn(x01, make_float2(cb.x, cb.y))));
      x23 = __fmul2_rn(x23, __fmul2_rn(__fmul2_rn(x23, ra2), __fmul2_rn(x23, make_float2(cb.z, cb.w))));
    }
    if (nrm) {
      x01.x = fmaxf(x01.x, 0.f); x01.y = fmaxf(x01.y, 0.f);
      x23.x = fmaxf(x23.x, 0.f); x23.y = fmaxf(x23.y, 0.f);
    }
    if (PHASE == 1) {
      x01 = __fmul2_rn(x01, x01);
      x23 = __fmul2_rn(x23, x23);
    } else {
      const float4 cn = *reinterpret_cast<const float4*>(fcn + 4 * i4);
      x01 = __fmul2_rn(x01, make_float2(cn.x, cn.y));
      x23 = __fmul2_rn(x23, make_float2(cn.z, cn.w));
    }
    w[4 * i4] = x01.x; w[4 * i4 + 1] = x01.y; w[4 * i4 + 2] = x23.x; w[4 * i4 + 3] = x23.y;
  }
  if (PHASE == 1) {
#pragma unroll
    for (int i = 0; i < 32; ++i) w[i] = row_ok ? w[i] : 0.f;
    const float cs = gp_transpose_reduce<false>(w, lane);
    if (col0 + lane < Nt) atomicAdd(colred + col0 + lane, cs);
    return;
  }
  // PHASE 2: transpose through the warp-private buffer (16-byte chunk j of row r lives at chunk j ^ (r & 7)), then
  // each group of 8 lanes stores one full 128-byte line: 4 rows per instruction
  __syncwarp();   // the previous block's reads of the buffer are complete
#pragma unroll
  for (int j = 0; j < 8; ++j)
    *reinterpret_cast<float4*>(stage + lane * 128 + ((j ^ (lane & 7)) << 4)) =
        make_float4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
  __syncwarp();
  const int j = lane & 7;
  const bool col_ok = col0 + 4 * j < Nt;   // Nt % 4 == 0: a 16-byte piece is entirely inside or outside
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int r = it * 4 + (lane >> 3);
    const float4 val = *reinterpret_cast<const float4*>(stage + r * 128 + ((j ^ (r & 7)) << 4));
    if (col_ok && s_base + r < Ns) st_cs_f4(out_rows0 + (long)r * Nt + col0 + 4 * j, val);
  }
}

template <int PHASE>
__global__ void __launch_bounds__(GP_THREADS, 1)
global_corr_persist_kernel(const __grid_constant__ CUtensorMap tm_src, const __grid_constant__ CUtensorMap tm_trg,
                           float* __restrict__ out, float* __restrict__ rowmax, float* __restrict__ colmax,
                           float* __restrict__ normsq, int C, long Ns, long Nt, int nST, int nTT, long total,
                           int mode) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* ring = smem + GP_SMEM_A;
  uint8_t* stage_all = ring + GP_SMEM_RING;
  GpShared* sh = reinterpret_cast<GpShared*>(stage_all + GP_SMEM_STAGE);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nkb = C / GP_BK;
  const bool mm = mode & 1, nrm = mode & 2;
  const long start = (long)blockIdx.x * total / gridDim.x;
  const long end = (long)(blockIdx.x + 1) * total / gridDim.x;

  if (tid == 0) {
    mbar_init(&sh->a_full, 1);
    mbar_init(&sh->a_empty, 1);
    for (int s = 0; s < GP_STAGES; ++s) {
      mbar_init(&sh->b_full[s], 1);
      mbar_init(&sh->b_empty[s], 1);
    }
    for (int a = 0; a < GP_ACC; ++a) {
      mbar_init(&sh->acc_full[a], 1);
      mbar_init(&sh->acc_empty[a], GP_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 8) tmem_alloc<512>(&sh->tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sh->tmem_base;

  if (warp == 9) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      tma_prefetch_desc(&tm_src);
      tma_prefetch_desc(&tm_trg);
      long prev_row = -1;
      uint32_t a_loads = 0, bk = 0;
      for (long i = start; i < end; ++i) {
        const GpTile t = gp_decode(i, nST, nTT);
        if (t.row != prev_row) {
          if (a_loads > 0) mbar_wait(&sh->a_empty, (a_loads - 1) & 1);   // every MMA reading the old tile is done
          mbar_expect_tx(&sh->a_full, (uint32_t)nkb * GP_KB_BYTES);
          for (int kb = 0; kb < nkb; ++kb)
#pragma unroll
            for (int blk = 0; blk < 4; ++blk)
              tma_load_3d(sA + kb * GP_KB_BYTES + blk * GP_BOX_BYTES, &tm_src, &sh->a_full,
                          t.s_tile * 128 + blk * 32, kb * GP_BK, t.b);
          ++a_loads;
          prev_row = t.row;
        }
        for (int kb = 0; kb < nkb; ++kb, ++bk) {
          const uint32_t st = bk % GP_STAGES;
          if (bk >= GP_STAGES) mbar_wait(&sh->b_empty[st], ((bk / GP_STAGES) - 1) & 1);
          uint8_t* dst = ring + st * GP_KB_BYTES;
          mbar_expect_tx(&sh->b_full[st], GP_KB_BYTES);
#pragma unroll
          for (int blk = 0; blk < 4; ++blk)
            tma_load_3d(dst + blk * GP_BOX_BYTES, &tm_trg, &sh->b_full[st], t.t_tile * 128 + blk * 32, kb * GP_BK,
                        t.b);
        }
      }
    }
  } else if (warp == 8) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t IDESC = make_idesc(FMT_TF32, 128, 128, 1, 1);
      long prev_row = -1;
      uint32_t a_cnt = 0, bk = 0, j = 0;
      for (long i = start; i < end; ++i, ++j) {
        const long row = i / nTT;
        if (row != prev_row) {
          mbar_wait(&sh->a_full, a_cnt & 1);
          ++a_cnt;
          prev_row = row;
        }
        const uint32_t buf = j % GP_ACC;
        if (j >= GP_ACC) mbar_wait(&sh->acc_empty[buf], ((j / GP_ACC) - 1) & 1);   // epilogue drained this accumulator
        tc_fence_after();
        for (int kb = 0; kb < nkb; ++kb, ++bk) {
          const uint32_t st = bk % GP_STAGES;
          mbar_wait(&sh->b_full[st], (bk / GP_STAGES) & 1);
          tc_fence_after();
          const uint64_t da = make_sdesc_sw128_base32(smem_u32(sA + kb * GP_KB_BYTES), GP_BOX_BYTES, 512);
          const uint64_t db = make_sdesc_sw128_base32(smem_u32(ring + st * GP_KB_BYTES), GP_BOX_BYTES, 512);
#pragma unroll
          for (int k = 0; k < GP_BK / 8; ++k)
            mma_tf32_ss(tmem + buf * 128, da + (uint64_t)(k * 64), db + (uint64_t)(k * 64), IDESC,
                        (kb > 0 || k > 0) ? 1u : 0u);
          tc_commit(&sh->b_empty[st]);
        }
        tc_commit(&sh->acc_full[buf]);
        if (i + 1 < end && (i + 1) / nTT != row) tc_commit(&sh->a_empty);   // the resident source tile may be replaced
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps: thread = accumulator row
    const int q = warp & 3, hf = warp >> 2;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    uint8_t* stage = stage_all + warp * GP_STAGE_BYTES;
    float* fcb = sh->fac[warp][0];
    float* fcn = sh->fac[warp][1];
    long prev_row = -1, prev_s = 0;
    int prev_b = 0;
    bool row_ok = false;
    float rmax = -INFINITY, ra = 1.f;
    uint32_t j = 0;
    for (long i = start; i < end; ++i, ++j) {
      const GpTile t = gp_decode(i, nST, nTT);
      if (t.row != prev_row) {
        if (PHASE == 0 && prev_row >= 0 && row_ok) gp_atomic_max_float(rowmax + (long)prev_b * Ns + prev_s, rmax);
        prev_row = t.row;
        prev_b = t.b;
        prev_s = (long)t.s_tile * 128 + q * 32 + lane;
        row_ok = prev_s < Ns;
        rmax = -INFINITY;
        if (PHASE >= 1) ra = (mm && row_ok) ? 1.f / (rowmax[(long)t.b * Ns + prev_s] + 1e-5f) : 1.f;   // 1 / (A + eps)
      }
      const long colh = (long)t.t_tile * 128 + hf * 64;   // first column of this warp's half
      if (PHASE >= 1) {
        __syncwarp();   // the previous tile's reads of fac[] are complete
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const long col = colh + c * 32 + lane;
          fcb[c * 32 + lane] = (mm && col < Nt) ? 1.f / (colmax[(long)t.b * Nt + col] + 1e-5f) : 1.f;   // 1 / (Bm + eps)
          if (PHASE == 2)
            fcn[c * 32 + lane] = (nrm && col < Nt) ? 1.f / fmaxf(sqrtf(normsq[(long)t.b * Nt + col]), 1e-12f) : 1.f;
        }
        __syncwarp();
      }
      const uint32_t buf = j % GP_ACC;
      mbar_wait(&sh->acc_full[buf], (j / GP_ACC) & 1);
      tc_fence_after();
      uint32_t v0[32], v1[32];
      const uint32_t taddr = tmem + lane_off + buf * 128 + hf * 64;
      tmem_ld32(taddr, v0);
      tmem_ld32(taddr + 32, v1);
      tc_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sh->acc_empty[buf]);   // the accumulator is in registers: release it to the MMA warp
      float* colred = PHASE == 0 ? colmax + (long)t.b * Nt : normsq + (long)t.b * Nt;
      const long s_base = (long)t.s_tile * 128 + q * 32;
      float* out_rows0 = out + ((long)t.b * Ns + s_base) * Nt;
      gp_block<PHASE>(v0, fcb, fcn, stage, lane, row_ok, ra, mm, nrm, colh, Nt, rmax, colred, out_rows0, s_base, Ns);
      gp_block<PHASE>(v1, fcb + 32, fcn + 32, stage, lane, row_ok, ra, mm, nrm, colh + 32, Nt, rmax, colred, out_rows0,
                      s_base, Ns);
    }
    if (PHASE == 0 && prev_row >= 0 && row_ok) gp_atomic_max_float(rowmax + (long)prev_b * Ns + prev_s, rmax);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc<512>(tmem);
}

bool global_corr_persist_supported(int C, long Ns, long Nt, const void* a, const void* b, const void* c) {
  return C % GP_BK == 0 && C / GP_BK <= GP_MAXKB && Ns % 4 == 0 && Nt % 4 == 0 && Ns < (1l << 31) - 128 &&
         Nt < (1l << 31) - 128 && (((uintptr_t)a | (uintptr_t)b | (uintptr_t)c) & 15) == 0;
}

// rowmax / colmax: [B,Ns] / [B,Nt] pre-filled with -inf by the caller when mode & 1; normsq: [B,Nt] scratch.
int global_corr_persist(const float* src, const float* trg, float* out, float* rowmax, float* colmax, float* normsq,
                        int B, int C, long Ns, long Nt, int mode, cudaStream_t st) {
  CUtensorMap ts, tt;
  int rc = make_tmap_3d(&ts, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, src, (uint64_t)Ns, (uint64_t)C, (uint64_t)B,
                        (uint64_t)Ns * 4, (uint64_t)C * Ns * 4, 32, GP_BK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  if (rc != RF_OK) return rc;
  rc = make_tmap_3d(&tt, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, trg, (uint64_t)Nt, (uint64_t)C, (uint64_t)B,
                    (uint64_t)Nt * 4, (uint64_t)C * Nt * 4, 32, GP_BK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  if (rc != RF_OK) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    RF_CUDA(cudaFuncSetAttribute(global_corr_persist_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, GP_SMEM));
    RF_CUDA(cudaFuncSetAttribute(global_corr_persist_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, GP_SMEM));
    RF_CUDA(cudaFuncSetAttribute(global_corr_persist_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, GP_SMEM));
    attr_set = true;
  }
  const int nST = (int)ceil_div(Ns, 128), nTT = (int)ceil_div(Nt, 128);
  const long total = (long)B * nST * nTT;
  const unsigned grid = (unsigned)(total < kNumSMs ? total : kNumSMs);
  if (mode & 1) {
    global_corr_persist_kernel<0><<<grid, GP_THREADS, GP_SMEM, st>>>(ts, tt, out, rowmax, colmax, normsq, C, Ns, Nt, nST,
                                                                    nTT, total, mode);
    RF_CHECK_LAUNCH("global_corr_persist_kernel<0>");
  }
  if (mode & 2) {
    RF_CUDA(cudaMemsetAsync(normsq, 0, sizeof(float) * (size_t)B * Nt, st));
    global_corr_persist_kernel<1><<<grid, GP_THREADS, GP_SMEM, st>>>(ts, tt, out, rowmax, colmax, normsq, C, Ns, Nt, nST,
                                                                    nTT, total, mode);
    RF_CHECK_LAUNCH("global_corr_persist_kernel<1>");
  }
  global_corr_persist_kernel<2><<<grid, GP_THREADS, GP_SMEM, st>>>(ts, tt, out, rowmax, colmax, normsq, C, Ns, Nt, nST,
                                                                  nTT, total, mode);
  RF_CHECK_LAUNCH("global_corr_persist_kernel<2>");
  return RF_OK;
}

}  // namespace rf
