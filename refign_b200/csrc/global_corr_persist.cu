// Persistent, warp-specialised tcgen05 (UMMA) TF32 global correlation for sm_100a -- the sweep-size path.
//
// Same mathematics as global_corr_umma.cu (GlobalFeatureCorrelationLayer, /root/reference/models/modules.py:
// 294-333,362-374: corr = src^T trg, mutual matching with the row / column maxima, ReLU, L2-norm over the
// source dimension); the volume is written ONCE and the reductions the reference needs are obtained by
// recomputing tiles on the tensor cores.  Structure (one CTA per SM, 320 threads):
//   * every CTA owns a CONTIGUOUS range of the (batch, row-tile, column-tile) sequence (column tile fastest), so
//     the 128-row "resident" operand (C x 128 fp32 = 64 KB at C = 128) is loaded ONCE per row tile and stays in
//     shared memory while only the "streaming" operand passes through a 3-stage ring of 32 KB K-blocks:
//     operand traffic per 128x128 of output is 64 KB instead of 128 KB;
//   * tiles are 128 x 256 (tcgen05.mma M128 N256 K8): with N = 128 the single issuing thread needed more cycles
//     to ISSUE an MMA (descriptor arithmetic, R2UR moves, election: ~17 SASS instructions) than the tensor pipe
//     needed to EXECUTE it (67 cycles) -- ncu showed the issuer busy 100 % and the epilogue warps idle on the
//     accumulator barrier at 1.6 us per 128x128 tile, whatever the epilogue did;
//   * warp 9 = TMA producer (the boxes of a stage are issued by 8 lanes of ONE instruction), warp 8 = MMA issuer,
//     warps 0-7 = epilogue (TMEM lane quarter = warp % 4, column half = warp / 4); two 256-column fp32
//     accumulators in TMEM (all 512 columns) decouple the MMA of tile j+1 from the epilogue of tile j;
//   * EVERY reduction is a per-thread ROW reduction (thread = TMEM lane = accumulator row) carried in a register
//     across the whole row of tiles: the column maxima / column norms of the volume are the row maxima / row
//     norms of the TRANSPOSED product, which the same kernel computes with the operands swapped (both operands
//     are MN-major, M = N = 128: the instruction is symmetric).  The first version reduced columns with a
//     31-shuffle butterfly + global REDs per 32x32 block and was issue-bound at 7 500 warp instructions per tile
//     (ncu: tensor pipe 27 % busy); the row form needs ~1 instruction per element and no atomics per tile;
//   * the write pass transposes each 32x32 block through a swizzled warp-private shared buffer, so a warp stores
//     four full 128-byte lines per instruction.
// Passes (kernel template parameter):
//   ROWMAX  rows = A:  out_row[r] = max_c D[r][c]                      (run for (src,trg) and for (trg,src))
//   ROWSSQ  rows = trg, cols = src:  out_row[t] += sum_s relu(v)^2,  v = c * ((c * ra_s) * (c * cb_t))
//   WRITE   rows = src, cols = trg:  out[s][t] = relu(v) / max(||.||_t, 1e-12)
// Operand layouts, descriptors and the instruction descriptor are those of global_corr_umma.cu (MN-major tf32,
// 128-byte swizzle with 32-byte atoms, boxes of 32 positions x 32 channels).
#include "rf_common.cuh"
#include "rf_sm100.cuh"

namespace rf {
using namespace sm100;

constexpr int GP_BK = 32;                        // channels per K block
constexpr int GP_MAXKB = 4;                      // C <= 128: the resident tile fits
constexpr int GP_BOX_BYTES = GP_BK * 128;        // 32 positions x 32 channels
constexpr int GP_KB_BYTES = 4 * GP_BOX_BYTES;    // resident operand: 128 positions x 32 channels = 16 KB
constexpr int GP_TN = 256;                       // tile columns (MMA N)
constexpr int GP_BKB_BYTES = (GP_TN / 32) * GP_BOX_BYTES;   // streaming operand: 256 positions x 32 channels = 32 KB
constexpr int GP_STAGES = 3;
constexpr int GP_ACC = 2;                        // TMEM accumulators (256 columns each)
constexpr int GP_EPI_WARPS = 8;
constexpr int GP_THREADS = (GP_EPI_WARPS + 2) * 32;
constexpr int GP_STAGE_BYTES = 32 * 128;         // one warp's 32x32 fp32 transpose buffer
constexpr int GP_SMEM_A = GP_MAXKB * GP_KB_BYTES;
constexpr int GP_SMEM_RING = GP_STAGES * GP_BKB_BYTES;
constexpr int GP_SMEM_STAGE = GP_EPI_WARPS * GP_STAGE_BYTES;
constexpr int GP_SMEM = GP_SMEM_A + GP_SMEM_RING + GP_SMEM_STAGE + 1024 /* alignment */ + 9216 /* GpShared */;

enum { GP_ROWMAX = 0, GP_ROWSSQ = 1, GP_WRITE = 2 };

struct __align__(16) GpShared {
  float fac[GP_EPI_WARPS][2][GP_TN / 2];   // per epilogue warp: column factor and 1/norm of its 128 columns
  uint64_t a_full, a_empty, b_full[GP_STAGES], b_empty[GP_STAGES], acc_full[GP_ACC], acc_empty[GP_ACC];
  uint32_t tmem_base;
};
static_assert(sizeof(GpShared) <= 9216, "GpShared too large");

__device__ __forceinline__ void gp_atomic_max_float(float* addr, float v) {
  if (v >= 0.f)
    atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else
    atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

// walks the tile sequence (batch, row tile, column tile) without divisions after the first decode
struct GpTile {
  int b, r_tile, c_tile;
};
__device__ __forceinline__ GpTile gp_decode(long i, int nRT, int nCT) {
  GpTile t;
  const long row = i / nCT;
  t.c_tile = (int)(i - row * nCT);
  t.b = (int)(row / nRT);
  t.r_tile = (int)(row - (long)t.b * nRT);
  return t;
}
__device__ __forceinline__ bool gp_next(GpTile& t, int nRT, int nCT) {   // returns true when the row tile changes
  if (++t.c_tile < nCT) return false;
  t.c_tile = 0;
  if (++t.r_tile == nRT) {
    t.r_tile = 0;
    ++t.b;
  }
  return true;
}

// v = c * ((c * fr) * (c * fc)) for a column pair (mutual matching), then ReLU
__device__ __forceinline__ float2 gp_value(float2 x, float2 fr2, float2 fc, bool mm, bool nrm) {
  if (mm) x = __fmul2_rn(x, __fmul2_rn(__fmul2_rn(x, fr2), __fmul2_rn(x, fc)));
  if (nrm) {
    x.x = fmaxf(x.x, 0.f);
    x.y = fmaxf(x.y, 0.f);
  }
  return x;
}

// One 32-column block of one accumulator row (thread = row).  ncols = valid columns of this block (>= 32: all).
template <int PASS>
// fcol / fnorm / stage are SHARED-space addresses: the accesses below are explicit ld.shared / st.shared (pointers into
// the manually aligned dynamic window lose their provenance and would compile to generic LD.E / ST.E)
__device__ __forceinline__ void gp_block(const uint32_t (&v)[32], uint32_t fcol, uint32_t fnorm, uint32_t stage, int lane, float frow,
                                         bool mm, bool nrm, int ncols, float& racc, float* __restrict__ out_blk,
                                         long ld_out, int nrows) {
  if (PASS == GP_ROWMAX) {
    if (ncols >= 32) {
      float m0 = racc, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        m0 = fmaxf(m0, __uint_as_float(v[i]));
        m1 = fmaxf(m1, __uint_as_float(v[i + 1]));
        m2 = fmaxf(m2, __uint_as_float(v[i + 2]));
        m3 = fmaxf(m3, __uint_as_float(v[i + 3]));
      }
      racc = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (i < ncols) racc = fmaxf(racc, __uint_as_float(v[i]));
    }
    return;
  }
  const float2 fr2 = make_float2(frow, frow);
  if (PASS == GP_ROWSSQ) {
    // zero-filled (out-of-range) columns give v = 0: no column check needed
    float2 acc = make_float2(0.f, 0.f);
#pragma unroll
    for (int i4 = 0; i4 < 8; ++i4) {
      const float4 fc = lds_f4(fcol + 16 * i4);   // warp-uniform address: broadcast
      const float2 a = gp_value(make_float2(__uint_as_float(v[4 * i4]), __uint_as_float(v[4 * i4 + 1])), fr2,
                                make_float2(fc.x, fc.y), mm, nrm);
      const float2 b = gp_value(make_float2(__uint_as_float(v[4 * i4 + 2]), __uint_as_float(v[4 * i4 + 3])), fr2,
                                make_float2(fc.z, fc.w), mm, nrm);
      acc = __ffma2_rn(a, a, acc);
      acc = __ffma2_rn(b, b, acc);
    }
    racc += acc.x + acc.y;
    return;
  }
  // GP_WRITE: transpose through the warp-private buffer (16-byte chunk j of row r lives at chunk j ^ (r & 7)), then
  // each group of 8 lanes stores one full 128-byte line: 4 rows per instruction
  __syncwarp();   // the previous block's reads of the buffer are complete
#pragma unroll
  for (int i4 = 0; i4 < 8; ++i4) {
    const float4 fc = lds_f4(fcol + 16 * i4);
    const float4 fn = lds_f4(fnorm + 16 * i4);
    float2 a = gp_value(make_float2(__uint_as_float(v[4 * i4]), __uint_as_float(v[4 * i4 + 1])), fr2,
                        make_float2(fc.x, fc.y), mm, nrm);
    float2 b = gp_value(make_float2(__uint_as_float(v[4 * i4 + 2]), __uint_as_float(v[4 * i4 + 3])), fr2,
                        make_float2(fc.z, fc.w), mm, nrm);
    a = __fmul2_rn(a, make_float2(fn.x, fn.y));
    b = __fmul2_rn(b, make_float2(fn.z, fn.w));
    sts_u4(stage + lane * 128 + ((i4 ^ (lane & 7)) << 4), __float_as_uint(a.x), __float_as_uint(a.y), __float_as_uint(b.x),
           __float_as_uint(b.y));
  }
  __syncwarp();
  const int j = lane & 7, r0 = lane >> 3;
  float* p = out_blk + (long)r0 * ld_out + 4 * j;   // row r0 of this block, 16-byte piece j
  if (ncols >= 32 && nrows >= 32) {
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int r = it * 4 + r0;
      st_cs_f4(p, lds_f4(stage + r * 128 + ((j ^ (r & 7)) << 4)));
      p += 4 * ld_out;
    }
  } else {
    const bool col_ok = 4 * j < ncols;   // the column count is a multiple of 4: a 16-byte piece is all in or all out
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int r = it * 4 + r0;
      if (col_ok && r < nrows) st_cs_f4(p, lds_f4(stage + r * 128 + ((j ^ (r & 7)) << 4)));
      p += 4 * ld_out;
    }
  }
}

// A = resident operand (rows, NA positions), B = streaming operand (columns, NB positions).
//   rowfac_in [B,NA] / colfac_in [B,NB]: the maxima whose reciprocals (+1e-5) are the mutual-matching factors
//   colnorm_in [B,NB]: sums of squares (GP_WRITE);  row_out [B,NA]: ROWMAX / ROWSSQ result;  out [B,NA,NB]: GP_WRITE
template <int PASS>
__global__ void __launch_bounds__(GP_THREADS, 1)
global_corr_persist_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                           float* __restrict__ out, const float* __restrict__ rowfac_in,
                           const float* __restrict__ colfac_in, const float* __restrict__ colnorm_in,
                           float* __restrict__ row_out, int C, long NA, long NB, int nRT, int nCT, long total,
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
  const int ntiles = (int)(end - start);

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
  const uint32_t tmem = uniform_u32(sh->tmem_base);   // provably warp-uniform: the MMA operands stay in uniform registers

  if (warp == 9) {
    // ------------------------------------------------------------------ TMA producer (whole warp: the boxes of
    // one stage are issued by different lanes of one instruction)
    if (elect_one()) {
      tma_prefetch_desc(&tm_a);
      tma_prefetch_desc(&tm_b);
    }
    GpTile t = gp_decode(start, nRT, nCT);
    bool new_row = true;
    uint32_t a_loads = 0, st = 0, fills = 0;
    for (int j = 0; j < ntiles; ++j) {
      if (new_row) {
        if (a_loads > 0) mbar_wait(&sh->a_empty, (a_loads - 1) & 1);   // every MMA reading the old tile is done
        if (elect_one()) {
          mbar_expect_tx(&sh->a_full, (uint32_t)nkb * GP_KB_BYTES);
          for (int kb = 0; kb < nkb; ++kb)
#pragma unroll
            for (int blk = 0; blk < 4; ++blk)
              tma_load_3d(sA + kb * GP_KB_BYTES + blk * GP_BOX_BYTES, &tm_a, &sh->a_full, t.r_tile * 128 + blk * 32,
                          kb * GP_BK, t.b);
        }
        ++a_loads;
      }
      for (int kb = 0; kb < nkb; ++kb) {
        if (fills >= GP_STAGES) mbar_wait(&sh->b_empty[st], ((fills / GP_STAGES) - 1) & 1);
        if (elect_one()) {
          uint8_t* dst = ring + st * GP_BKB_BYTES;
          mbar_expect_tx(&sh->b_full[st], GP_BKB_BYTES);
#pragma unroll
          for (int blk = 0; blk < GP_TN / 32; ++blk)
            tma_load_3d(dst + blk * GP_BOX_BYTES, &tm_b, &sh->b_full[st], t.c_tile * GP_TN + blk * 32, kb * GP_BK, t.b);
        }
        ++fills;
        st = (st + 1 == GP_STAGES) ? 0 : st + 1;
      }
      new_row = gp_next(t, nRT, nCT);
    }
  } else if (warp == 8) {
    // ------------------------------------------------------------------ MMA issuer (whole warp, elected issue)
    constexpr uint32_t IDESC = make_idesc(FMT_TF32, 128, GP_TN, 1, 1);
    const uint64_t da0 = make_sdesc_sw128_base32(smem_u32(sA), GP_BOX_BYTES, 512);
    const uint64_t db0 = make_sdesc_sw128_base32(smem_u32(ring), GP_BOX_BYTES, 512);
    GpTile t = gp_decode(start, nRT, nCT);
    bool new_row = true;
    uint32_t a_cnt = 0, st = 0, ph = 0;
    for (int j = 0; j < ntiles; ++j) {
      if (new_row) {
        mbar_wait(&sh->a_full, a_cnt & 1);
        ++a_cnt;
      }
      const uint32_t buf = (uint32_t)j & (GP_ACC - 1);
      if (j >= GP_ACC) mbar_wait(&sh->acc_empty[buf], ((j / GP_ACC) - 1) & 1);   // epilogue drained this accumulator
      tc_fence_after();
      const uint32_t d_tmem = tmem + buf * GP_TN;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(&sh->b_full[st], ph);
        tc_fence_after();
        const uint64_t da = da0 + (uint64_t)(kb * (GP_KB_BYTES >> 4));
        const uint64_t db = db0 + (uint64_t)(st * (GP_BKB_BYTES >> 4));
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < GP_BK / 8; ++k)
            mma_tf32_ss(d_tmem, da + (uint64_t)(k * 64), db + (uint64_t)(k * 64), IDESC, (kb > 0 || k > 0) ? 1u : 0u);
          tc_commit(&sh->b_empty[st]);
        }
        __syncwarp();
        if (++st == GP_STAGES) {
          st = 0;
          ph ^= 1;
        }
      }
      new_row = gp_next(t, nRT, nCT);
      if (elect_one()) {
        tc_commit(&sh->acc_full[buf]);
        if (new_row && j + 1 < ntiles) tc_commit(&sh->a_empty);   // the resident tile may be replaced
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps: thread = accumulator row
    const int q = warp & 3, hf = warp >> 2;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    uint8_t* stage = stage_all + warp * GP_STAGE_BYTES;
    float* fcol = sh->fac[warp][0];
    float* fnorm = sh->fac[warp][1];
    GpTile t = gp_decode(start, nRT, nCT);
    bool new_row = true;
    long row = 0;          // this thread's row (position in A) and its validity
    bool row_ok = false;
    float racc = 0.f, frow = 1.f;
    for (int j = 0; j < ntiles; ++j) {
      if (new_row) {
        row = (long)t.r_tile * 128 + q * 32 + lane;
        row_ok = row < NA;
        racc = PASS == GP_ROWMAX ? -INFINITY : 0.f;
        if (PASS != GP_ROWMAX) frow = (mm && row_ok) ? 1.f / (rowfac_in[(long)t.b * NA + row] + 1e-5f) : 1.f;
      }
      const long colh = (long)t.c_tile * GP_TN + hf * (GP_TN / 2);   // first column of this warp's half
      const int ncols = (int)(NB - colh < GP_TN / 2 ? NB - colh : GP_TN / 2);   // valid columns of the half (may be <= 0)
      if (PASS != GP_ROWMAX) {
        __syncwarp();   // the previous tile's reads of fac[] are complete
#pragma unroll
        for (int c = 0; c < GP_TN / 64; ++c) {
          const int cc = c * 32 + lane;
          fcol[cc] = (mm && cc < ncols) ? 1.f / (colfac_in[(long)t.b * NB + colh + cc] + 1e-5f) : 1.f;
          if (PASS == GP_WRITE)
            fnorm[cc] = (nrm && cc < ncols) ? 1.f / fmaxf(sqrtf(colnorm_in[(long)t.b * NB + colh + cc]), 1e-12f) : 1.f;
        }
        __syncwarp();
      }
      const uint32_t buf = (uint32_t)j & (GP_ACC - 1);
      mbar_wait(&sh->acc_full[buf], (j / GP_ACC) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem + lane_off + buf * GP_TN + hf * (GP_TN / 2);
      const long r_base = (long)t.r_tile * 128 + q * 32;
      const int nrows = (int)(NA - r_base < 32 ? NA - r_base : 32);
      float* out_blk = PASS == GP_WRITE ? out + ((long)t.b * NA + r_base) * NB + colh : nullptr;
#pragma unroll
      for (int ch = 0; ch < GP_TN / 128; ++ch) {   // 64 columns at a time
        uint32_t v0[32], v1[32];
        tmem_ld32(taddr + ch * 64, v0);
        tmem_ld32(taddr + ch * 64 + 32, v1);
        tc_wait_ld();
        if (ch == GP_TN / 128 - 1) {   // the whole accumulator has been read: release it to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&sh->acc_empty[buf]);
        }
        gp_block<PASS>(v0, smem_u32(fcol + ch * 64), smem_u32(fnorm + ch * 64), smem_u32(stage), lane, frow, mm, nrm, ncols - ch * 64, racc,
                       out_blk + ch * 64, NB, nrows);
        gp_block<PASS>(v1, smem_u32(fcol + ch * 64 + 32), smem_u32(fnorm + ch * 64 + 32), smem_u32(stage), lane, frow, mm, nrm, ncols - ch * 64 - 32,
                       racc, out_blk + ch * 64 + 32, NB, nrows);
      }
      const int pb = t.b;
      new_row = gp_next(t, nRT, nCT);
      if ((new_row || j + 1 == ntiles) && row_ok) {   // end of this CTA's share of the row: publish the register
        if (PASS == GP_ROWMAX) gp_atomic_max_float(row_out + (long)pb * NA + row, racc);
        if (PASS == GP_ROWSSQ) atomicAdd(row_out + (long)pb * NA + row, racc);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc<512>(tmem);
}

bool global_corr_persist_supported(int C, long Ns, long Nt, const void* a, const void* b, const void* c) {
  return C % GP_BK == 0 && C / GP_BK <= GP_MAXKB && Ns % 4 == 0 && Nt % 4 == 0 && Ns < (1l << 31) - 128 &&
         Nt < (1l << 31) - 128 && (((uintptr_t)a | (uintptr_t)b | (uintptr_t)c) & 15) == 0;
}

template <int PASS>
static int gp_launch(const CUtensorMap& ta, const CUtensorMap& tb, float* out, const float* rowfac, const float* colfac,
                     const float* colnorm, float* row_out, int B, int C, long NA, long NB, int mode, cudaStream_t st) {
  static bool attr_set = false;   // per template instantiation
  if (!attr_set) {
    RF_CUDA(cudaFuncSetAttribute(global_corr_persist_kernel<PASS>, cudaFuncAttributeMaxDynamicSharedMemorySize, GP_SMEM));
    attr_set = true;
  }
  const int nRT = (int)ceil_div(NA, 128), nCT = (int)ceil_div(NB, GP_TN);
  const long total = (long)B * nRT * nCT;
  const unsigned grid = (unsigned)(total < kNumSMs ? total : kNumSMs);
  global_corr_persist_kernel<PASS><<<grid, GP_THREADS, GP_SMEM, st>>>(ta, tb, out, rowfac, colfac, colnorm, row_out, C, NA,
                                                                      NB, nRT, nCT, total, mode);
  RF_CHECK_LAUNCH("global_corr_persist_kernel");
  return RF_OK;
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
  if (mode & 1) {   // A[s] = max_t corr (rows = src) and Bm[t] = max_s corr (rows = trg: the transposed product)
    rc = gp_launch<GP_ROWMAX>(ts, tt, nullptr, nullptr, nullptr, nullptr, rowmax, B, C, Ns, Nt, mode, st);
    if (rc != RF_OK) return rc;
    rc = gp_launch<GP_ROWMAX>(tt, ts, nullptr, nullptr, nullptr, nullptr, colmax, B, C, Nt, Ns, mode, st);
    if (rc != RF_OK) return rc;
  }
  if (mode & 2) {   // ||v[:, t]||^2 as row sums of the transposed product (rows = trg, columns = src)
    RF_CUDA(cudaMemsetAsync(normsq, 0, sizeof(float) * (size_t)B * Nt, st));
    rc = gp_launch<GP_ROWSSQ>(tt, ts, nullptr, colmax, rowmax, nullptr, normsq, B, C, Nt, Ns, mode, st);
    if (rc != RF_OK) return rc;
  }
  return gp_launch<GP_WRITE>(ts, tt, out, rowmax, colmax, normsq, nullptr, B, C, Ns, Nt, mode, st);
}

}  // namespace rf
