// bf16 GEMM on tcgen05 with fused epilogues for the Linear layers of the MiT encoder and the DAFormer head
// (reference: models/backbones/mix_transformer.py:96-103,137-164 q / kv / proj / fc1 / fc2 nn.Linear;
// models/modules.py:59-68 MLP embedding) -- forward, input gradient and weight gradient through ONE kernel:
//     out[m, n] (+)= sum_k A(m, k) * B(n, k)  (+ bias[n])
//   forward  y  = x W^T + b : A = x  [T, K]  (K-major),  B = W [N, K] (K-major)            -> bf16 out, fp32 bias fused
//   dgrad    dx = dy W      : A = dy [T, N]  (K-major),  B = W [N, K] read MN-major        -> bf16 out
//   wgrad    dW += dy^T x   : A = dy [T, N] read MN-major, B = x [T, K] read MN-major      -> fp32 out, accumulated in place
//                             (split over the token dimension, partial sums leave with red.global.add.f32 straight
//                             into the flat gradient buffer: no scratch, no memset, no separate add)
// Both operand layouts are consumed in place through TMA + UMMA shared-memory descriptors (128-byte swizzle; an
// MN-major operand is a set of [64 k x 64 m] boxes, 8 KiB each): no transposed copies of weights or activations.
//
// Persistent, warp-specialised: warp 8 = TMA producer, warp 9 = MMA issuer, warps 0-7 = two epilogue warpgroups (thread =
// accumulator row, alternate 128-byte column chunks).  128 x BN output tiles (BN = 128, or 64 for narrow outputs), BK = 64, a ring of smem stages, and the 512 TMEM
// columns as 4 (BN = 128) / 8 (BN = 64) accumulator buffers so the MMA warp runs ahead of the epilogue.  Tiles are
// walked n-fastest so consecutive CTAs share the A rows in L2.  All barrier arrivals are one per warp / tcgen05.commit.
// The epilogue leaves through swizzled shared-memory chunks and TMA stores (cp.reduce.async.bulk .add for the weight
// gradient): full 128-byte lines whatever the output pitch (per-thread row stores ran at 2 TB/s on 128 k x 256 outputs).
#include <cuda_bf16.h>
#include <stdlib.h>

#include "rf_common.cuh"
#include "rf_sm100.cuh"
#include "rf_trace.cuh"

namespace rf {
using namespace sm100;

constexpr int GM_BM = 128, GM_BK = 64;
constexpr int GM_THREADS = 352;   // 2 epilogue warpgroups (warps 0-7) + TMA-load warp (8) + MMA warp (9) + TMA-store warp (10)
constexpr int GM_A_BYTES = GM_BM * GM_BK * 2;          // 16 KiB per stage

template <int BN, int MODE, bool CL2 = false>
struct GmCfg {
  // CTA pair (CL2, one cta_group::2 MMA over two SMs): every CTA stages its own 128 rows of A and HALF of the B tile
  static constexpr int B_BYTES = BN * GM_BK * 2 / (CL2 ? 2 : 1);
  static constexpr int STAGE = GM_A_BYTES + B_BYTES;
  // Staging buffers per epilogue warpgroup ([128 rows x 128 B] TMA-store sources).  A TMA store takes ~1 500 cycles from
  // issue until its shared-memory source may be rewritten (timeline traces: 2 000 cycles per chunk with one buffer,
  // the epilogue of a K = 320 tile longer than its MMAs), so two buffers alternate and a chunk only waits for the store
  // before the previous one.  The long-K 256-wide convolution tiles keep the fourth operand stage instead (their
  // epilogue hides behind 144 k-blocks of MMAs).
  static constexpr int EPI = (MODE == 1 && BN == 256 && !CL2) ? 1 : 2;
  static constexpr int ACC = 512 / BN;                 // accumulator buffers in TMEM (2 for BN = 192 / 256)
  static constexpr int STAGING = 2 * EPI * 16384;
  static constexpr int BIAS_BYTES = EPI * BN * 4;       // one bias slice per tile in flight
  static constexpr int FIXED = STAGING + 1024 + 384 + BIAS_BYTES;   // + alignment slack, barrier block
#ifdef WS_TRACE
  static constexpr int BUDGET = 232448 - 22528;        // (the trace log takes 22 KiB of static smem)
#else
  static constexpr int BUDGET = 232448;
#endif
  static constexpr int FIT = (BUDGET - FIXED) / STAGE;
  static constexpr int STAGES = FIT > 8 ? 8 : FIT;
  static constexpr int SMEM = STAGES * STAGE + FIXED;
  static_assert(STAGES >= 2, "operand ring too small");
};

struct GmBars {
  uint64_t full[8], empty[8];
  uint64_t acc_full[8], acc_empty[8];
  uint64_t st_full[4], st_free[4];     // epilogue staging buffers [warpgroup * 2 + buffer]: filled by 4 warps / read out by the TMA store
  uint32_t tmem_base;
};

struct GmParams {
  const float* bias;       // [N] or nullptr
  void* out;               // bf16 or fp32 [M, N] row-major
  int M, N, K;
  int m_tiles, n_tiles, k_blocks, splits, kb_per_split;
  int out_f32, accumulate;
  float* colsum;           // MODE 0, optional: colsum[m] += sum_k A[m, k] (the bias gradient riding on the weight-gradient GEMM)
  // implicit-GEMM 3x3 convolution (MODE 1: forward / input gradient, MODE 2: weight gradient)
  int taps;                // 1 (GEMM, MODE 1) or 9 (MODE 2: one output tile set per filter tap)
  int H, W, tiles_h, tiles_w;      // MODE 1: pixel tiles of 8 x 16;  MODE 2: pixel k-blocks of 4 x 16
  int cin_blocks;          // MODE 1: 64-channel blocks per tap
  int act;                 // MODE 1 epilogue: 0 none, 1 ReLU, 2 LeakyReLU(slope)
  float slope;
  int dil;                 // MODE 1: dilation (= padding) of the 3x3 filter
};

enum { GM_GEMM = 0, GM_CONV = 1, GM_CONV_WGRAD = 2 };

// A / B tiles of one k-block: K-major = one box [rows x 64 k]; MN-major = rows/64 boxes [64 k x 64 rows]
template <bool MN>
__device__ __forceinline__ void gm_load(void* smem_dst, const CUtensorMap* tm, uint64_t* bar, int row0, int k0, int rows) {
  if (!MN) {
    tma_load_3d(smem_dst, tm, bar, k0, row0, 0);
  } else {
    for (int blk = 0; blk < rows / 64; ++blk)
      tma_load_3d(reinterpret_cast<uint8_t*>(smem_dst) + blk * 8192, tm, bar, row0 + blk * 64, k0, 0);
  }
}

template <int BN, bool A_MN, bool B_MN, int MODE, bool CL2, bool CS>
__global__ void __launch_bounds__(GM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                 const __grid_constant__ CUtensorMap tm_out, const GmParams p) {
  using Cfg = GmCfg<BN, MODE, CL2>;
  WS_T_INIT();
  WS_T(0);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* staging = smem + Cfg::STAGES * Cfg::STAGE;
  GmBars* bars = reinterpret_cast<GmBars*>(staging + Cfg::STAGING);
  const int tid = threadIdx.x, warp = tid >> 5;

  if (warp == 8) {
    if (elect_one()) {
      tma_prefetch_desc(&tm_a);
      tma_prefetch_desc(&tm_b);
      tma_prefetch_desc(&tm_out);
      for (int s = 0; s < Cfg::STAGES; ++s) {
        mbar_init(&bars->full[s], 1);              // CTA pair: the leader's barrier collects both CTAs' loads
        mbar_init(&bars->empty[s], 1);
      }
      for (int s = 0; s < Cfg::ACC; ++s) {      // (all of them; with column sums only the first `nacc` are used)
        mbar_init(&bars->acc_full[s], 1);
        // CTA pair: the leader's MMAs also wait for ONE forwarded arrival per tile from the peer CTA (see the MMA warp)
        mbar_init(&bars->acc_empty[s], (CL2 && cluster_ctarank() == 0) ? 9 : 8);
      }
      for (int s = 0; s < 4; ++s) {
        mbar_init(&bars->st_full[s], 4);
        mbar_init(&bars->st_free[s], 1);
      }
      fence_barrier_init();
    }
    __syncwarp();
  }
  if (warp == 9) {
    if (CL2) tmem_alloc_2cta<512>(&bars->tmem_base); else tmem_alloc<512>(&bars->tmem_base);
  }
  // Column sums of A (the bias gradient db[n] = sum_t dy[t, n] of a Linear layer, riding on its weight-gradient GEMM
  // dW = dy^T x whose A operand IS dy^T): one extra N = 16 MMA per k-step against a tile of ONES puts sum_k A[m, k] into
  // 16 spare accumulator columns -- the tensor core does the reduction, a separate column-sum launch per layer goes away.
  // The ones tile (no-swizzle K-major core matrices, 512 B) lives in the bias slice, which an accumulating GEMM never uses;
  // the spare columns are 448 + 16 * buffer, so at most 448 / BN (<= 4) accumulator buffers are in flight.
  // (CS is a template parameter: with the buffer count a run-time value every `local % nacc` of the three roles became an
  //  integer division and the plain GEMMs lost ~10 %)
  constexpr bool do_cs = CS;
  constexpr int nacc = CS ? ((448 / BN) < 4 ? (448 / BN) : 4) : Cfg::ACC;
  static_assert(!CS || (MODE == GM_GEMM && !CL2 && nacc >= 1), "column sums ride on the plain single-CTA GEMM");
  if (do_cs && tid < 256) {
    reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(bars) + 384)[tid] = 0x3F803F80u;   // bf16 (1.0, 1.0) x 512
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  if (CL2) cluster_sync_all();   // the peer's barriers are initialised before anything arrives on them
  tc_fence_after();
  const uint32_t tmem = uniform_u32(bars->tmem_base);
  WS_T(1);

  // CTA pairs (CL2): the two CTAs of a cluster work on the m-tiles 2 mp and 2 mp + 1 of the SAME n-tile as ONE
  // tcgen05.mma.cta_group::2 of M = 256: each CTA stages its own A tile and half of the B tile (the tensor core reads
  // both halves), so the operand traffic per SM and k-block drops from (128 + BN) to (128 + BN / 2) rows -- these GEMMs
  // are bound by the L2 -> SM operand path (~50 B/clk per SM measured).  The leader (rank 0) issues the MMAs and its
  // commits arrive on both CTAs' barriers; each CTA's epilogue drains its own 128 accumulator rows.  An m-tile beyond
  // the tensor loads zeros and stores nothing (TMA fills / clips), so the pair never diverges.
  const int rank = CL2 ? (int)cluster_ctarank() : 0;
  const int m_units = CL2 ? (p.m_tiles + 1) / 2 : p.m_tiles;
  const int tiles = p.taps * m_units * p.n_tiles;
  const long items = (long)tiles * p.splits;          // work items: (tile, k split), tile-major
  const long w_first = CL2 ? blockIdx.x / 2 : blockIdx.x, w_step = CL2 ? gridDim.x / 2 : gridDim.x;

  if (warp == 8) {
    // ------------------------------------------------------------------ TMA producer
    int it = 0;
    for (long w = w_first; w < items; w += w_step) {
      const int tile = (int)(w / p.splits), split = (int)(w % p.splits);
      const int tap = tile / (m_units * p.n_tiles);                   // MODE 2 only (else 0)
      const int mt = ((tile / p.n_tiles) % m_units) * (CL2 ? 2 : 1) + rank;
      const int m0 = mt * GM_BM, n0 = (tile % p.n_tiles) * BN;
      const int kb0 = split * p.kb_per_split, kb1 = min(p.k_blocks, kb0 + p.kb_per_split);
      // MODE 1: the 128 output pixels of this tile are the 8 x 16 patch (th, tw) of image b
      const int cb_img = mt / (p.tiles_h * p.tiles_w), cth = (mt / p.tiles_w) % p.tiles_h, ctw = mt % p.tiles_w;
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        const int st = it % Cfg::STAGES;
        if (it >= Cfg::STAGES) mbar_wait(&bars->empty[st], ((it / Cfg::STAGES) - 1) & 1);
        uint8_t* stage = smem + st * Cfg::STAGE;
        if (elect_one()) {
          uint8_t* sB = stage + GM_A_BYTES;
          if (!CL2) {
            mbar_expect_tx(&bars->full[st], Cfg::STAGE);
            if (MODE == GM_GEMM) {
              gm_load<A_MN>(stage, &tm_a, &bars->full[st], m0, kb * GM_BK, GM_BM);
              gm_load<B_MN>(sB, &tm_b, &bars->full[st], n0, kb * GM_BK, BN);
            } else if (MODE == GM_CONV) {
              // k-block = (filter tap, 64 input channels): the activation box is shifted by the tap, rows / columns
              // outside the image are zero-filled by TMA (= the convolution's zero padding)
              const int ctap = kb / p.cin_blocks, cb = kb % p.cin_blocks;
              tma_load_4d(stage, &tm_a, &bars->full[st], cb * 64, ctw * 16 + (ctap % 3 - 1) * p.dil, cth * 8 + (ctap / 3 - 1) * p.dil, cb_img);
              tma_load_3d(sB, &tm_b, &bars->full[st], cb * 64, ctap, n0);
            } else {
              // weight gradient: k-block = 4 x 16 pixels of one image; A = dy (output channels m0.. as MN-major blocks),
              // B = x shifted by the tap (input channels n0..)
              const int kw = kb % p.tiles_w, kh = (kb / p.tiles_w) % p.tiles_h, kimg = kb / (p.tiles_w * p.tiles_h);
              for (int blk = 0; blk < GM_BM / 64; ++blk)
                tma_load_4d(stage + blk * 8192, &tm_a, &bars->full[st], m0 + blk * 64, kw * 16, kh * 4, kimg);
              for (int blk = 0; blk < BN / 64; ++blk)
                tma_load_4d(sB + blk * 8192, &tm_b, &bars->full[st], n0 + blk * 64, kw * 16 + tap % 3 - 1, kh * 4 + tap / 3 - 1, kimg);
            }
          } else {
            // both CTAs' loads complete on the LEADER's barrier, which the leader arms with the bytes of both (a remote
            // arrive.expect_tx from the peer cost ~1 000 cycles per k-block in the timeline trace); the peer cannot run a
            // phase ahead: its next load into this stage waits for the commit of the MMAs that consumed this one
            const uint32_t fb = mapa_u32(smem_u32(&bars->full[st]), 0);
            if (rank == 0) mbar_expect_tx(&bars->full[st], 2 * Cfg::STAGE);
            const int nh = n0 + rank * (BN / 2);            // this CTA's half of the B tile
            if (MODE == GM_GEMM) {
              if (!A_MN) {
                tma_load_3d_2cta(stage, &tm_a, fb, kb * GM_BK, m0, 0);
              } else {
                for (int blk = 0; blk < GM_BM / 64; ++blk) tma_load_3d_2cta(stage + blk * 8192, &tm_a, fb, m0 + blk * 64, kb * GM_BK, 0);
              }
              if (!B_MN) {
                tma_load_3d_2cta(sB, &tm_b, fb, kb * GM_BK, nh, 0);
              } else {
                for (int blk = 0; blk < BN / 128; ++blk) tma_load_3d_2cta(sB + blk * 8192, &tm_b, fb, nh + blk * 64, kb * GM_BK, 0);
              }
            } else if (MODE == GM_CONV) {
              const int ctap = kb / p.cin_blocks, cb = kb % p.cin_blocks;
              tma_load_4d_2cta(stage, &tm_a, fb, cb * 64, ctw * 16 + (ctap % 3 - 1) * p.dil, cth * 8 + (ctap / 3 - 1) * p.dil, cb_img);
              tma_load_3d_2cta(sB, &tm_b, fb, cb * 64, ctap, nh);
            } else {
              const int kw = kb % p.tiles_w, kh = (kb / p.tiles_w) % p.tiles_h, kimg = kb / (p.tiles_w * p.tiles_h);
              for (int blk = 0; blk < GM_BM / 64; ++blk)
                tma_load_4d_2cta(stage + blk * 8192, &tm_a, fb, m0 + blk * 64, kw * 16, kh * 4, kimg);
              for (int blk = 0; blk < BN / 128; ++blk)
                tma_load_4d_2cta(sB + blk * 8192, &tm_b, fb, nh + blk * 64, kw * 16 + tap % 3 - 1, kh * 4 + tap / 3 - 1, kimg);
            }
          }
        }
        __syncwarp();
        WS_T(30);
      }
    }
  } else if (warp == 9) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t IDESC = make_idesc(FMT_BF16, CL2 ? 2 * GM_BM : GM_BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
    constexpr uint64_t STEP = (uint64_t)(Cfg::STAGE >> 4);
    const uint32_t base = smem_u32(smem);
    const uint64_t dA0 = A_MN ? make_sdesc_sw128(base, 8192, 1024) : make_sdesc_sw128(base, 16, 1024);
    const uint64_t dB0 = B_MN ? make_sdesc_sw128(base + GM_A_BYTES, 8192, 1024) : make_sdesc_sw128(base + GM_A_BYTES, 16, 1024);
    constexpr uint64_t KA = A_MN ? 128 : 2, KB = B_MN ? 128 : 2;   // descriptor advance per 16-element k step
    int it = 0, local = 0;
    if (CL2 && rank != 0) {
      // Peer CTA of a pair: no MMAs to issue.  This warp forwards "my epilogue has drained accumulator buffer b" to the
      // leader, so the (slow: ~2 000 cycles with release semantics at cluster scope) remote arrival is not paid by the
      // eight epilogue warps inside their tile loop.
      for (long w = w_first; w < items; w += w_step, ++local) {
        const int buf = local % nacc;
        mbar_wait(&bars->acc_empty[buf], (local / nacc) & 1);
        if (elect_one()) mbar_arrive_cluster(mapa_u32(smem_u32(&bars->acc_empty[buf]), 0));
        __syncwarp();
      }
    }
    for (long w = (CL2 && rank != 0) ? items : w_first; w < items; w += w_step, ++local) {   // pair: the leader issues
      const int split = (int)(w % p.splits);
      const int kb0 = split * p.kb_per_split, kb1 = min(p.k_blocks, kb0 + p.kb_per_split);
      const int buf = local % nacc;
      if (local >= nacc) mbar_wait(&bars->acc_empty[buf], ((local / nacc) - 1) & 1);
      tc_fence_after();
      const uint32_t d = tmem + buf * BN;
      const bool cs_item = do_cs && ((int)(w / p.splits) % p.n_tiles) == 0;     // one n-tile per m-tile carries the column sum
      constexpr uint32_t IDESC_CS = make_idesc(FMT_BF16, GM_BM, 16, A_MN ? 1 : 0, 0);
      const uint64_t d_ones = make_sdesc(smem_u32(reinterpret_cast<uint8_t*>(bars) + 384), 128, 256, 0);
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        const int st = it % Cfg::STAGES;
        mbar_wait(&bars->full[st], (it / Cfg::STAGES) & 1);
        tc_fence_after();
        WS_T(20);
        const uint64_t dA = dA0 + st * STEP, dB = dB0 + st * STEP;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < GM_BK / 16; ++k) {
            if (CL2) mma_f16_ss_2cta(d, dA + k * KA, dB + k * KB, IDESC, (kb > kb0 || k > 0) ? 1u : 0u);
            else mma_f16_ss(d, dA + k * KA, dB + k * KB, IDESC, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          if (cs_item) {
#pragma unroll
            for (int k = 0; k < GM_BK / 16; ++k)
              mma_f16_ss(tmem + 448 + buf * 16, dA + k * KA, d_ones, IDESC_CS, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          if (CL2) tc_commit_2cta(&bars->empty[st], 3); else tc_commit(&bars->empty[st]);
          if (kb == kb1 - 1) {
            if (CL2) tc_commit_2cta(&bars->acc_full[buf], 3); else tc_commit(&bars->acc_full[buf]);
          }
        }
        __syncwarp();
        WS_T(21);
      }
    }
  } else if (warp == 10) {
    // ------------------------------------------------------------------ TMA-store issuer
    // One thread turns the staging chunks the epilogue warps fill into TMA stores (or TMA reduce-adds for the accumulating
    // fp32 outputs).  Timeline traces showed why this is a warp of its own: issuing one bulk store costs the issuing thread
    // 300-500 cycles and its shared-memory source is only reusable ~1 500 cycles later; with the issue inside the epilogue
    // warpgroups (and a bar.sync around it) every 128-byte column chunk cost ~2 300 cycles and the epilogue of a K = 320
    // tile took longer than its MMAs.
    const int cpc = p.out_f32 ? 32 : 64;
    const int nchunk = BN / cpc;
    const int lane = tid & 31;
    int cnt0 = 0, cnt1 = 0;              // chunks stored so far per warpgroup (mirrors the epilogue warps' counters)
    int prev = -1;                       // staging buffer of the most recently issued store
    for (long w = w_first; w < items; w += w_step) {
      const int tile = (int)(w / p.splits);
      const int tap = tile / (m_units * p.n_tiles);
      const int mt = ((tile / p.n_tiles) % m_units) * (CL2 ? 2 : 1) + rank;
      const int m0 = mt * GM_BM, n0 = (tile % p.n_tiles) * BN;
      const int cb_img = mt / (p.tiles_h * p.tiles_w), cth = (mt / p.tiles_w) % p.tiles_h, ctw = mt % p.tiles_w;
#pragma unroll 1
      for (int c = 0; c < nchunk; ++c) {
        const int wg = c & 1;
        const int cnt = wg ? cnt1 : cnt0;
        if (wg) ++cnt1; else ++cnt0;
        const int b = Cfg::EPI == 2 ? (cnt & 1) : 0, use = cnt / Cfg::EPI;
        const int slot = wg * 2 + b;
        mbar_wait(&bars->st_full[slot], use & 1);
        if (lane == 0) {
          const uint8_t* sbuf = staging + (wg * Cfg::EPI + b) * 16384;
          const int col = n0 + c * cpc;
          if (col < p.N) {
            if (MODE == GM_CONV) tma_store_4d(&tm_out, sbuf, col, ctw * 16, cth * 8, cb_img);
            else if (MODE == GM_CONV_WGRAD) tma_reduce_add_3d(&tm_out, sbuf, col, tap, m0);
            else if (p.accumulate) tma_reduce_add_3d(&tm_out, sbuf, col, m0, 0);
            else tma_store_3d(&tm_out, sbuf, col, m0, 0);
          }
          tma_store_commit();            // one group per chunk, empty or not: wait_group.read<1> counts groups
          WS_T(11);
          if (prev >= 0) {               // every store but the one just issued has read its source: hand that buffer back
            tma_store_wait_read<1>();
            mbar_arrive(&bars->st_free[prev]);
          }
        }
        prev = slot;
        __syncwarp();
      }
    }
    if (lane == 0) tma_store_wait_read<0>();   // the staging smem may go away; the global writes complete with the grid
    WS_T(12);
  } else {
    // ------------------------------------------------------------------ epilogue warpgroups (thread = accumulator row)
    // TMEM -> registers (+ bias, -> bf16) -> swizzled smem chunk [128 rows x 128 B], handed to the store warp: full
    // 128-byte lines leave the SM whatever the row pitch, the tensor map clips ragged M / N edges, and the stores run
    // asynchronously.  Two warpgroups take alternate 128-byte column chunks of every tile; each warp runs on its own
    // (mbarrier hand-offs per staging buffer, no CTA-wide barrier inside a tile).
    const int wg = warp >> 2, r = tid & 127;
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    const int cpc = p.out_f32 ? 32 : 64;                 // output columns per 128-byte chunk
    const int nchunk = BN / cpc;
    const uint32_t sw = (uint32_t)(r & 7);
    uint8_t* const sbuf0 = staging + wg * (Cfg::EPI * 16384);
    int nstore = 0;                      // chunks this warpgroup has handed over so far (selects the staging buffer)
    // Bias: EPI slices of BN floats behind the barrier block.  Each warpgroup stages and reads only the columns of ITS
    // chunks (one element per thread), so the slice hand-over is a 128-thread named barrier per warpgroup and the two
    // warpgroups never wait for each other.
    float* const sbias0 = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 384);
    const uint32_t sbias0_s = smem_u32(sbias0), sbuf0_s = smem_u32(sbuf0);   // explicit shared-space accesses below
    const int my_col = ((r / cpc) * 2 + wg) * cpc + r % cpc;       // the tile column whose bias this thread stages
    auto bias_at = [&](long ww) -> float {   // (0 beyond N / the end of the work list)
      if (ww >= items || my_col >= BN) return 0.f;
      const int nn = (int)((ww / p.splits) % p.n_tiles) * BN + my_col;
      return nn < p.N ? __ldg(p.bias + nn) : 0.f;
    };
    auto wg_sync = [&]() {
      if (wg == 0) asm volatile("bar.sync 1, 128;" ::: "memory"); else asm volatile("bar.sync 2, 128;" ::: "memory");
    };
    // one 32-column half of a bf16 chunk: + bias, activation, pack -> 16 words
    auto convert_half = [&](const uint32_t (&v)[32], uint32_t sb, uint32_t* pkh) {   // sb: shared address of 32 bias floats
#pragma unroll
      for (int e = 0; e < 16; e += 2) {
        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.bias) bv = lds_f4(sb + 8 * e);
        float4 x = make_float4(__uint_as_float(v[2 * e]) + bv.x, __uint_as_float(v[2 * e + 1]) + bv.y,
                               __uint_as_float(v[2 * e + 2]) + bv.z, __uint_as_float(v[2 * e + 3]) + bv.w);
        if (MODE == GM_CONV && p.act) {   // fused ReLU / LeakyReLU of the frozen conv stacks
          const float ng = p.act == 1 ? 0.f : p.slope;
          x.x = x.x > 0.f ? x.x : x.x * ng;
          x.y = x.y > 0.f ? x.y : x.y * ng;
          x.z = x.z > 0.f ? x.z : x.z * ng;
          x.w = x.w > 0.f ? x.w : x.w * ng;
        }
        const __nv_bfloat162 h0 = __floats2bfloat162_rn(x.x, x.y);
        const __nv_bfloat162 h1 = __floats2bfloat162_rn(x.z, x.w);
        pkh[e] = *reinterpret_cast<const uint32_t*>(&h0);
        pkh[e + 1] = *reinterpret_cast<const uint32_t*>(&h1);
      }
    };
    float bias_next = 0.f;
    if (p.bias && Cfg::EPI == 2 && my_col < BN) sts_f1(sbias0_s + 4 * my_col, bias_at(w_first));
    int local = 0;
    for (long w = w_first; w < items; w += w_step, ++local) {
      const int buf = local % nacc;
      const uint32_t sbias = sbias0_s + (Cfg::EPI == 2 ? (local & 1) * BN * 4 : 0);
      if (p.bias) {   // this tile's bias slice, zero beyond N, read back as broadcast shared-memory vectors
        if (Cfg::EPI == 2) {
          // slice `local` was written during the previous tile (or before the loop); the next tile's element is
          // fetched now and stored when this tile is done, so no global-load latency sits in front of a tile
          wg_sync();
          bias_next = bias_at(w + w_step);
        } else {
          wg_sync();                                          // the warpgroup is done with the previous tile's slice
          if (my_col < BN) sts_f1(sbias + 4 * my_col, bias_at(w));
          wg_sync();
        }
      }
      mbar_wait(&bars->acc_full[buf], (local / nacc) & 1);
      tc_fence_after();
      WS_T(10);
      const uint32_t t = tmem + lane_off + buf * BN;
      // hand one finished chunk (32 packed words per thread = one 128-byte row) to the store warp
      auto hand_over = [&](const uint32_t (&pk)[32]) {
        WS_T(13);
        const int b = Cfg::EPI == 2 ? (nstore & 1) : 0, use = nstore / Cfg::EPI;
        ++nstore;
        if (use > 0) mbar_wait(&bars->st_free[wg * 2 + b], (use - 1) & 1);   // the store that last used this buffer has read it
        WS_T(14);
        const uint32_t srow = sbuf0_s + b * 16384 + r * 128;
#pragma unroll
        for (int q = 0; q < 8; ++q) sts_u4(srow + ((q ^ sw) << 4), pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
        fence_proxy_async();
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(&bars->st_full[wg * 2 + b]);
        WS_T(15);
      };
      auto release_acc = [&]() {   // this warpgroup's share of the accumulator is in registers: hand the TMEM buffer back
        tc_fence_before();
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(&bars->acc_empty[buf]);
      };
      if (p.out_f32) {
#pragma unroll 1
        for (int c = wg; c < nchunk; c += 2) {
          uint32_t pk[32];
          tmem_ld32(t + c * 32, pk);
          tc_wait_ld();
          if (p.bias) {
#pragma unroll
            for (int e = 0; e < 32; e += 4) {
              const float4 bv = lds_f4(sbias + 4 * (c * 32 + e));
              pk[e] = __float_as_uint(__uint_as_float(pk[e]) + bv.x);
              pk[e + 1] = __float_as_uint(__uint_as_float(pk[e + 1]) + bv.y);
              pk[e + 2] = __float_as_uint(__uint_as_float(pk[e + 2]) + bv.z);
              pk[e + 3] = __float_as_uint(__uint_as_float(pk[e + 3]) + bv.w);
            }
          }
          if (MODE == GM_CONV && p.act) {
#pragma unroll
            for (int e = 0; e < 32; ++e) {
              const float x = __uint_as_float(pk[e]);
              pk[e] = __float_as_uint(x > 0.f ? x : (p.act == 1 ? 0.f : x * p.slope));
            }
          }
          if (c + 2 >= nchunk) {
            if (do_cs && wg == 0 && ((int)(w / p.splits) % p.n_tiles) == 0) {
              uint32_t cs[8];
              tmem_ld8(tmem + lane_off + 448 + buf * 16, cs);
              tc_wait_ld();
              const int m = (int)((w / p.splits) / p.n_tiles) % p.m_tiles * GM_BM + r;
              if (m < p.M) atomicAdd(p.colsum + m, __uint_as_float(cs[0]));
            }
            release_acc();
          }
          hand_over(pk);
        }
      } else {
        // bf16 output: 64-column chunks read as two 32-column halves, software-pipelined -- the TMEM load of the next
        // half (also the first half of the warpgroup's NEXT chunk) is in flight while this one is converted and stored
        uint32_t va[32], vb[32];
        if (wg < nchunk) tmem_ld32(t + wg * 64, va);
#pragma unroll 1
        for (int c = wg; c < nchunk; c += 2) {
          uint32_t pk[32];
          tc_wait_ld();
          tmem_ld32(t + c * 64 + 32, vb);
          convert_half(va, sbias + 4 * (c * 64), pk);
          tc_wait_ld();
          if (c + 2 < nchunk) tmem_ld32(t + (c + 2) * 64, va); else release_acc();
          convert_half(vb, sbias + 4 * (c * 64 + 32), pk + 16);
          hand_over(pk);
        }
      }
      if (nchunk == 1 && wg == 1) {   // BN = 64 bf16: a single chunk, warpgroup 1 only releases the accumulator
        tc_fence_before();
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(&bars->acc_empty[buf]);
      }
      if (p.bias && Cfg::EPI == 2 && my_col < BN) sts_f1(sbias0_s + 4 * (((local + 1) & 1) * BN + my_col), bias_next);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CL2) cluster_sync_all();   // no CTA leaves while the pair's MMAs may still read its smem / arrive on its barriers
  WS_T(2);
  WS_T_FLUSH();
  if (warp == 9) {
    if (CL2) tmem_dealloc_2cta<512>(tmem); else tmem_dealloc<512>(tmem);
  }
}

// CTA pairs need a B tile that splits in two halves of whole swizzle groups / 64-column blocks
constexpr bool gm_pairable(int BN, bool B_MN) { return B_MN ? (BN % 128 == 0) : (BN % 16 == 0); }

template <int BN, bool A_MN, bool B_MN, int MODE, bool CL2, bool CS = false>
static int gemm_launch_impl(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, GmParams p, cudaStream_t st) {
  using Cfg = GmCfg<BN, MODE, CL2>;
  static bool attr = false;
  auto kern = gemm_bf16_kernel<BN, A_MN, B_MN, MODE, CL2, CS>;
  if (!attr) {
    RF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    attr = true;
  }
  const long m_units = CL2 ? (p.m_tiles + 1) / 2 : p.m_tiles;
  const long items = (long)p.taps * m_units * p.n_tiles * p.splits;
  const int per = CL2 ? 2 : 1;
  const long slots = kNumSMs / per;
  const int grid = (int)((items < slots ? items : slots) * per);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(GM_THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = per;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  RF_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, to, p));
  return RF_OK;
}

// CTA pairs (one tcgen05.mma.cta_group::2 of M = 256 over two SMs, each CTA staging half of the B tile) are opt-in
// (RF_GEMM_PAIRS=1).  Measured on the 16 MiT / DAFormer layer shapes in CUDA-graph replay: 154 / 149 / 154 us (forward /
// dgrad / wgrad totals) against 152 / 148 / 153 us for single CTAs.  The pair's main loop IS faster (613 cycles per
// 64-wide k-block of a 256 x 256 tile against ~870: the operand ingest per SM is halved), but at K <= 512 both variants
// are bound by the epilogue (~4 800 cycles per 128 x 256 tile per CTA in the timeline traces), so the totals do not move.
static bool gm_use_pairs() {
  static const bool on = [] { const char* e = getenv("RF_GEMM_PAIRS"); return e && e[0] == '1'; }();
  return on;
}

template <int BN, bool A_MN, bool B_MN, int MODE = GM_GEMM>
static int gemm_launch(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, GmParams p, cudaStream_t st,
                       bool pairs = false) {
  if constexpr (gm_pairable(BN, B_MN)) {
    if (pairs) return gemm_launch_impl<BN, A_MN, B_MN, MODE, true>(ta, tb, to, p, st);
  }
  if constexpr (MODE == GM_GEMM && A_MN && B_MN && BN <= 192) {     // the weight-gradient layouts can carry the column sum
    if (p.colsum != nullptr) return gemm_launch_impl<BN, A_MN, B_MN, MODE, false, true>(ta, tb, to, p, st);
  }
  if (p.colsum != nullptr) {
    set_error("rf_gemm_bf16: the fused column sum needs a_mn_major = b_mn_major = 1 (the weight-gradient layouts)");
    return RF_EINVAL;
  }
  return gemm_launch_impl<BN, A_MN, B_MN, MODE, false>(ta, tb, to, p, st);
}

// Tile width: these small-K GEMMs are bound by the operand traffic L2 -> SM (~50 B/clk per SM measured) and by wave
// quantisation over the persistent CTAs, so pick the BN in {64, 128, 192, 256} that minimises
//   rounds x (operand rows per CTA and k-block)  =  rounds x (128 + BN)        single CTAs, 148 slots
//                                                   rounds x (128 + BN / 2)    CTA pairs sharing B, 74 slots
static int pick_bn(long m_tiles, int N, bool b_mn, bool* pairs_out) {
  const bool want_pairs = gm_use_pairs() && m_tiles >= 2;
  int BN = 64;
  bool pr = want_pairs && gm_pairable(64, b_mn);
  if (N > 64) {
    long best = -1;
    for (int cand : {256, 192, 128}) {
      const bool cp = want_pairs && gm_pairable(cand, b_mn);
      const long n_tiles = (N + cand - 1) / cand;
      const long rounds = cp ? (((m_tiles + 1) / 2) * n_tiles + kNumSMs / 2 - 1) / (kNumSMs / 2)
                             : (m_tiles * n_tiles + kNumSMs - 1) / kNumSMs;
      const long cost = rounds * (128 + (cp ? cand / 2 : cand));
      if (best < 0 || cost < best) {
        best = cost;
        BN = cand;
        pr = cp;
      }
    }
  }
  *pairs_out = pr;
  return BN;
}

}  // namespace rf

using namespace rf;

extern "C" int rf_gemm_bf16(const void* a, const void* b, const float* bias, void* out, int M, int N, int K, int a_mn_major,
                            int b_mn_major, int out_f32, int accumulate, float* colsum, void* stream) {
  RF_REQUIRE(a && b && out && M > 0 && N > 0 && K > 0, "rf_gemm_bf16: bad arguments");
  RF_REQUIRE(!colsum || (accumulate && out_f32), "rf_gemm_bf16: the fused column sum rides on an accumulating fp32 GEMM");
  RF_REQUIRE((((uintptr_t)a | (uintptr_t)b | (uintptr_t)out) & 15) == 0, "rf_gemm_bf16: operands must be 16-byte aligned");
  RF_REQUIRE(!(accumulate && !out_f32), "rf_gemm_bf16: accumulation needs an fp32 output");
  RF_REQUIRE(!(accumulate && bias), "rf_gemm_bf16: bias and accumulation are exclusive");
  // TMA: every global row pitch must be a multiple of 16 bytes
  const long a_pitch = a_mn_major ? M : K, b_pitch = b_mn_major ? N : K;
  RF_REQUIRE(a_pitch % 8 == 0 && b_pitch % 8 == 0 && N % (out_f32 ? 4 : 8) == 0,
             "rf_gemm_bf16: M / N / K pitches must be multiples of 8 elements (got M %d N %d K %d)", M, N, K);
  bool pairs = false;
  int BN = pick_bn((M + GM_BM - 1) / GM_BM, N, b_mn_major != 0, &pairs);
  if (colsum) {            // 16 spare accumulator columns per buffer: single CTAs, tiles of at most 192 columns
    pairs = false;
    if (BN == 256) BN = 128;
  }
  CUtensorMap ta, tb;
  int rc;
  if (a_mn_major)   // stored [K, M]: box = 64 m (inner) x 64 k
    rc = make_tmap_3d(&ta, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a, (uint64_t)M, (uint64_t)K, 1, (uint64_t)M * 2, (uint64_t)M * K * 2, 64, 64);
  else              // stored [M, K]: box = 64 k (inner) x 128 m
    rc = make_tmap_3d(&ta, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a, (uint64_t)K, (uint64_t)M, 1, (uint64_t)K * 2, (uint64_t)M * K * 2, 64, GM_BM);
  if (rc != RF_OK) return rc;
  if (b_mn_major)
    rc = make_tmap_3d(&tb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, b, (uint64_t)N, (uint64_t)K, 1, (uint64_t)N * 2, (uint64_t)N * K * 2, 64, 64);
  else
    rc = make_tmap_3d(&tb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, b, (uint64_t)K, (uint64_t)N, 1, (uint64_t)K * 2, (uint64_t)N * K * 2, 64,
                      (uint32_t)(pairs ? BN / 2 : BN));   // a CTA pair loads half of the rows each
  if (rc != RF_OK) return rc;
  CUtensorMap to;   // output [M, N]: 128-byte chunks of 128 rows
  if (out_f32)
    rc = make_tmap_3d(&to, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, out, (uint64_t)N, (uint64_t)M, 1, (uint64_t)N * 4, (uint64_t)M * N * 4, 32, GM_BM);
  else
    rc = make_tmap_3d(&to, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, out, (uint64_t)N, (uint64_t)M, 1, (uint64_t)N * 2, (uint64_t)M * N * 2, 64, GM_BM);
  if (rc != RF_OK) return rc;
  GmParams p;
  p.bias = bias;
  p.out = out;
  p.M = M;
  p.N = N;
  p.K = K;
  p.m_tiles = (M + GM_BM - 1) / GM_BM;
  p.n_tiles = (N + BN - 1) / BN;
  p.k_blocks = (K + GM_BK - 1) / GM_BK;
  p.out_f32 = out_f32;
  p.accumulate = accumulate;
  p.colsum = colsum;
  p.taps = 1;
  p.H = p.W = p.tiles_h = p.tiles_w = p.cin_blocks = 1;
  p.act = 0;
  p.slope = 0.f;
  p.dil = 1;
  // split the contraction only for accumulating fp32 outputs (the weight gradient: few output tiles, a contraction over
  // every token): enough (tile, split) items for ~2 per SM, at least 4 k-blocks each
  p.splits = 1;
  if (accumulate) {
    const long tiles = (long)(pairs ? (p.m_tiles + 1) / 2 : p.m_tiles) * p.n_tiles;
    long s = (2l * (pairs ? kNumSMs / 2 : kNumSMs)) / tiles;          // floor: (tile, split) items must not spill into a third round
    if (s > p.k_blocks / 4) s = p.k_blocks / 4;
    if (s < 1) s = 1;
    p.splits = (int)s;
  }
  p.kb_per_split = (p.k_blocks + p.splits - 1) / p.splits;
  p.splits = (p.k_blocks + p.kb_per_split - 1) / p.kb_per_split;
  cudaStream_t st = (cudaStream_t)stream;
#define RF_GM(BN_)                                                                        \
  (a_mn_major ? (b_mn_major ? gemm_launch<BN_, true, true>(ta, tb, to, p, st, pairs) : gemm_launch<BN_, true, false>(ta, tb, to, p, st, pairs)) \
              : (b_mn_major ? gemm_launch<BN_, false, true>(ta, tb, to, p, st, pairs) : gemm_launch<BN_, false, false>(ta, tb, to, p, st, pairs)))
  return BN == 64 ? RF_GM(64) : (BN == 128 ? RF_GM(128) : (BN == 192 ? RF_GM(192) : RF_GM(256)));
#undef RF_GM
}

// ---------------------------------------------------------------------------------------------------- 3x3 convolution
// channels-last activation [B, H, W, C] as a rank-4 map (C, W, H, B) with box (64, 16, bh, 1)
static int act_map(CUtensorMap* m, const void* base, int B, int H, int W, int C, int bh, CUtensorMapDataType dt, int elem,
                   uint32_t box0) {
  const uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)B};
  const uint64_t str[3] = {(uint64_t)C * elem, (uint64_t)W * C * elem, (uint64_t)H * W * C * elem};
  const uint32_t box[4] = {box0, 16, (uint32_t)bh, 1};
  return make_tmap_nd(m, dt, elem, base, 4, dims, str, box);
}
// channels-last filter [Cout][3][3][Cin] as a rank-3 map (Cin, tap, Cout)
static int filt_map(CUtensorMap* m, const void* base, int Cin, int Cout, CUtensorMapDataType dt, int elem, uint32_t box0,
                    uint32_t rows) {
  const uint64_t dims[3] = {(uint64_t)Cin, 9, (uint64_t)Cout};
  const uint64_t str[2] = {(uint64_t)Cin * elem, (uint64_t)9 * Cin * elem};
  const uint32_t box[3] = {box0, 1, rows};
  return make_tmap_nd(m, dt, elem, base, 3, dims, str, box);
}

extern "C" int rf_conv3x3_bf16(const void* x, const void* w, const float* bias, void* out, int B, int H, int W, int Cin,
                               int Cout, int out_f32, int act, float slope, int dilation, void* stream) {
  RF_REQUIRE(x && w && out && B > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, "rf_conv3x3_bf16: bad arguments");
  RF_REQUIRE((((uintptr_t)x | (uintptr_t)w | (uintptr_t)out) & 15) == 0, "rf_conv3x3_bf16: operands must be 16-byte aligned");
  RF_REQUIRE(Cin % 8 == 0 && Cout % 8 == 0, "rf_conv3x3_bf16: channel counts must be multiples of 8 (got %d -> %d)", Cin, Cout);
  RF_REQUIRE(act >= 0 && act <= 2, "rf_conv3x3_bf16: act must be 0 (none), 1 (ReLU) or 2 (LeakyReLU)");
  RF_REQUIRE(dilation >= 1 && dilation <= 64, "rf_conv3x3_bf16: dilation must be in [1, 64]");
  GmParams p;
  p.bias = bias;
  p.out = out;
  p.tiles_h = (H + 7) / 8;
  p.tiles_w = (W + 15) / 16;
  p.m_tiles = B * p.tiles_h * p.tiles_w;
  p.M = p.m_tiles * GM_BM;
  p.N = Cout;
  p.K = 9 * Cin;
  bool pairs = false;
  const int BN = pick_bn(p.m_tiles, Cout, false, &pairs);
  p.n_tiles = (Cout + BN - 1) / BN;
  p.cin_blocks = (Cin + 63) / 64;
  p.k_blocks = 9 * p.cin_blocks;
  p.splits = 1;
  p.kb_per_split = p.k_blocks;
  p.out_f32 = out_f32;
  p.accumulate = 0;
  p.colsum = nullptr;
  p.taps = 1;
  p.H = H;
  p.W = W;
  p.act = act;
  p.slope = slope;
  p.dil = dilation;
  CUtensorMap ta, tb, to;
  int rc = act_map(&ta, x, B, H, W, Cin, 8, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, 64);
  if (rc != RF_OK) return rc;
  rc = filt_map(&tb, w, Cin, Cout, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, 64, (uint32_t)(pairs ? BN / 2 : BN));
  if (rc != RF_OK) return rc;
  rc = out_f32 ? act_map(&to, out, B, H, W, Cout, 8, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, 32)
               : act_map(&to, out, B, H, W, Cout, 8, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, 64);
  if (rc != RF_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  switch (BN) {
    case 64: return gemm_launch<64, false, false, GM_CONV>(ta, tb, to, p, st, pairs);
    case 128: return gemm_launch<128, false, false, GM_CONV>(ta, tb, to, p, st, pairs);
    case 192: return gemm_launch<192, false, false, GM_CONV>(ta, tb, to, p, st, pairs);
    default: return gemm_launch<256, false, false, GM_CONV>(ta, tb, to, p, st, pairs);
  }
}

// dw[co][tap][ci] (fp32, channels-last filter layout) += sum over pixels dy[.., co] * x[.. shifted by tap .., ci]
extern "C" int rf_conv3x3_wgrad_bf16(const void* dy, const void* x, float* dw, int B, int H, int W, int Cin, int Cout,
                                     void* stream) {
  RF_REQUIRE(dy && x && dw && B > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, "rf_conv3x3_wgrad_bf16: bad arguments");
  RF_REQUIRE((((uintptr_t)dy | (uintptr_t)x | (uintptr_t)dw) & 15) == 0, "rf_conv3x3_wgrad_bf16: operands must be 16-byte aligned");
  RF_REQUIRE(Cin % 8 == 0 && Cout % 8 == 0, "rf_conv3x3_wgrad_bf16: channel counts must be multiples of 8");
  GmParams p;
  p.bias = nullptr;
  p.out = dw;
  p.taps = 9;
  p.M = Cout;
  p.N = Cin;
  p.m_tiles = (Cout + GM_BM - 1) / GM_BM;
  bool pairs = false;
  int BN = pick_bn(p.m_tiles, Cin, true, &pairs);     // pairs: the two CTAs take adjacent 128-channel slices of Cout
  p.n_tiles = (Cin + BN - 1) / BN;
  p.tiles_h = (H + 3) / 4;            // pixel k-blocks of 4 x 16
  p.tiles_w = (W + 15) / 16;
  p.k_blocks = B * p.tiles_h * p.tiles_w;
  p.K = p.k_blocks * GM_BK;
  p.cin_blocks = 1;
  p.H = H;
  p.W = W;
  p.out_f32 = 1;
  p.accumulate = 1;
  p.colsum = nullptr;
  p.act = 0;
  p.slope = 0.f;
  p.dil = 1;
  const long tiles = 9l * (pairs ? (p.m_tiles + 1) / 2 : p.m_tiles) * p.n_tiles;
  long s = (2l * (pairs ? kNumSMs / 2 : kNumSMs)) / tiles;
  if (s > p.k_blocks / 4) s = p.k_blocks / 4;
  if (s < 1) s = 1;
  p.kb_per_split = (int)((p.k_blocks + s - 1) / s);
  p.splits = (p.k_blocks + p.kb_per_split - 1) / p.kb_per_split;
  CUtensorMap ta, tb, to;
  int rc = act_map(&ta, dy, B, H, W, Cout, 4, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, 64);
  if (rc != RF_OK) return rc;
  rc = act_map(&tb, x, B, H, W, Cin, 4, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, 64);
  if (rc != RF_OK) return rc;
  rc = filt_map(&to, dw, Cin, Cout, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, 32, GM_BM);
  if (rc != RF_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  switch (BN) {
    case 64: return gemm_launch<64, true, true, GM_CONV_WGRAD>(ta, tb, to, p, st, pairs);
    case 128: return gemm_launch<128, true, true, GM_CONV_WGRAD>(ta, tb, to, p, st, pairs);
    case 192: return gemm_launch<192, true, true, GM_CONV_WGRAD>(ta, tb, to, p, st, pairs);
    default: return gemm_launch<256, true, true, GM_CONV_WGRAD>(ta, tb, to, p, st, pairs);
  }
}
