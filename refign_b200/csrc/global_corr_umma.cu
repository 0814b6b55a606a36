// tcgen05 (UMMA) TF32 path of the global correlation volume for sm_100a.
//
// out[b, s, t] = sum_c src[b,c,s] * trg[b,c,t]  followed by mutual matching, ReLU and the L2-norm over
// the source dimension (GlobalFeatureCorrelationLayer.forward, /root/reference/models/modules.py:294-333,
// 362-374; the reference runs torch.bmm + ~10 elementwise / reduction passes over the [B,Ns,Nt] volume).
//
// The volume is HBM-write-bound at the sweep sizes (64 flop per output byte at C = 128), so the design
// goal is ONE pass over it: the GEMM tile is recomputed on the tensor cores instead of re-reading the
// volume for the two reductions the reference needs.  Three launches of the same kernel:
//   phase 0  tile GEMM -> row maxima A[s] (max over t) and column maxima Bm[t] (max over s); nothing stored
//   phase 1  tile GEMM -> v = c * ((c/(A+eps)) * (c/(Bm+eps))), relu, column sums of v^2 (norm over s)
//   phase 2  tile GEMM -> v, relu, / max(||.||, 1e-12) -> the only write of the volume
// Tile: 128 (s) x 128 (t) fp32 accumulator in TMEM (128 columns), K loop over 32-channel blocks.
// Both operands are read IN PLACE from the [B,C,N] layout (N contiguous), i.e. both are "MN-major"
// UMMA operands: TMA boxes of 32 positions (128 bytes) x 32 channels with the 128-byte swizzle, four
// boxes per operand tile, kind::tf32 M128 N128 K8.  MN-major 32-bit operands only exist in the
// "128-byte swizzle with 32-byte atoms" shared-memory layout (UMMA layout type 1, TMA
// CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): four K rows per swizzle group.
// 256 threads: thread t owns accumulator row t % 128 (TMEM lane) and the column half t / 128; thread 0
// also drives TMA and the MMAs.  The mutual-matching ratios use reciprocals (this is the TF32 path; the
// exact-division arithmetic of the reference lives in the fp32 FFMA path).
#include "rf_common.cuh"
#include "rf_sm100.cuh"

namespace rf {
using namespace sm100;

constexpr int GU_BK = 32;                    // channels per K block
constexpr int GU_STAGES = 3;
constexpr int GU_BOX_BYTES = GU_BK * 128;    // one 32-position x 32-channel box
constexpr int GU_OP_BYTES = 4 * GU_BOX_BYTES;  // 128 positions
constexpr int GU_STAGE_BYTES = 2 * GU_OP_BYTES;
constexpr int GU_SMEM = GU_STAGES * GU_STAGE_BYTES + 1024 + 2048;

struct __align__(8) GuShared {
  uint64_t full[GU_STAGES], empty[GU_STAGES], done;
  uint32_t tmem_base;
  float col[128];   // per-column reduction (max as ordered int / sum of squares)
  float colB[128];  // colmax + eps (phases 1, 2)
  float colN[128];  // 1 / norm (phase 2)
};

__device__ __forceinline__ void smem_atomic_max_float(float* addr, float v) {
  if (v >= 0.f)
    atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else
    atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}
__device__ __forceinline__ void gmem_atomic_max_float(float* addr, float v) {
  if (v >= 0.f)
    atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else
    atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

// Reduction ACROSS the 32 lanes of a warp of a 32-element per-lane array: on return lane l holds
// op_{lanes} w[l].  Butterfly with halving payload (16+8+4+2+1 shuffles), static register indices only.
template <bool MAX>
__device__ __forceinline__ float warp_transpose_reduce(float (&w)[32], int lane) {
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

template <int PHASE>
__global__ void __launch_bounds__(256, 2)
global_corr_umma_kernel(const __grid_constant__ CUtensorMap tm_src, const __grid_constant__ CUtensorMap tm_trg,
                        float* __restrict__ out, float* __restrict__ rowmax, float* __restrict__ colmax,
                        float* __restrict__ normsq, int C, long Ns, long Nt, int mode) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  GuShared* sh = reinterpret_cast<GuShared*>(smem + GU_STAGES * GU_STAGE_BYTES);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int b = blockIdx.z;
  const long s0 = (long)blockIdx.y * 128, t0 = (long)blockIdx.x * 128;
  const int nkb = C / GU_BK;
  const bool mm = mode & 1, nrm = mode & 2;

  if (tid == 0) {
    for (int s = 0; s < GU_STAGES; ++s) {
      mbar_init(&sh->full[s], 1);
      mbar_init(&sh->empty[s], 1);
    }
    mbar_init(&sh->done, 1);
    fence_barrier_init();
  }
  if (tid < 128) {
    if (PHASE == 0) sh->col[tid] = -INFINITY;
    if (PHASE == 1) sh->col[tid] = 0.f;
    if (PHASE >= 1) {
      const long t = t0 + tid;
      sh->colB[tid] = (mm && t < Nt) ? 1.f / (colmax[(long)b * Nt + t] + 1e-5f) : 1.f;   // 1 / (Bm + eps)
      if (PHASE == 2) sh->colN[tid] = (nrm && t < Nt) ? 1.f / fmaxf(sqrtf(normsq[(long)b * Nt + t]), 1e-12f) : 1.f;
    }
  }
  if (warp == 0) tmem_alloc<128>(&sh->tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sh->tmem_base;

  if (tid == 0) {
    tma_prefetch_desc(&tm_src);
    tma_prefetch_desc(&tm_trg);
    constexpr uint32_t IDESC = make_idesc(FMT_TF32, 128, 128, 1, 1);
    auto load_stage = [&](int kb) {
      const int st = kb % GU_STAGES;
      uint8_t* a = smem + st * GU_STAGE_BYTES;
      uint8_t* bb = a + GU_OP_BYTES;
      mbar_expect_tx(&sh->full[st], GU_STAGE_BYTES);
#pragma unroll
      for (int blk = 0; blk < 4; ++blk) {
        tma_load_3d(a + blk * GU_BOX_BYTES, &tm_src, &sh->full[st], (int)(s0 + blk * 32), kb * GU_BK, b);
        tma_load_3d(bb + blk * GU_BOX_BYTES, &tm_trg, &sh->full[st], (int)(t0 + blk * 32), kb * GU_BK, b);
      }
    };
    for (int kb = 0; kb < GU_STAGES && kb < nkb; ++kb) load_stage(kb);
    for (int kb = 0; kb < nkb; ++kb) {
      const int st = kb % GU_STAGES;
      mbar_wait(&sh->full[st], (kb / GU_STAGES) & 1);
      tc_fence_after();
      const uint64_t da = make_sdesc_sw128_base32(smem_u32(smem + st * GU_STAGE_BYTES), GU_BOX_BYTES, 512);
      const uint64_t db = make_sdesc_sw128_base32(smem_u32(smem + st * GU_STAGE_BYTES + GU_OP_BYTES), GU_BOX_BYTES, 512);
#pragma unroll
      for (int k = 0; k < GU_BK / 8; ++k)
        mma_tf32_ss(tmem, da + (uint64_t)(k * 64), db + (uint64_t)(k * 64), IDESC, (kb > 0 || k > 0) ? 1u : 0u);
      tc_commit(&sh->empty[st]);
      if (kb >= 1 && kb - 1 + GU_STAGES < nkb) {  // refill the stage consumed one iteration ago
        const int pk = kb - 1;
        mbar_wait(&sh->empty[pk % GU_STAGES], (pk / GU_STAGES) & 1);
        load_stage(pk + GU_STAGES);
      }
    }
    tc_commit(&sh->done);
  }
  __syncwarp();
  mbar_wait(&sh->done, 0);
  tc_fence_after();

  // ---------------------------------------------------------------- epilogue: thread = row s
  const int half = tid >> 7;            // warps 0-3: columns [0,64), warps 4-7: columns [64,128)
  const long s = s0 + (tid & 127);
  const bool row_ok = s < Ns;
  const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  const float ra = (PHASE >= 1 && mm && row_ok) ? 1.f / (rowmax[(long)b * Ns + s] + 1e-5f) : 1.f;   // 1 / (A + eps)
  float rmax = -INFINITY;
  float* orow = out + ((long)b * Ns + s) * Nt + t0;
#pragma unroll
  for (int cc = 0; cc < 2; ++cc) {
    const int c = half * 2 + cc;
    uint32_t v[32];
    tmem_ld32(trow + c * 32, v);
    tc_wait_ld();
    if (PHASE == 0) {
      float w[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        w[i] = row_ok ? __uint_as_float(v[i]) : -INFINITY;
        if (t0 + c * 32 + i < Nt) rmax = fmaxf(rmax, w[i]);
      }
      const float cm = warp_transpose_reduce<true>(w, tid & 31);      // lane l: max over this warp's 32 rows of column l
      smem_atomic_max_float(&sh->col[c * 32 + (tid & 31)], cm);       // 4 warps, conflict-free
    } else {
      float w[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        float x = __uint_as_float(v[i]);
        if (mm) x = x * ((x * ra) * (x * sh->colB[c * 32 + i]));
        if (nrm) x = fmaxf(x, 0.f);
        w[i] = x;
      }
      if (PHASE == 1) {
#pragma unroll
        for (int i = 0; i < 32; ++i) w[i] = row_ok ? w[i] * w[i] : 0.f;
        const float cs = warp_transpose_reduce<false>(w, tid & 31);
        atomicAdd(&sh->col[c * 32 + (tid & 31)], cs);
      } else if (row_ok) {
        const bool vec = (Nt % 4 == 0) && (t0 + c * 32 + 32 <= Nt);
        if (vec) {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            st_cs_f4(orow + c * 32 + 4 * i,
                     make_float4(w[4 * i] * sh->colN[c * 32 + 4 * i], w[4 * i + 1] * sh->colN[c * 32 + 4 * i + 1],
                                 w[4 * i + 2] * sh->colN[c * 32 + 4 * i + 2], w[4 * i + 3] * sh->colN[c * 32 + 4 * i + 3]));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (t0 + c * 32 + i < Nt) orow[c * 32 + i] = w[i] * sh->colN[c * 32 + i];
        }
      }
    }
  }
  if (PHASE == 0 && row_ok) gmem_atomic_max_float(rowmax + (long)b * Ns + s, rmax);
  tc_fence_before();
  __syncthreads();
  if (PHASE == 0 && tid < 128 && t0 + tid < Nt) gmem_atomic_max_float(colmax + (long)b * Nt + t0 + tid, sh->col[tid]);
  if (PHASE == 1 && tid < 128 && t0 + tid < Nt) atomicAdd(normsq + (long)b * Nt + t0 + tid, sh->col[tid]);
  if (warp == 0) tmem_dealloc<128>(tmem);
}

bool global_corr_umma_supported(int C, long Ns, long Nt, const void* a, const void* b, const void* c) {
  return C % GU_BK == 0 && Ns % 4 == 0 && Nt % 4 == 0 && Ns < (1l << 31) && Nt < (1l << 31) &&
         (((uintptr_t)a | (uintptr_t)b | (uintptr_t)c) & 15) == 0;
}

// rowmax / colmax: [B,Ns] / [B,Nt] pre-filled with -inf by the caller when mode & 1; normsq: [B,Nt] scratch.
int global_corr_umma(const float* src, const float* trg, float* out, float* rowmax, float* colmax, float* normsq,
                     int B, int C, long Ns, long Nt, int mode, cudaStream_t st) {
  CUtensorMap ts, tt;
  int rc = make_tmap_3d(&ts, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, src, (uint64_t)Ns, (uint64_t)C, (uint64_t)B,
                        (uint64_t)Ns * 4, (uint64_t)C * Ns * 4, 32, GU_BK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  if (rc != RF_OK) return rc;
  rc = make_tmap_3d(&tt, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, trg, (uint64_t)Nt, (uint64_t)C, (uint64_t)B,
                    (uint64_t)Nt * 4, (uint64_t)C * Nt * 4, 32, GU_BK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  if (rc != RF_OK) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    RF_CUDA(cudaFuncSetAttribute(global_corr_umma_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, GU_SMEM));
    RF_CUDA(cudaFuncSetAttribute(global_corr_umma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, GU_SMEM));
    RF_CUDA(cudaFuncSetAttribute(global_corr_umma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, GU_SMEM));
    attr_set = true;
  }
  dim3 grid((unsigned)ceil_div(Nt, 128), (unsigned)ceil_div(Ns, 128), (unsigned)B);
  if (mode & 1) {
    global_corr_umma_kernel<0><<<grid, 256, GU_SMEM, st>>>(ts, tt, out, rowmax, colmax, normsq, C, Ns, Nt, mode);
    RF_CHECK_LAUNCH("global_corr_umma_kernel<0>");
  }
  if (mode & 2) {
    RF_CUDA(cudaMemsetAsync(normsq, 0, sizeof(float) * (size_t)B * Nt, st));
    global_corr_umma_kernel<1><<<grid, 256, GU_SMEM, st>>>(ts, tt, out, rowmax, colmax, normsq, C, Ns, Nt, mode);
    RF_CHECK_LAUNCH("global_corr_umma_kernel<1>");
  }
  global_corr_umma_kernel<2><<<grid, 256, GU_SMEM, st>>>(ts, tt, out, rowmax, colmax, normsq, C, Ns, Nt, mode);
  RF_CHECK_LAUNCH("global_corr_umma_kernel<2>");
  return RF_OK;
}

}  // namespace rf
