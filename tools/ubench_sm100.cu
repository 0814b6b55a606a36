// Micro-benchmarks of the sm_100a resources the attention / GEMM kernels are designed around.  Not product code:
// built by tools/build_ubench.sh into tools/_bin/ubench_sm100 and run on the GPU box; results are committed under
// profiles/ and quoted in DESIGN.md.  Measures (per SM, in SM clocks):
//   tmem_ld / tmem_st : tcgen05.ld / tcgen05.st bandwidth with 4, 8, 16 warps
//   mufu              : ex2.approx throughput
//   mma               : tcgen05.mma issue-to-completion rate for SS / TS operand modes and N = 64 / 128 / 256
//   tma_red           : cp.reduce.async.bulk.tensor (fp32 add) chip throughput, distinct and 8-way shared tiles
//   red_v4            : red.global.add.v4.f32 chip throughput
#include <cuda_bf16.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../refign_b200/csrc/rf_common.cuh"
#include "../refign_b200/csrc/rf_sm100.cuh"

namespace rf {
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vfprintf(stderr, fmt, ap);
  va_end(ap);
  fputc('\n', stderr);
}
}  // namespace rf
using namespace rf;
using namespace rf::sm100;

#define CK(x)                                                                         \
  do {                                                                                \
    cudaError_t e_ = (x);                                                             \
    if (e_ != cudaSuccess) {                                                          \
      fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
      exit(1);                                                                        \
    }                                                                                 \
  } while (0)

// ------------------------------------------------------------------------------------------------ TMEM ld / st
template <int MODE>  // 0 = ld, wait once per 4 loads; 1 = ld, wait after every load; 2 = st
__global__ void tmem_rw_kernel(int iters, long long* cycles, uint32_t* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc<512>(&slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) & 3) * 128;
  uint32_t acc = 0;
  uint32_t v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = threadIdx.x + i;
  if (MODE != 2) {  // defined contents
    for (int c = 0; c < 4; ++c) tmem_st32(tmem + c * 32, v);
    tc_wait_st();
  }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 2) {
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_st32(tmem + c * 32, v);
      tc_wait_st();
    } else {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        tmem_ld32(tmem + c * 32, v);
        if (MODE == 1) {
          tc_wait_ld();
          acc ^= v[0] ^ v[31];
        }
      }
      if (MODE == 0) {
        tc_wait_ld();
        acc ^= v[0] ^ v[31];
      }
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678u) sink[0] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(slot);
}

// ------------------------------------------------------------------------------------------------ MUFU
__global__ void mufu_kernel(int iters, long long* cycles, float* sink) {
  float x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = -0.001f * (threadIdx.x + i);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = fast_exp2(x[i]) - 1.0f;
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
  if (s == 1234.5f) sink[0] = s;
}

// ------------------------------------------------------------------------------------------------ MMA rate
// MODE 0: SS N=128   1: SS N=64   2: TS N=64   3: SS N=256   4: SS N=64 with MN-major B   5: TS N=128
template <int MODE>
__global__ void __launch_bounds__(128, 1) mma_kernel(int rounds, int per_round, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  // non-trivial operand bytes (bf16 ~ 1.0 / 0.5 patterns)
  for (int i = tid; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3f803f00u + (i & 0x7f);
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<512>(&slot);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = uniform_u32(slot);
  constexpr int N = (MODE == 0 || MODE == 5) ? 128 : (MODE == 3 ? 256 : 64);
  constexpr uint32_t IDESC = make_idesc(FMT_BF16, 128, N, 0, MODE == 4 ? 1 : 0);
  const uint64_t dA = make_sdesc_sw128(smem_u32(smem), 16, 1024);
  const uint64_t dB = MODE == 4 ? make_sdesc_sw128(smem_u32(smem + 16384), 8192, 1024)
                                : make_sdesc_sw128(smem_u32(smem + 16384), 16, 1024);
  long long t0 = 0, t1 = 0;
  if (warp == 0) {
    t0 = clock64();
    for (int r = 0; r < rounds; ++r) {
      if (elect_one()) {
        for (int i = 0; i < per_round; ++i) {
          const int k = i & 3;
          if (MODE == 2 || MODE == 5)
            mma_f16_ts(tmem, tmem + 256 + k * 8, dB + (uint64_t)(k * 2), IDESC, 1u);
          else if (MODE == 4)
            mma_f16_ss(tmem, dA + (uint64_t)(k * 2), dB + (uint64_t)(k * 128), IDESC, 1u);
          else
            mma_f16_ss(tmem, dA + (uint64_t)(k * 2), dB + (uint64_t)(k * 2), IDESC, 1u);
        }
        tc_commit(&bar);
      }
      __syncwarp();
      mbar_wait(&bar, r & 1);
    }
    t1 = clock64();
  }
  if (tid == 0) cycles[blockIdx.x] = t1 - t0;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}


// ------------------------------------------------------------------------------------------------ commit latency
// Does the mbarrier of a tcgen05.commit fire when ITS batch completes, while the issuing warp then waits (mbarrier
// spin) and other warps keep the SM busy?   warp 0 = MMA issuer, warp 1 = observer, warps 2..9 = load generators.
// LOAD 0: idle   1: FMA + MUFU loop   2: tcgen05.ld loop   WAITMODE 0: issuer spins on clock64   1: issuer waits on an mbarrier
template <int LOAD, int WAITMODE>
__global__ void __launch_bounds__(320, 1) commit_latency_kernel(long long* out, int gap) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t barA, barB, barX;
  __shared__ uint32_t slot;
  __shared__ long long t_issue[2];
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (16384 + 32768) / 4; i += 320) reinterpret_cast<uint32_t*>(smem)[i] = 0x3f803f00u + (i & 0x7f);
  if (tid == 0) {
    mbar_init(&barA, 1);
    mbar_init(&barB, 1);
    mbar_init(&barX, 256);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<512>(&slot);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = uniform_u32(slot);
  constexpr uint32_t IDESC_S = make_idesc(FMT_BF16, 128, 128, 0, 0);
  constexpr uint32_t IDESC_A = make_idesc(FMT_BF16, 128, 64, 0, 1);
  const uint64_t dA = make_sdesc_sw128(smem_u32(smem), 16, 1024);
  const uint64_t dB = make_sdesc_sw128(smem_u32(smem + 16384), 16, 1024);
  const uint64_t dBmn = make_sdesc_sw128(smem_u32(smem + 16384), 8192, 1024);
  const long long t_start = clock64();
  if (warp == 0) {
    if (elect_one()) {
      t_issue[0] = clock64() - t_start;
      for (int k = 0; k < 4; ++k) mma_f16_ss(tmem, dA + (uint64_t)(k * 2), dB + (uint64_t)(k * 2), IDESC_S, k > 0);
      for (int k = 0; k < 4; ++k) mma_f16_ss(tmem + 128, dA + (uint64_t)(k * 2), dB + (uint64_t)(k * 2), IDESC_S, k > 0);
      tc_commit(&barA);
    }
    __syncwarp();
    if (WAITMODE == 0) {
      while (clock64() - t_start < gap) {}
    } else {
      mbar_wait(&barX, 0);
    }
    tc_fence_after();
    if (elect_one()) {
      t_issue[1] = clock64() - t_start;
      for (int k = 0; k < 8; ++k) mma_f16_ts(tmem + 448, tmem + 256 + k * 8, dBmn + (uint64_t)(k * 128), IDESC_A, k > 0);
      for (int k = 0; k < 8; ++k) mma_f16_ts(tmem + 384, tmem + 320 + k * 8, dBmn + (uint64_t)(k * 128), IDESC_A, k > 0);
      tc_commit(&barB);
    }
    __syncwarp();
  } else if (warp == 1) {
    mbar_wait(&barA, 0);
    const long long ta = clock64() - t_start;
    mbar_wait(&barB, 0);
    const long long tb = clock64() - t_start;
    if ((tid & 31) == 0 && blockIdx.x == 0) {
      out[0] = ta;
      out[1] = tb;
    }
  } else {
    const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    float x = 0.001f * tid, y = 0.f;
    uint32_t v[32];
    uint32_t acc = 0;
    while (clock64() - t_start < gap) {
      if (LOAD == 1) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          x = fast_exp2(x) - 1.0f;
          y = fmaf(y, 1.0001f, x);
        }
      } else if (LOAD == 2) {
        tmem_ld32(taddr, v);
        tc_wait_ld();
        acc ^= v[3];
      }
    }
    if (x + y == 123.f || acc == 77u) out[7] = 1;
    tc_fence_before();
    mbar_arrive(&barX);
  }
  __syncthreads();
  if (tid == 0 && blockIdx.x == 0) {
    out[2] = t_issue[0];
    out[3] = t_issue[1];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}


// ------------------------------------------------------------------------------------------------ LDTM under MMA load
// 8 warps run the softmax warps' load pattern (4 x tcgen05.ld.32x32b.x32 + wait) while warp 8 keeps the tensor pipe busy.
// MMAMODE 0: tensor pipe idle   1: SS N=128 MMAs   2: TS N=64 MMAs (A operand read from TMEM)
template <int MMAMODE>
__global__ void __launch_bounds__(288, 1) ldtm_under_mma_kernel(long long* out, int span) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  __shared__ int iters[8];
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (16384 + 32768) / 4; i += 288) reinterpret_cast<uint32_t*>(smem)[i] = 0x3f803f00u + (i & 0x7f);
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<512>(&slot);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = uniform_u32(slot);
  const long long t_start = clock64();
  if (warp == 8) {
    constexpr uint32_t IDESC_S = make_idesc(FMT_BF16, 128, 128, 0, 0);
    constexpr uint32_t IDESC_A = make_idesc(FMT_BF16, 128, 64, 0, 1);
    const uint64_t dA = make_sdesc_sw128(smem_u32(smem), 16, 1024);
    const uint64_t dB = make_sdesc_sw128(smem_u32(smem + 16384), 16, 1024);
    const uint64_t dBmn = make_sdesc_sw128(smem_u32(smem + 16384), 8192, 1024);
    int r = 0;
    while (MMAMODE != 0 && clock64() - t_start < span) {
      if (elect_one()) {
        for (int k = 0; k < 8; ++k) {
          if (MMAMODE == 1) mma_f16_ss(tmem + 256, dA + (uint64_t)((k & 3) * 2), dB + (uint64_t)((k & 3) * 2), IDESC_S, 1u);
          else mma_f16_ts(tmem + 448, tmem + 384 + k * 8, dBmn + (uint64_t)(k * 128), IDESC_A, 1u);
        }
        tc_commit(&bar);
      }
      __syncwarp();
      mbar_wait(&bar, r & 1);
      ++r;
    }
  } else {
    const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 64;
    uint32_t a[32], b[32], c[32], d[32];
    uint32_t acc = 0;
    int n = 0;
    while (clock64() - t_start < span) {
      tmem_ld32(taddr, a);
      tmem_ld32(taddr + 32, b);
      tmem_ld32(taddr + 128, c);
      tmem_ld32(taddr + 160, d);
      tc_wait_ld();
      acc ^= a[1] ^ b[2] ^ c[3] ^ d[4];
      ++n;
    }
    if ((tid & 31) == 0) iters[warp] = n;
    if (acc == 77u) out[7] = 1;
  }
  __syncthreads();
  if (tid == 0 && blockIdx.x == 0) {
    int tot = 0;
    for (int w = 0; w < 8; ++w) tot += iters[w];
    out[0] = tot;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

// ------------------------------------------------------------------------------------------------ TMA reduce
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }

// each CTA reduces [128 rows x 32 fp32] tiles (16 KiB) into a [rows][64] fp32 tensor; share = how many CTAs hit the
// same tile sequence (1 = all distinct, 8 = the dQ case of 8 key blocks)
__global__ void __launch_bounds__(128, 1) tma_red_kernel(const __grid_constant__ CUtensorMap tm, int iters, int rows,
                                                         int share) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  for (int i = threadIdx.x; i < 2 * 16384 / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 1.0f;
  fence_proxy_async();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int group = blockIdx.x / share;
    const int ngroups = (gridDim.x + share - 1) / share;
    for (int it = 0; it < iters; ++it) {
      const long tile = (long)it * ngroups + group;
      const int row = (int)((tile * 128) % rows);
      tma_reduce_add_3d(&tm, smem, 0, row, 0);
      tma_reduce_add_3d(&tm, smem + 16384, 32, row, 0);
      bulk_commit();
      bulk_wait_read<1>();
    }
    bulk_wait_read<0>();
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

__global__ void red_v4_kernel(float* base, int iters, long rows_mask) {
  const long t = blockIdx.x * (long)blockDim.x + threadIdx.x;
  const long nthreads = (long)gridDim.x * blockDim.x;
  for (int it = 0; it < iters; ++it) {
    const long idx = ((t + (long)it * nthreads) * 4) & rows_mask;
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(base + idx), "f"(1.f), "f"(1.f), "f"(1.f), "f"(1.f)
                 : "memory");
  }
}

static double max_cycles(long long* d, int n) {
  std::vector<long long> h(n);
  CK(cudaMemcpy(h.data(), d, n * sizeof(long long), cudaMemcpyDeviceToHost));
  long long m = 0;
  for (auto v : h) m = v > m ? v : m;
  return (double)m;
}

int main() {
  CK(cudaSetDevice(0));
  CK(cudaFree(0));
  long long* cyc;
  uint32_t* sink;
  CK(cudaMalloc(&cyc, 148 * sizeof(long long)));
  CK(cudaMalloc(&sink, 16));
  printf("{\n");
  // ---- TMEM
  const int iters = 4000;
  for (int warps : {4, 8, 16}) {
    for (int mode = 0; mode < 3; ++mode) {
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) tmem_rw_kernel<0><<<148, warps * 32>>>(iters, cyc, sink);
        if (mode == 1) tmem_rw_kernel<1><<<148, warps * 32>>>(iters, cyc, sink);
        if (mode == 2) tmem_rw_kernel<2><<<148, warps * 32>>>(iters, cyc, sink);
        CK(cudaDeviceSynchronize());
      }
      const double c = max_cycles(cyc, 148);
      const double bytes = (double)iters * warps * 4 * 32 * 32 * 4;
      printf(" \"tmem_%s_w%d_bytes_per_clk_per_sm\": %.1f,\n", mode == 0 ? "ld_wait4" : (mode == 1 ? "ld_wait1" : "st"), warps,
             bytes / c);
    }
  }
  // ---- MUFU
  for (int warps : {4, 8, 16, 32}) {
    for (int rep = 0; rep < 2; ++rep) {
      mufu_kernel<<<148, warps * 32>>>(2000, cyc, (float*)sink);
      CK(cudaDeviceSynchronize());
    }
    const double c = max_cycles(cyc, 148);
    printf(" \"mufu_ex2_w%d_per_clk_per_sm\": %.2f,\n", warps, 2000.0 * 8 * warps * 32 / c);
  }
  // ---- MMA
  {
    const int smem = 16384 + 32768 + 1024;
    CK(cudaFuncSetAttribute(mma_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaFuncSetAttribute(mma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaFuncSetAttribute(mma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaFuncSetAttribute(mma_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaFuncSetAttribute(mma_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaFuncSetAttribute(mma_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const char* names[6] = {"ss_n128", "ss_n64", "ts_n64", "ss_n256", "ss_n64_bmn", "ts_n128"};
    for (int mode = 0; mode < 6; ++mode) {
      for (int per : {4, 32}) {
        const int rounds = 512;
        for (int rep = 0; rep < 2; ++rep) {
          if (mode == 0) mma_kernel<0><<<148, 128, smem>>>(rounds, per, cyc);
          if (mode == 1) mma_kernel<1><<<148, 128, smem>>>(rounds, per, cyc);
          if (mode == 2) mma_kernel<2><<<148, 128, smem>>>(rounds, per, cyc);
          if (mode == 3) mma_kernel<3><<<148, 128, smem>>>(rounds, per, cyc);
          if (mode == 4) mma_kernel<4><<<148, 128, smem>>>(rounds, per, cyc);
          if (mode == 5) mma_kernel<5><<<148, 128, smem>>>(rounds, per, cyc);
          CK(cudaDeviceSynchronize());
        }
        printf(" \"mma_%s_batch%d_clk_per_mma\": %.1f,\n", names[mode], per, max_cycles(cyc, 148) / ((double)rounds * per));
      }
    }
  }
  // ---- commit latency
  {
    const int smem = 16384 + 32768 + 1024;
    long long* out;
    CK(cudaMalloc(&out, 64));
#define RUN_CL(L, W)                                                                                                   \
  {                                                                                                                    \
    CK(cudaFuncSetAttribute(commit_latency_kernel<L, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));          \
    long long h[4];                                                                                                    \
    for (int rep = 0; rep < 2; ++rep) {                                                                                \
      commit_latency_kernel<L, W><<<148, 320, smem>>>(out, 4000);                                                      \
      CK(cudaDeviceSynchronize());                                                                                     \
    }                                                                                                                  \
    CK(cudaMemcpy(h, out, 32, cudaMemcpyDeviceToHost));                                                                \
    printf(" \"commit_latency_load%d_wait%d\": {\"issueA\": %lld, \"barA_seen\": %lld, \"issueB\": %lld, \"barB_seen\": %lld},\n", L, W, \
           h[2], h[0], h[3], h[1]);                                                                                    \
  }
    RUN_CL(0, 0) RUN_CL(0, 1) RUN_CL(1, 0) RUN_CL(1, 1) RUN_CL(2, 0) RUN_CL(2, 1)
  }
  // ---- LDTM bandwidth while the tensor pipe runs
  {
    const int smem = 16384 + 32768 + 1024;
    long long* out;
    CK(cudaMalloc(&out, 64));
#define RUN_LM(M)                                                                                              \
  {                                                                                                            \
    CK(cudaFuncSetAttribute(ldtm_under_mma_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));     \
    long long h[1];                                                                                            \
    for (int rep = 0; rep < 2; ++rep) {                                                                        \
      ldtm_under_mma_kernel<M><<<148, 288, smem>>>(out, 200000);                                               \
      CK(cudaDeviceSynchronize());                                                                             \
    }                                                                                                          \
    CK(cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost));                                                         \
    printf(" \"ldtm_x32_8warps_mma_mode%d_bytes_per_clk_per_sm\": %.1f,\n", M, h[0] * 16384.0 / 200000.0);      \
  }
    RUN_LM(0) RUN_LM(1) RUN_LM(2)
  }
  // ---- TMA reduce-add and red.v4
  {
    const int rows = 1 << 20;  // [rows][64] fp32 = 256 MiB
    float* buf;
    CK(cudaMalloc(&buf, (size_t)rows * 64 * 4));
    CK(cudaMemset(buf, 0, (size_t)rows * 64 * 4));
    CUtensorMap tm;
    if (make_tmap_3d(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, buf, 64, rows, 1, 64 * 4, (uint64_t)rows * 64 * 4, 32, 128) != 0) return 1;
    CK(cudaFuncSetAttribute(tma_red_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 16384 + 1024));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int share : {1, 8}) {
      const int it = 2000;
      float ms = 0;
      for (int rep = 0; rep < 2; ++rep) {
        CK(cudaEventRecord(e0));
        tma_red_kernel<<<148, 128, 2 * 16384 + 1024>>>(tm, it, rows, share);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        CK(cudaEventElapsedTime(&ms, e0, e1));
      }
      printf(" \"tma_reduce_add_f32_share%d_gbs\": %.1f,\n", share, 148.0 * it * 32768 / (ms * 1e6));
    }
    {
      std::vector<float> h(64);
      CK(cudaMemcpy(h.data(), buf, 64 * 4, cudaMemcpyDeviceToHost));
      printf(" \"tma_reduce_add_sample\": %.1f,\n", h[0]);
    }
    for (int blocks : {148 * 4, 148 * 16}) {
      const int it = 200;
      float ms = 0;
      for (int rep = 0; rep < 2; ++rep) {
        CK(cudaEventRecord(e0));
        red_v4_kernel<<<blocks, 256>>>(buf, it, (long)rows * 64 - 1);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        CK(cudaEventElapsedTime(&ms, e0, e1));
      }
      printf(" \"red_v4_f32_blocks%d_gbs\": %.1f,\n", blocks, (double)blocks * 256 * it * 16 / (ms * 1e6));
    }
  }
  printf(" \"done\": 1\n}\n");
  return 0;
}
