// Global (4-D) correlation volume with mutual matching, ReLU and L2-norm.
//
// Restates GlobalFeatureCorrelationLayer.forward
// (/root/reference/models/modules.py:294-308): the bmm of :362-374
// (out[b, s, t] = <src[b,:,s], trg[b,:,t]>, s row-major over the source map),
// mutual_matching (:310-333, eps 1e-5, corr * (corr/(maxA+eps) * corr/(maxB+eps)))
// and relu + F.normalize over the source dimension (:307).  The reference runs
// one cuBLAS bmm plus ~10 elementwise/reduction kernels over the volume; here:
//   pass 1  GEMM tile kernel, epilogue folds the row / column maxima
//           (tensor-core path: global_corr_umma.cu, tcgen05 TF32 with the
//            accumulator in TMEM; this file: fp32 FFMA path for any shape);
//   pass 2  per target column: apply the mutual-matching ratio + ReLU,
//           accumulate the squared norm over s, then rescale in place.
// Both operands are read in their native [B,C,N] layout (N contiguous), i.e.
// no transposed copies.
#include <stdlib.h>

#include "rf_common.cuh"

namespace rf {

constexpr int GC_TS = 64, GC_TT = 64, GC_CK = 16;

__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
  // total order trick: non-negative floats compare as ints, negatives reversed as uints
  if (v >= 0.f)
    atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else
    atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

__global__ void fill_kernel(float* p, float v, long n) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) p[i] = v;
}

// out tile 64(s) x 64(t); 256 threads, 4x4 outputs each.
__global__ void __launch_bounds__(256)
global_corr_ffma_kernel(const float* __restrict__ src, const float* __restrict__ trg, float* __restrict__ out,
                        float* __restrict__ rowmax, float* __restrict__ colmax, int C, long Ns, long Nt) {
  __shared__ __align__(16) float As[GC_CK][GC_TS];
  __shared__ __align__(16) float Bs[GC_CK][GC_TT];
  const int b = blockIdx.z;
  const long s0 = (long)blockIdx.y * GC_TS, t0 = (long)blockIdx.x * GC_TT;
  const float* S = src + (long)b * C * Ns;
  const float* T = trg + (long)b * C * Nt;
  const int tt = threadIdx.x & 15, ts = threadIdx.x >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int c0 = 0; c0 < C; c0 += GC_CK) {
    for (int i = threadIdx.x; i < GC_CK * GC_TS; i += 256) {
      const int c = i / GC_TS, k = i % GC_TS;
      As[c][k] = (c0 + c < C && s0 + k < Ns) ? __ldg(S + (long)(c0 + c) * Ns + s0 + k) : 0.f;
      Bs[c][k] = (c0 + c < C && t0 + k < Nt) ? __ldg(T + (long)(c0 + c) * Nt + t0 + k) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < GC_CK; ++c) {
      const float4 a = *reinterpret_cast<const float4*>(&As[c][4 * ts]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[c][4 * tt]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* O = out + (long)b * Ns * Nt;
  float cm[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long s = s0 + 4 * ts + i;
    float rm = -INFINITY;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long t = t0 + 4 * tt + j;
      if (s < Ns && t < Nt) {
        O[s * Nt + t] = acc[i][j];
        rm = fmaxf(rm, acc[i][j]);
        cm[j] = fmaxf(cm[j], acc[i][j]);
      }
    }
    if (rowmax) {
      // the 16 threads sharing `ts` are 16 consecutive lanes
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) rm = fmaxf(rm, __shfl_xor_sync(0xffffffffu, rm, o));
      if (tt == 0 && s < Ns) atomic_max_float(rowmax + (long)b * Ns + s, rm);
    }
  }
  if (colmax) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      // lanes l and l^16 hold the same tt, different ts
      cm[j] = fmaxf(cm[j], __shfl_xor_sync(0xffffffffu, cm[j], 16));
      const long t = t0 + 4 * tt + j;
      if ((threadIdx.x & 16) == 0 && t < Nt) atomic_max_float(colmax + (long)b * Nt + t, cm[j]);
    }
  }
}

// pass 2: one CTA column-group: threads x = 32 consecutive t, y = GC2_SY slices of s.
constexpr int GC2_SY = 8;
__global__ void __launch_bounds__(32 * GC2_SY)
global_corr_finish_kernel(float* __restrict__ out, const float* __restrict__ rowmax, const float* __restrict__ colmax,
                          long Ns, long Nt, int mode) {
  __shared__ float part[GC2_SY][32];
  const int b = blockIdx.y;
  const long t = (long)blockIdx.x * 32 + threadIdx.x;
  float* O = out + (long)b * Ns * Nt;
  const bool mm = mode & 1, nrm = mode & 2;
  const bool live = t < Nt;
  const float cb = (mm && live) ? colmax[(long)b * Nt + t] + 1e-5f : 1.f;
  float ss = 0.f;
  if (live) {
    for (long s = threadIdx.y; s < Ns; s += GC2_SY) {
      float v = O[s * Nt + t];
      if (mm) {
        const float ra = v / (rowmax[(long)b * Ns + s] + 1e-5f);
        const float rb = v / cb;
        v = v * (ra * rb);
      }
      if (nrm) v = fmaxf(v, 0.f);
      ss = fmaf(v, v, ss);
      if (mm || nrm) O[s * Nt + t] = v;
    }
  }
  if (!nrm) return;
  part[threadIdx.y][threadIdx.x] = ss;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int y = 0; y < GC2_SY; ++y) tot += part[y][threadIdx.x];
  const float d = fmaxf(sqrtf(tot), 1e-12f);
  if (live)
    for (long s = threadIdx.y; s < Ns; s += GC2_SY) O[s * Nt + t] = O[s * Nt + t] / d;
}

int global_corr_umma(const float* src, const float* trg, float* out, float* rowmax, float* colmax, float* normsq,
                     int B, int C, long Ns, long Nt, int mode, cudaStream_t st);  // global_corr_umma.cu
bool global_corr_umma_supported(int C, long Ns, long Nt, const void* a, const void* b, const void* c);
// global_corr_persist.cu
int global_corr_persist(const float* src, const float* trg, float* out, float* rowmax, float* colmax, float* normsq,
                        int B, int C, long Ns, long Nt, int mode, cudaStream_t st);
bool global_corr_persist_supported(int C, long Ns, long Nt, const void* a, const void* b, const void* c);

}  // namespace rf

using namespace rf;

extern "C" int64_t rf_global_corr_workspace_bytes(int B, int64_t Ns, int64_t Nt) {
  return (int64_t)sizeof(float) * B * (Ns + 2 * Nt);  // row maxima, column maxima, column norms
}

extern "C" int rf_global_corr_fwd(const float* src, const float* trg, float* out, void* workspace, int B, int C,
                                  int64_t Ns, int64_t Nt, int mode, int use_tensor_cores, void* stream) {
  RF_REQUIRE(src && trg && out, "rf_global_corr_fwd: null pointer");
  RF_REQUIRE(B > 0 && B <= 65535 && C > 0 && Ns > 0 && Nt > 0, "rf_global_corr_fwd: bad shape");
  RF_REQUIRE(!(mode & 1) || workspace, "rf_global_corr_fwd: mutual matching needs a workspace");
  cudaStream_t st = (cudaStream_t)stream;
  float* rowmax = (mode & 1) ? (float*)workspace : nullptr;
  float* colmax = (mode & 1) ? rowmax + (long)B * Ns : nullptr;
  if (mode & 1) {
    const long n = (long)B * (Ns + Nt);
    fill_kernel<<<(int)((n + 255) / 256 < 1024 ? (n + 255) / 256 : 1024), 256, 0, st>>>(rowmax, -INFINITY, n);
    RF_CHECK_LAUNCH("fill_kernel");
  }
  const bool can_tc = global_corr_umma_supported(C, Ns, Nt, src, trg, out);
  const bool can_pp = global_corr_persist_supported(C, Ns, Nt, src, trg, out);
  RF_REQUIRE(use_tensor_cores != 1 || can_tc,
             "rf_global_corr_fwd: tcgen05 path needs C%%32==0, Ns%%4==0, Nt%%4==0 and 16B-aligned pointers");
  RF_REQUIRE(use_tensor_cores != 2 || can_pp,
             "rf_global_corr_fwd: persistent tcgen05 path needs C%%32==0, C<=128, Ns%%4==0, Nt%%4==0 and 16B-aligned pointers");
  RF_REQUIRE(use_tensor_cores <= 2, "rf_global_corr_fwd: use_tensor_cores must be -1 (auto), 0, 1 or 2");
  // automatic choice: the tensor-core paths pay off once the volume no longer fits a handful of CTAs; the
  // persistent kernel (source tile resident, C <= 128) is the sweep-size path
  static const bool persist_off = [] { const char* e = getenv("RF_GCORR_PERSIST"); return e && e[0] == '0'; }();
  const bool big = Ns * Nt >= (1l << 20);
  if (use_tensor_cores == 2 || (use_tensor_cores < 0 && can_pp && big && !persist_off)) {
    RF_REQUIRE(workspace != nullptr, "rf_global_corr_fwd: the tcgen05 path needs the workspace");
    float* ws = (float*)workspace;
    return global_corr_persist(src, trg, out, ws, ws + (long)B * Ns, ws + (long)B * (Ns + Nt), B, C, Ns, Nt, mode, st);
  }
  if (use_tensor_cores == 1 || (use_tensor_cores < 0 && can_tc && big)) {
    RF_REQUIRE(workspace != nullptr, "rf_global_corr_fwd: the tcgen05 path needs the workspace");
    float* ws = (float*)workspace;
    return global_corr_umma(src, trg, out, ws, ws + (long)B * Ns, ws + (long)B * (Ns + Nt), B, C, Ns, Nt, mode, st);
  }
  {
    dim3 grid((unsigned)ceil_div(Nt, GC_TT), (unsigned)ceil_div(Ns, GC_TS), B);
    global_corr_ffma_kernel<<<grid, 256, 0, st>>>(src, trg, out, rowmax, colmax, C, Ns, Nt);
    RF_CHECK_LAUNCH("global_corr_ffma_kernel");
  }
  if (mode & 3) {
    dim3 grid((unsigned)ceil_div(Nt, 32), B), block(32, GC2_SY);
    global_corr_finish_kernel<<<grid, block, 0, st>>>(out, rowmax, colmax, Ns, Nt, mode);
    RF_CHECK_LAUNCH("global_corr_finish_kernel");
  }
  return RF_OK;
}
