// Training-mode BatchNorm (+ ReLU) over channels-last activations [rows = B*H*W, C] for sm_100a:
// the BN + ReLU that follows every convolution of the DAFormer head (ConvModule / depthwise-separable
// ASPP branches and the 3x3 bottleneck: /root/reference/models/modules.py:16-56,
// /root/reference/models/heads/daformer.py:26-35,102-108; SyncBatchNorm under DDP).
// ATen's channels-last batch-norm kernels run ~8x off the HBM roofline at these shapes
// ([2..4, 1024, 256, 256] bf16); here every pass is one coalesced 16-byte-vector sweep:
//   forward : bn_reduce (per-channel sum, sum of squares)  ->  [all-reduce for SyncBN, by the caller]
//             -> bn_finalize (mean, rstd, scale/shift, running statistics)  ->  bn_apply (+ ReLU)
//   backward: bn_bwd_reduce (sum g, sum g * xhat; the ReLU mask is recomputed from x)  ->  [all-reduce]
//             -> bn_bwd_apply  dx = a (g - mean(g) - xhat mean(g xhat))
// Reductions: CTA = 32 channel groups (8 channels each; lanes along channels) x 8 row lanes, fp32 register
// partials, shared-memory combine, one red.global per channel per CTA.
#include <cuda_bf16.h>

#include "rf_common.cuh"

namespace rf {

constexpr int BN_CG = 32, BN_RL = 8;

template <typename T>
__device__ __forceinline__ void bn_load8(const T* p, float (&v)[8]);
template <>
__device__ __forceinline__ void bn_load8<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[8]) {
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}
template <>
__device__ __forceinline__ void bn_load8<float>(const float* p, float (&v)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void bn_store8(__nv_bfloat16* p, const float (&v)[8]) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ void bn_store8(float* p, const float (&v)[8]) {
  reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}

// MODE 0: out[c] += sum x,           out[C + c] += sum x^2
// MODE 1: out[c] += sum g,           out[C + c] += sum g * xhat,   g = dy * (relu ? (x a + b > 0) : 1)
template <typename T, int MODE>
__global__ void __launch_bounds__(BN_CG * BN_RL)
bn_reduce_kernel(const T* __restrict__ x, const T* __restrict__ dy, const float* __restrict__ mean,
                 const float* __restrict__ rstd, const float* __restrict__ scale, const float* __restrict__ shift,
                 float* __restrict__ out, long rows, int C, long strip, int relu) {
  __shared__ float red[16][BN_CG];
  const int lane_cg = threadIdx.x % BN_CG, rl = threadIdx.x / BN_CG;
  const int cg = blockIdx.x * BN_CG + lane_cg;
  const bool live = cg * 8 < C;
  for (int i = threadIdx.x; i < 16 * BN_CG; i += BN_CG * BN_RL) (&red[0][0])[i] = 0.f;
  __syncthreads();
  if (live) {
    float a0[8], a1[8], mu[8], rs[8], sc[8], sh[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      a0[k] = a1[k] = 0.f;
      if (MODE == 1) {
        mu[k] = __ldg(mean + cg * 8 + k);
        rs[k] = __ldg(rstd + cg * 8 + k);
        sc[k] = __ldg(scale + cg * 8 + k);
        sh[k] = __ldg(shift + cg * 8 + k);
      }
    }
    const long r0 = (long)blockIdx.y * strip;
    const long r1 = (r0 + strip < rows) ? r0 + strip : rows;
    // four rows in flight per thread (independent 16-byte loads), then the tail
    constexpr int U = 4;
    long r = r0 + rl;
    for (; r + (U - 1) * BN_RL < r1; r += U * BN_RL) {
      float v[U][8], g[U][8];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        bn_load8<T>(x + (r + u * BN_RL) * C + cg * 8, v[u]);
        if (MODE == 1) bn_load8<T>(dy + (r + u * BN_RL) * C + cg * 8, g[u]);
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          if (MODE == 0) {
            a0[k] += v[u][k];
            a1[k] = fmaf(v[u][k], v[u][k], a1[k]);
          } else {
            const float gg = (relu && fmaf(v[u][k], sc[k], sh[k]) <= 0.f) ? 0.f : g[u][k];
            a0[k] += gg;
            a1[k] = fmaf(gg, (v[u][k] - mu[k]) * rs[k], a1[k]);
          }
        }
    }
    for (; r < r1; r += BN_RL) {
      float v[8];
      bn_load8<T>(x + r * C + cg * 8, v);
      if (MODE == 0) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          a0[k] += v[k];
          a1[k] = fmaf(v[k], v[k], a1[k]);
        }
      } else {
        float g[8];
        bn_load8<T>(dy + r * C + cg * 8, g);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float gg = (relu && fmaf(v[k], sc[k], sh[k]) <= 0.f) ? 0.f : g[k];
          a0[k] += gg;
          a1[k] = fmaf(gg, (v[k] - mu[k]) * rs[k], a1[k]);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      atomicAdd(&red[k][lane_cg], a0[k]);
      atomicAdd(&red[8 + k][lane_cg], a1[k]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 16 * BN_CG; i += BN_CG * BN_RL) {
    const int e = i / BN_CG, l = i % BN_CG;
    const int c = (blockIdx.x * BN_CG + l) * 8 + (e & 7);
    if (c < C) atomicAdd(out + (e < 8 ? 0 : C) + c, red[e][l]);
  }
}

// sums[0..C) = sum x, sums[C..2C) = sum x^2 over `count` samples (already all-reduced for SyncBN)
__global__ void bn_finalize_kernel(const float* __restrict__ sums, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ mean, float* __restrict__ rstd,
                                   float* __restrict__ scale, float* __restrict__ shift, float* __restrict__ run_mean,
                                   float* __restrict__ run_var, int C, float count, float eps, float momentum) {
  if (count <= 0.f) count = sums[2 * C];   // SyncBN with per-rank sample counts: the all-reduced count travels behind the sums
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float m = sums[c] / count;
  const float var = fmaxf(sums[C + c] / count - m * m, 0.f);   // biased, as used for normalisation
  const float r = rsqrtf(var + eps);
  const float a = (gamma ? gamma[c] : 1.f) * r;
  mean[c] = m;
  rstd[c] = r;
  scale[c] = a;
  shift[c] = (beta ? beta[c] : 0.f) - m * a;
  if (run_mean != nullptr) {
    run_mean[c] = (1.f - momentum) * run_mean[c] + momentum * m;
    const float unbiased = count > 1.f ? var * count / (count - 1.f) : var;
    run_var[c] = (1.f - momentum) * run_var[c] + momentum * unbiased;
  }
}

// MODE 0: y = act(x * scale + shift)
// MODE 1: dx = scale * (g - sum_g / n - xhat * sum_gx / n),  g = dy * relu-mask
// Launched with a thread count that is a multiple of C / 8, so a thread keeps ONE channel group for its
// whole grid-stride loop and holds the per-channel constants in registers.
template <typename T, int MODE>
__global__ void __launch_bounds__(256)
bn_apply_kernel(const T* __restrict__ x, const T* __restrict__ dy, const float* __restrict__ mean,
                const float* __restrict__ rstd, const float* __restrict__ scale, const float* __restrict__ shift,
                const float* __restrict__ sums, T* __restrict__ out, long rows, int C, float inv_count, int relu) {
  if (MODE == 1 && inv_count <= 0.f) inv_count = 1.f / __ldg(sums + 2 * C);   // all-reduced sample count behind the sums (SyncBN)
  const int CG = C / 8;
  const long tid = blockIdx.x * (long)blockDim.x + threadIdx.x;
  const long nthr = (long)gridDim.x * blockDim.x;       // multiple of CG
  const int cg = (int)(tid % CG);
  float sc[8], sh[8], mu[8], rs[8], s1[8], s2[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = cg * 8 + k;
    sc[k] = __ldg(scale + c);
    sh[k] = __ldg(shift + c);
    if (MODE == 1) {
      mu[k] = __ldg(mean + c);
      rs[k] = __ldg(rstd + c);
      s1[k] = __ldg(sums + c) * inv_count;
      s2[k] = __ldg(sums + C + c) * inv_count;
    }
  }
  for (long r = tid / CG; r < rows; r += nthr / CG) {
    float v[8], o[8];
    bn_load8<T>(x + r * C + cg * 8, v);
    if (MODE == 0) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float t = fmaf(v[k], sc[k], sh[k]);
        o[k] = relu ? fmaxf(t, 0.f) : t;
      }
    } else {
      float g[8];
      bn_load8<T>(dy + r * C + cg * 8, g);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float gg = (relu && fmaf(v[k], sc[k], sh[k]) <= 0.f) ? 0.f : g[k];
        const float xh = (v[k] - mu[k]) * rs[k];
        o[k] = sc[k] * (gg - s1[k] - xh * s2[k]);
      }
    }
    bn_store8(out + r * C + cg * 8, o);
  }
}

static long bn_apply_blocks(long rows, int C) {
  const int CG = C / 8;
  // threads = multiple of lcm(256, CG); ~16 CTAs per SM at most
  long lcm = 256;
  while (lcm % CG != 0) lcm += 256;
  long want = (rows * CG + 255) / 256;
  const long cap = (long)kNumSMs * 16;
  if (want > cap) want = cap;
  const long unit = lcm / 256;
  long blocks = (want + unit - 1) / unit * unit;
  return blocks < unit ? unit : blocks;
}

template <typename T, int MODE>
static int bn_reduce_launch(const void* x, const void* dy, const float* mean, const float* rstd, const float* scale,
                            const float* shift, float* out, long rows, int C, int relu, cudaStream_t st) {
  const int gx = (C / 8 + BN_CG - 1) / BN_CG;
  long strip = rows * gx / ((long)kNumSMs * 8);
  if (strip < 64) strip = 64;
  if (strip > 2048) strip = 2048;
  const long gy = (rows + strip - 1) / strip;
  RF_REQUIRE(gy <= 65535, "batch-norm reduction: too many row strips");
  dim3 grid((unsigned)gx, (unsigned)gy);
  bn_reduce_kernel<T, MODE><<<grid, BN_CG * BN_RL, 0, st>>>((const T*)x, (const T*)dy, mean, rstd, scale, shift, out, rows,
                                                            C, strip, relu);
  RF_CHECK_LAUNCH("bn_reduce_kernel");
  return RF_OK;
}

static int bn_check(const void* x, long rows, int C, int dtype, const char* name) {
  RF_REQUIRE(x != nullptr && rows > 0 && C > 0, "%s: bad argument", name);
  RF_REQUIRE(C % 8 == 0, "%s: C=%d must be a multiple of 8", name, C);
  RF_REQUIRE(((uintptr_t)x & 15) == 0, "%s: tensors must be 16-byte aligned", name);
  RF_REQUIRE(dtype == 0 || dtype == 1, "%s: dtype must be 0 (f32) or 1 (bf16)", name);
  return RF_OK;
}

}  // namespace rf

using namespace rf;

extern "C" int rf_bn_stats(const void* x, float* sums, int64_t rows, int C, int dtype, void* stream) {
  int rc = bn_check(x, rows, C, dtype, "rf_bn_stats");
  if (rc != RF_OK) return rc;
  RF_REQUIRE(sums != nullptr, "rf_bn_stats: null output");
  cudaStream_t st = (cudaStream_t)stream;
  RF_CUDA(cudaMemsetAsync(sums, 0, sizeof(float) * 2 * C, st));
  if (dtype == 1)
    return bn_reduce_launch<__nv_bfloat16, 0>(x, nullptr, nullptr, nullptr, nullptr, nullptr, sums, rows, C, 0, st);
  return bn_reduce_launch<float, 0>(x, nullptr, nullptr, nullptr, nullptr, nullptr, sums, rows, C, 0, st);
}

extern "C" int rf_bn_finalize(const float* sums, const float* gamma, const float* beta, float* mean, float* rstd,
                              float* scale, float* shift, float* running_mean, float* running_var, int C,
                              double count, float eps, float momentum, void* stream) {
  RF_REQUIRE(sums && mean && rstd && scale && shift && C > 0 && (count >= 1.0 || count == 0.0), "rf_bn_finalize: bad argument");
  RF_REQUIRE((running_mean == nullptr) == (running_var == nullptr), "rf_bn_finalize: running statistics go together");
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(sums, gamma, beta, mean, rstd, scale, shift,
                                                                         running_mean, running_var, C, (float)count, eps,
                                                                         momentum);
  RF_CHECK_LAUNCH("bn_finalize_kernel");
  return RF_OK;
}

extern "C" int rf_bn_apply(const void* x, const float* scale, const float* shift, void* y, int64_t rows, int C,
                           int relu, int dtype, void* stream) {
  int rc = bn_check(x, rows, C, dtype, "rf_bn_apply");
  if (rc != RF_OK) return rc;
  RF_REQUIRE(scale && shift && y, "rf_bn_apply: null pointer");
  const long blocks = bn_apply_blocks(rows, C);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == 1)
    bn_apply_kernel<__nv_bfloat16, 0><<<(int)blocks, 256, 0, st>>>((const __nv_bfloat16*)x, nullptr, nullptr, nullptr,
                                                                   scale, shift, nullptr, (__nv_bfloat16*)y, rows, C,
                                                                   0.f, relu);
  else
    bn_apply_kernel<float, 0><<<(int)blocks, 256, 0, st>>>((const float*)x, nullptr, nullptr, nullptr, scale, shift,
                                                           nullptr, (float*)y, rows, C, 0.f, relu);
  RF_CHECK_LAUNCH("bn_apply_kernel");
  return RF_OK;
}

extern "C" int rf_bn_bwd_reduce(const void* x, const void* grad_y, const float* mean, const float* rstd,
                                const float* scale, const float* shift, float* sums, int64_t rows, int C, int relu,
                                int dtype, void* stream) {
  int rc = bn_check(x, rows, C, dtype, "rf_bn_bwd_reduce");
  if (rc != RF_OK) return rc;
  RF_REQUIRE(grad_y && mean && rstd && scale && shift && sums, "rf_bn_bwd_reduce: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  RF_CUDA(cudaMemsetAsync(sums, 0, sizeof(float) * 2 * C, st));
  if (dtype == 1) return bn_reduce_launch<__nv_bfloat16, 1>(x, grad_y, mean, rstd, scale, shift, sums, rows, C, relu, st);
  return bn_reduce_launch<float, 1>(x, grad_y, mean, rstd, scale, shift, sums, rows, C, relu, st);
}

extern "C" int rf_bn_bwd_apply(const void* x, const void* grad_y, const float* mean, const float* rstd,
                               const float* scale, const float* shift, const float* sums, void* grad_x, int64_t rows,
                               int C, double count, int relu, int dtype, void* stream) {
  int rc = bn_check(x, rows, C, dtype, "rf_bn_bwd_apply");
  if (rc != RF_OK) return rc;
  RF_REQUIRE(grad_y && mean && rstd && scale && shift && sums && grad_x && (count >= 1.0 || count == 0.0), "rf_bn_bwd_apply: bad argument");
  const long blocks = bn_apply_blocks(rows, C);
  cudaStream_t st = (cudaStream_t)stream;
  const float inv = count > 0.0 ? (float)(1.0 / count) : 0.f;   // 0: the kernel reads the all-reduced count from sums[2C]
  if (dtype == 1)
    bn_apply_kernel<__nv_bfloat16, 1><<<(int)blocks, 256, 0, st>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)grad_y,
                                                                   mean, rstd, scale, shift, sums, (__nv_bfloat16*)grad_x,
                                                                   rows, C, inv, relu);
  else
    bn_apply_kernel<float, 1><<<(int)blocks, 256, 0, st>>>((const float*)x, (const float*)grad_y, mean, rstd, scale, shift,
                                                           sums, (float*)grad_x, rows, C, inv, relu);
  RF_CHECK_LAUNCH("bn_apply_kernel");
  return RF_OK;
}
