// Residual-add + LayerNorm over the channel dimension of token tensors [rows, C] for sm_100a.
//
// Restates the pre-LN residual structure of the MiT block
//   x = x + drop_path(attn(norm1(x)));  x = x + drop_path(mlp(norm2(x)))
// (/root/reference/models/backbones/mix_transformer.py:203-207, LayerNorm eps 1e-6 :304; drop-path
// per-sample scaling /root/reference/models/modules.py:587-596) and the plain LayerNorms of
// OverlapPatchEmbed (:234,240, eps 1e-5), the SR branch (:135,148) and the stage norms (:378-426):
//   xn  = x + scale[b] * branch          (optional; scale = drop-path mask / keep-prob, or 1)
//   y   = (xn - mean) * rstd * gamma + beta
// The reference runs add, (cast,) LayerNorm, (cast) as separate passes; here the new residual stream
// (fp32) and the normalised activations (bf16 for the following tensor-core GEMM, or fp32) leave
// one kernel.  HBM-bound: one warp per row, the row lives in registers, two-pass mean / variance,
// 8- or 16-byte vector accesses.
// Backward: d(xn) = d(xn)_downstream + LN'(dy);  d(branch) = scale[b] * d(xn);  gamma / beta
// gradients are accumulated per warp in registers over its rows, combined in shared memory and
// added to global fp32 buffers with one atomic per channel per CTA.
#include <cuda_bf16.h>

#include <stdlib.h>

#include "rf_common.cuh"

namespace rf {

constexpr int LN_MAXV = 16;  // max elements per lane -> C <= 512

template <typename T>
__device__ __forceinline__ float ld1(const T* p);
template <>
__device__ __forceinline__ float ld1<float>(const float* p) { return __ldg(p); }
template <>
__device__ __forceinline__ float ld1<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void st1(float* p, float v) { *p = v; }
__device__ __forceinline__ void st1(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

// lane owns elements {(i*32 + lane)*2, +1} for i < NP (pairs): coalesced 8-byte (fp32) / 4-byte (bf16) accesses
template <typename T>
__device__ __forceinline__ float2 ld2(const T* p);
template <>
__device__ __forceinline__ float2 ld2<float>(const float* p) { return __ldg(reinterpret_cast<const float2*>(p)); }
template <>
__device__ __forceinline__ float2 ld2<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p));
}
__device__ __forceinline__ void st2(float* p, float2 v) { *reinterpret_cast<float2*>(p) = v; }
__device__ __forceinline__ void st2(__nv_bfloat16* p, float2 v) {
  *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(v.x, v.y);
}

// NP = C / 64 pairs per lane (C % 64 == 0), or generic scalar path when NP == 0 (C % 32 == 0, C <= 512).
template <typename TX, typename TB, typename TY, int NP>
__global__ void __launch_bounds__(256)
add_layernorm_fwd_kernel(const TX* __restrict__ x, const TB* __restrict__ branch, const float* __restrict__ scale,
                         const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ xn_out,
                         TY* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out, long rows,
                         int C, long rows_per_sample, float eps) {
  const int lane = threadIdx.x & 31;
  const long warp = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5;
  const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
  constexpr int NV = NP > 0 ? 2 * NP : LN_MAXV;
  const int nscalar = C / 32;  // generic path
  for (long r = warp; r < rows; r += nwarps) {
    float v[NV];
    const TX* xr = x + r * C;
    const float s = (branch != nullptr && scale != nullptr) ? __ldg(scale + r / rows_per_sample) : 1.f;
    float sum = 0.f;
    if (NP > 0) {
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        const int c = (i * 32 + lane) * 2;
        float2 a = ld2<TX>(xr + c);
        if (branch != nullptr) {
          const float2 bb = ld2<TB>(branch + r * C + c);
          a.x = fmaf(s, bb.x, a.x);
          a.y = fmaf(s, bb.y, a.y);
        }
        v[2 * i] = a.x;
        v[2 * i + 1] = a.y;
        sum += a.x + a.y;
      }
    } else {
#pragma unroll
      for (int i = 0; i < LN_MAXV; ++i)
        if (i < nscalar) {
          const int c = i * 32 + lane;
          float a = ld1<TX>(xr + c);
          if (branch != nullptr) a = fmaf(s, ld1<TB>(branch + r * C + c), a);
          v[i] = a;
          sum += a;
        }
    }
    const float mean = warp_sum(sum) / (float)C;
    float sq = 0.f;
    if (NP > 0) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const float d = v[i] - mean;
        sq = fmaf(d, d, sq);
      }
    } else {
#pragma unroll
      for (int i = 0; i < LN_MAXV; ++i)
        if (i < nscalar) {
          const float d = v[i] - mean;
          sq = fmaf(d, d, sq);
        }
    }
    const float rstd = rsqrtf(warp_sum(sq) / (float)C + eps);
    if (lane == 0) {
      if (mean_out) mean_out[r] = mean;
      if (rstd_out) rstd_out[r] = rstd;
    }
    if (NP > 0) {
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        const int c = (i * 32 + lane) * 2;
        if (xn_out != nullptr) st2(xn_out + r * C + c, make_float2(v[2 * i], v[2 * i + 1]));
        const float2 g = __ldg(reinterpret_cast<const float2*>(gamma + c));
        const float2 b = __ldg(reinterpret_cast<const float2*>(beta + c));
        st2(y + r * C + c, make_float2(fmaf((v[2 * i] - mean) * rstd, g.x, b.x),
                                       fmaf((v[2 * i + 1] - mean) * rstd, g.y, b.y)));
      }
    } else {
#pragma unroll
      for (int i = 0; i < LN_MAXV; ++i)
        if (i < nscalar) {
          const int c = i * 32 + lane;
          if (xn_out != nullptr) xn_out[r * C + c] = v[i];
          st1(y + r * C + c, fmaf((v[i] - mean) * rstd, __ldg(gamma + c), __ldg(beta + c)));
        }
    }
  }
}

// dxn = (dxn_in) + rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma,  xhat = (xn - mean) * rstd
// dbranch = scale[b] * dxn (optional).  dgamma += dy * xhat, dbeta += dy.
template <typename TXN, typename TY, typename TB, int NP>
__global__ void __launch_bounds__(256)
add_layernorm_bwd_kernel(const TXN* __restrict__ xn, const TY* __restrict__ dy, const float* __restrict__ dxn_in,
                         const float* __restrict__ mean, const float* __restrict__ rstd,
                         const float* __restrict__ gamma, const float* __restrict__ scale, float* __restrict__ dxn,
                         TB* __restrict__ dbranch, float* __restrict__ dgamma, float* __restrict__ dbeta, long rows,
                         int C, long rows_per_sample) {
  extern __shared__ float red[];  // [2][C]
  const int lane = threadIdx.x & 31;
  const long warp = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5;
  const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
  constexpr int NV = NP > 0 ? 2 * NP : LN_MAXV;
  const int nscalar = C / 32;
  float ag[NV], ab[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) ag[i] = ab[i] = 0.f;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
  for (long r = warp; r < rows; r += nwarps) {
    const float mu = __ldg(mean + r), rs = __ldg(rstd + r);
    float xh[NV], g[NV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (NP == 0 && i >= nscalar) break;
      int c;
      float xv, dv;
      if (NP > 0) {
        c = ((i >> 1) * 32 + lane) * 2 + (i & 1);
        if ((i & 1) == 0) {
          const float2 xx = ld2<TXN>(xn + r * C + c);
          const float2 dd = ld2<TY>(dy + r * C + c);
          xh[i] = (xx.x - mu) * rs;
          xh[i + 1] = (xx.y - mu) * rs;
          g[i] = dd.x;
          g[i + 1] = dd.y;
        }
        xv = xh[i];
        dv = g[i];
      } else {
        c = i * 32 + lane;
        xv = (ld1<TXN>(xn + r * C + c) - mu) * rs;
        dv = ld1<TY>(dy + r * C + c);
        xh[i] = xv;
      }
      ag[i] = fmaf(dv, xv, ag[i]);
      ab[i] += dv;
      const float gg = dv * __ldg(gamma + c);
      g[i] = gg;
      s1 += gg;
      s2 = fmaf(gg, xv, s2);
    }
    s1 = warp_sum(s1) / (float)C;
    s2 = warp_sum(s2) / (float)C;
    const float sc = (dbranch != nullptr && scale != nullptr) ? __ldg(scale + r / rows_per_sample) : 1.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (NP == 0 && i >= nscalar) break;
      g[i] = rs * (g[i] - s1 - xh[i] * s2);
    }
    if (NP > 0) {
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        const int c = (i * 32 + lane) * 2;
        float2 d = make_float2(g[2 * i], g[2 * i + 1]);
        if (dxn_in != nullptr) {
          const float2 e = __ldg(reinterpret_cast<const float2*>(dxn_in + r * C + c));
          d.x += e.x;
          d.y += e.y;
        }
        st2(dxn + r * C + c, d);
        if (dbranch != nullptr) st2(dbranch + r * C + c, make_float2(sc * d.x, sc * d.y));
      }
    } else {
#pragma unroll
      for (int i = 0; i < LN_MAXV; ++i)
        if (i < nscalar) {
          const int c = i * 32 + lane;
          float d = g[i];
          if (dxn_in != nullptr) d += __ldg(dxn_in + r * C + c);
          dxn[r * C + c] = d;
          if (dbranch != nullptr) st1(dbranch + r * C + c, sc * d);
        }
    }
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    if (NP == 0 && i >= nscalar) break;
    const int c = NP > 0 ? ((i >> 1) * 32 + lane) * 2 + (i & 1) : i * 32 + lane;
    atomicAdd(&red[c], ag[i]);
    atomicAdd(&red[C + c], ab[i]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    atomicAdd(dgamma + i, red[i]);
    atomicAdd(dbeta + i, red[C + i]);
  }
}

static int ln_grid(long rows) {
  long blocks = (rows + 7) / 8;  // 8 warps per CTA, one row per warp per iteration
  const long cap = (long)kNumSMs * 8;
  return (int)(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
}

// The backward ends with 2 C global atomics per CTA (dgamma / dbeta).  One CTA per 8 rows means 1 024-way contention per
// address at MiT stage 3 (8 192 rows, the shape of 160 of the 208 calls of a step), where the launch is short enough
// for that to show: 10.8 -> 6.7 us with 3 CTAs per SM looping over their rows (tools/bench_layernorm.py).  Long launches
// (stage 1: 131 072 rows) keep the full grid, they are bandwidth-bound and lose with fewer warps in flight.
// RF_LN_BWD_CTAS_PER_SM overrides the short-launch value.
static int ln_bwd_grid(long rows) {
  static const int per_sm = [] {
    const char* e = getenv("RF_LN_BWD_CTAS_PER_SM");
    const int v = e ? atoi(e) : 3;
    return v < 1 ? 1 : (v > 8 ? 8 : v);
  }();
  long blocks = (rows + 7) / 8;
  const long cap = (long)kNumSMs * (rows <= 16384 ? per_sm : 8);
  return (int)(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
}

template <typename TX, typename TB, typename TY>
static int ln_fwd_dispatch(const void* x, const void* branch, const float* scale, const float* gamma,
                           const float* beta, float* xn, void* y, float* mean, float* rstd, long rows, int C,
                           long rps, float eps, cudaStream_t st) {
  const int grid = ln_grid(rows);
#define RF_LN_FWD(NP)                                                                                         \
  add_layernorm_fwd_kernel<TX, TB, TY, NP><<<grid, 256, 0, st>>>((const TX*)x, (const TB*)branch, scale, gamma, \
                                                                  beta, xn, (TY*)y, mean, rstd, rows, C, rps, eps)
  if (C % 64 == 0 && C / 64 <= 8) {
    switch (C / 64) {
      case 1: RF_LN_FWD(1); break;
      case 2: RF_LN_FWD(2); break;
      case 4: RF_LN_FWD(4); break;
      case 5: RF_LN_FWD(5); break;
      case 8: RF_LN_FWD(8); break;
      default: RF_LN_FWD(0); break;
    }
  } else {
    RF_LN_FWD(0);
  }
#undef RF_LN_FWD
  RF_CHECK_LAUNCH("add_layernorm_fwd_kernel");
  return RF_OK;
}

template <typename TXN, typename TY, typename TB>
static int ln_bwd_dispatch(const void* xn, const void* dy, const float* dxn_in, const float* mean, const float* rstd,
                           const float* gamma, const float* scale, float* dxn, void* dbranch, float* dgamma,
                           float* dbeta, long rows, int C, long rps, cudaStream_t st) {
  const int grid = ln_bwd_grid(rows);
  const size_t smem = sizeof(float) * 2 * C;
#define RF_LN_BWD(NP)                                                                                            \
  add_layernorm_bwd_kernel<TXN, TY, TB, NP><<<grid, 256, smem, st>>>((const TXN*)xn, (const TY*)dy, dxn_in, mean, \
                                                                      rstd, gamma, scale, dxn, (TB*)dbranch,     \
                                                                      dgamma, dbeta, rows, C, rps)
  if (C % 64 == 0 && C / 64 <= 8) {
    switch (C / 64) {
      case 1: RF_LN_BWD(1); break;
      case 2: RF_LN_BWD(2); break;
      case 4: RF_LN_BWD(4); break;
      case 5: RF_LN_BWD(5); break;
      case 8: RF_LN_BWD(8); break;
      default: RF_LN_BWD(0); break;
    }
  } else {
    RF_LN_BWD(0);
  }
#undef RF_LN_BWD
  RF_CHECK_LAUNCH("add_layernorm_bwd_kernel");
  return RF_OK;
}

}  // namespace rf

using namespace rf;

extern "C" int rf_add_layernorm_fwd(const void* x, const void* branch, const float* scale, const float* gamma,
                                    const float* beta, float* xn_out, void* y, float* mean, float* rstd,
                                    int64_t rows, int C, int64_t rows_per_sample, float eps, int x_dtype,
                                    int branch_dtype, int y_dtype, void* stream) {
  RF_REQUIRE(x && gamma && beta && y, "rf_add_layernorm_fwd: null pointer");
  RF_REQUIRE(rows > 0 && C > 0 && C % 32 == 0 && C <= 32 * LN_MAXV, "rf_add_layernorm_fwd: C=%d must be a multiple of 32, <= %d",
             C, 32 * LN_MAXV);
  RF_REQUIRE(rows_per_sample > 0, "rf_add_layernorm_fwd: rows_per_sample must be positive");
  RF_REQUIRE(branch != nullptr || x_dtype == 0 || xn_out == nullptr, "rf_add_layernorm_fwd: xn_out without a branch needs f32 x");
  cudaStream_t st = (cudaStream_t)stream;
  const int key = (x_dtype << 2) | ((branch ? branch_dtype : 0) << 1) | y_dtype;
  switch (key) {
    case 0: return ln_fwd_dispatch<float, float, float>(x, branch, scale, gamma, beta, xn_out, y, mean, rstd, rows, C, rows_per_sample, eps, st);
    case 1: return ln_fwd_dispatch<float, float, __nv_bfloat16>(x, branch, scale, gamma, beta, xn_out, y, mean, rstd, rows, C, rows_per_sample, eps, st);
    case 2: return ln_fwd_dispatch<float, __nv_bfloat16, float>(x, branch, scale, gamma, beta, xn_out, y, mean, rstd, rows, C, rows_per_sample, eps, st);
    case 3: return ln_fwd_dispatch<float, __nv_bfloat16, __nv_bfloat16>(x, branch, scale, gamma, beta, xn_out, y, mean, rstd, rows, C, rows_per_sample, eps, st);
    case 4: return ln_fwd_dispatch<__nv_bfloat16, float, float>(x, branch, scale, gamma, beta, xn_out, y, mean, rstd, rows, C, rows_per_sample, eps, st);
    case 5: return ln_fwd_dispatch<__nv_bfloat16, float, __nv_bfloat16>(x, branch, scale, gamma, beta, xn_out, y, mean, rstd, rows, C, rows_per_sample, eps, st);
    case 6: return ln_fwd_dispatch<__nv_bfloat16, __nv_bfloat16, float>(x, branch, scale, gamma, beta, xn_out, y, mean, rstd, rows, C, rows_per_sample, eps, st);
    default: return ln_fwd_dispatch<__nv_bfloat16, __nv_bfloat16, __nv_bfloat16>(x, branch, scale, gamma, beta, xn_out, y, mean, rstd, rows, C, rows_per_sample, eps, st);
  }
}

extern "C" int rf_add_layernorm_bwd(const void* xn, const void* dy, const float* dxn_in, const float* mean,
                                    const float* rstd, const float* gamma, const float* scale, float* dxn,
                                    void* dbranch, float* dgamma, float* dbeta, int64_t rows, int C,
                                    int64_t rows_per_sample, int xn_dtype, int dy_dtype, int branch_dtype,
                                    int accumulate, void* stream) {
  RF_REQUIRE(xn && dy && mean && rstd && gamma && dxn && dgamma && dbeta, "rf_add_layernorm_bwd: null pointer");
  RF_REQUIRE(rows > 0 && C > 0 && C % 32 == 0 && C <= 32 * LN_MAXV, "rf_add_layernorm_bwd: C=%d must be a multiple of 32, <= %d",
             C, 32 * LN_MAXV);
  RF_REQUIRE(rows_per_sample > 0, "rf_add_layernorm_bwd: rows_per_sample must be positive");
  cudaStream_t st = (cudaStream_t)stream;
  if (!accumulate) {
    RF_CUDA(cudaMemsetAsync(dgamma, 0, sizeof(float) * C, st));
    RF_CUDA(cudaMemsetAsync(dbeta, 0, sizeof(float) * C, st));
  }
  const int key = (xn_dtype << 2) | (dy_dtype << 1) | (dbranch ? branch_dtype : 0);
  switch (key) {
    case 0: return ln_bwd_dispatch<float, float, float>(xn, dy, dxn_in, mean, rstd, gamma, scale, dxn, dbranch, dgamma, dbeta, rows, C, rows_per_sample, st);
    case 1: return ln_bwd_dispatch<float, float, __nv_bfloat16>(xn, dy, dxn_in, mean, rstd, gamma, scale, dxn, dbranch, dgamma, dbeta, rows, C, rows_per_sample, st);
    case 2: return ln_bwd_dispatch<float, __nv_bfloat16, float>(xn, dy, dxn_in, mean, rstd, gamma, scale, dxn, dbranch, dgamma, dbeta, rows, C, rows_per_sample, st);
    case 3: return ln_bwd_dispatch<float, __nv_bfloat16, __nv_bfloat16>(xn, dy, dxn_in, mean, rstd, gamma, scale, dxn, dbranch, dgamma, dbeta, rows, C, rows_per_sample, st);
    case 4: return ln_bwd_dispatch<__nv_bfloat16, float, float>(xn, dy, dxn_in, mean, rstd, gamma, scale, dxn, dbranch, dgamma, dbeta, rows, C, rows_per_sample, st);
    case 5: return ln_bwd_dispatch<__nv_bfloat16, float, __nv_bfloat16>(xn, dy, dxn_in, mean, rstd, gamma, scale, dxn, dbranch, dgamma, dbeta, rows, C, rows_per_sample, st);
    case 6: return ln_bwd_dispatch<__nv_bfloat16, __nv_bfloat16, float>(xn, dy, dxn_in, mean, rstd, gamma, scale, dxn, dbranch, dgamma, dbeta, rows, C, rows_per_sample, st);
    default: return ln_bwd_dispatch<__nv_bfloat16, __nv_bfloat16, __nv_bfloat16>(xn, dy, dxn_in, mean, rstd, gamma, scale, dxn, dbranch, dgamma, dbeta, rows, C, rows_per_sample, st);
  }
}
