// Local (windowed) correlation for sm_100a.
//
//   out[n,ph,pw,y,x] = sum_c sum_{i<kH,j<kW} in1[n,c,y',x'] * in2[n,c,y'+dy,x'+dx]
//
// Semantics follow the reference sampler
// (/root/reference/models/correlation_ops/correlation.cpp:14-42,80-129 and
//  correlation_cuda_kernel.cu:26-88): out-of-image terms contribute zero.
//
// Two forward paths:
//  * tiled kernel (the shape Refign uses: kernel 1, stride 1, pad 0, odd patch
//    <= 9, W % 4 == 0): one CTA owns an 8x32 tile of target pixels and ALL
//    displacements, streams channels through a 3-stage cp.async pipeline
//    (target tile + haloed source tile in shared memory, zero-filled outside
//    the image), keeps 4 px x P x 3 accumulators per thread in registers and
//    optionally fuses LocalFeatureCorrelationLayer's ReLU + L2-norm over the
//    P*P displacement channels (models/modules.py:271-273) because the CTA owns
//    every displacement of its pixels.  No NHWC permute copies, no output
//    zero-fill, launched on the caller's stream.
//  * generic kernel: any kernel/stride/pad/dilation, one thread per output.
//
// Backward: gather formulation (no atomics), one thread per input-gradient
// element, valid for every parameter combination.
#include "rf_common.cuh"

namespace rf {
// local_corr_tc.cu
int local_corr_tc_launch(const float* in1, const float* in2, float* out, float* norm_out, int B, int C, int H, int W, bool fuse,
                         cudaStream_t st);
bool local_corr_tc_ok(int C, int H, int W);
}  // namespace rf

namespace rf {

// ----------------------------------------------------------------------------
// tiled forward
// ----------------------------------------------------------------------------
constexpr int LC_TH = 8;        // tile rows
constexpr int LC_TW = 32;       // tile cols
constexpr int LC_CK = 8;        // channels per pipeline stage
constexpr int LC_STAGES = 3;
constexpr int LC_R4 = 4;        // halo columns each side (max radius, 16B aligned)
constexpr int LC_HW = LC_TW + 2 * LC_R4;  // 40 halo columns

template <int P>
struct LcCfg {
  static constexpr int R = (P - 1) / 2;
  static constexpr int NG = (P + 2) / 3;            // ph groups of 3
  static constexpr int THREADS = 64 * NG;
  static constexpr int HROWS = LC_TH + P - 1;
  static constexpr int IN1_FLOATS = LC_CK * LC_TH * LC_TW;
  static constexpr int IN2_FLOATS = LC_CK * HROWS * LC_HW;
  static constexpr int STAGE_FLOATS = IN1_FLOATS + IN2_FLOATS;
  static constexpr int SMEM_BYTES = LC_STAGES * STAGE_FLOATS * 4 + NG * LC_TH * LC_TW * 4;
};

template <int P>
__device__ __forceinline__ void lc_issue_stage(float* stage, const float* __restrict__ in1,
                                               const float* __restrict__ in2, int c0, int C,
                                               int H, int W, int y0, int x0) {
  using Cfg = LcCfg<P>;
  const long plane = (long)H * W;
  float* s1 = stage;
  float* s2 = stage + Cfg::IN1_FLOATS;
  // target tile: CK x TH rows of TW/4 16-byte chunks
  constexpr int N1 = LC_CK * LC_TH * (LC_TW / 4);
  for (int i = threadIdx.x; i < N1; i += Cfg::THREADS) {
    const int q = i % (LC_TW / 4);
    const int r = (i / (LC_TW / 4)) % LC_TH;
    const int c = i / (LC_TW / 4 * LC_TH);
    const int y = y0 + r, x = x0 + 4 * q, cc = c0 + c;
    const bool ok = (cc < C) && (y < H) && (x < W);
    const float* g = ok ? in1 + (long)cc * plane + (long)y * W + x : in1;
    cp_async16(s1 + (c * LC_TH + r) * LC_TW + 4 * q, g, ok ? 16 : 0);
  }
  constexpr int N2 = LC_CK * Cfg::HROWS * (LC_HW / 4);
  for (int i = threadIdx.x; i < N2; i += Cfg::THREADS) {
    const int q = i % (LC_HW / 4);
    const int r = (i / (LC_HW / 4)) % Cfg::HROWS;
    const int c = i / (LC_HW / 4 * Cfg::HROWS);
    const int y = y0 - Cfg::R + r, x = x0 - LC_R4 + 4 * q, cc = c0 + c;
    const bool ok = (cc < C) && (y >= 0) && (y < H) && (x >= 0) && (x < W);
    const float* g = ok ? in2 + (long)cc * plane + (long)y * W + x : in2;
    cp_async16(s2 + (c * Cfg::HROWS + r) * LC_HW + 4 * q, g, ok ? 16 : 0);
  }
}

// Accumulators of one thread: 3 displacement rows x P displacement columns x 4 pixels, stored in the pairing
// the packed FFMA2 inner loop needs: for pixel i the displacement columns are grouped as (P-1)/2 aligned pairs
// plus one single (the last column when i + OFF is even, the first when it is odd), because the source values
// bb[idx], bb[idx+1] of a pair must be an even-aligned register pair of the float4 shared-memory loads.
template <int P, int OFF>
struct LcAcc {
  static constexpr int NP = (P - 1) / 2;
  float2 pr[3][4][NP];
  float sg[3][4];
  __device__ __forceinline__ float& at(int p3, int pw, int i) {
    const int f = (i + OFF) & 1;
    if (f == 0) {
      if (pw == P - 1) return sg[p3][i];
      return (pw & 1) ? pr[p3][i][pw >> 1].y : pr[p3][i][pw >> 1].x;
    }
    if (pw == 0) return sg[p3][i];
    return ((pw - 1) & 1) ? pr[p3][i][(pw - 1) >> 1].y : pr[p3][i][(pw - 1) >> 1].x;
  }
};

template <int P, bool FUSE>
__global__ void __launch_bounds__(LcCfg<P>::THREADS, 2)
local_corr_tiled_kernel(const float* __restrict__ in1, const float* __restrict__ in2,
                        float* __restrict__ out, float* __restrict__ norm_out, int C, int H,
                        int W) {
  using Cfg = LcCfg<P>;
  extern __shared__ __align__(16) float smem[];
  float* ssq = smem + LC_STAGES * Cfg::STAGE_FLOATS;  // [NG][TH*TW]

  const int n = blockIdx.z;
  const int y0 = blockIdx.y * LC_TH, x0 = blockIdx.x * LC_TW;
  const long plane = (long)H * W;
  in1 += (long)n * C * plane;
  in2 += (long)n * C * plane;

  const int sx = threadIdx.x & 7;          // 4-pixel strip within the row
  const int r = (threadIdx.x >> 3) & 7;    // tile row
  const int g = threadIdx.x >> 6;          // ph group (warp-uniform)

  constexpr int OFF = LC_R4 - Cfg::R;
  LcAcc<P, OFF> acc;
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < P; ++b)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc.at(a, b, i) = 0.f;

  const int nchunks = (C + LC_CK - 1) / LC_CK;
#pragma unroll
  for (int s = 0; s < LC_STAGES - 1; ++s) {
    if (s < nchunks) lc_issue_stage<P>(smem + s * Cfg::STAGE_FLOATS, in1, in2, s * LC_CK, C, H, W, y0, x0);
    cp_async_commit();
  }

  for (int k = 0; k < nchunks; ++k) {
    cp_async_wait<LC_STAGES - 2>();
    __syncthreads();
    {  // prefetch chunk k + STAGES-1 into the slot consumed at iteration k-1
      const int kn = k + LC_STAGES - 1;
      if (kn < nchunks)
        lc_issue_stage<P>(smem + (kn % LC_STAGES) * Cfg::STAGE_FLOATS, in1, in2, kn * LC_CK, C, H, W, y0, x0);
      cp_async_commit();
    }
    const float* s1 = smem + (k % LC_STAGES) * Cfg::STAGE_FLOATS;
    const float* s2 = s1 + Cfg::IN1_FLOATS;
#pragma unroll
    for (int c = 0; c < LC_CK; ++c) {
      const float4 av = *reinterpret_cast<const float4*>(s1 + (c * LC_TH + r) * LC_TW + 4 * sx);
      const float a[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
      for (int p3 = 0; p3 < 3; ++p3) {
        const int ph = 3 * g + p3;
        if ((P % 3 == 0) || ph < P) {
          const float4* brow = reinterpret_cast<const float4*>(s2 + (c * Cfg::HROWS + r + ph) * LC_HW + 4 * sx);
          const float4 b0 = brow[0], b1 = brow[1], b2 = brow[2];
          const float bb[12] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w, b2.x, b2.y, b2.z, b2.w};
          // packed f32x2 FMAs over the displacement pairs of one pixel (see LcAcc): 5 issue slots per pixel
          // instead of 9, same products and the same per-accumulator summation order as a scalar loop.
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 a2 = make_float2(a[i], a[i]);
            constexpr int NPAIR = (P - 1) / 2;
            const int first = (i + OFF) & 1;
#pragma unroll
            for (int q = 0; q < NPAIR; ++q) {
              const int pw = first + 2 * q;
              acc.pr[p3][i][q] = __ffma2_rn(a2, make_float2(bb[i + pw + OFF], bb[i + pw + 1 + OFF]), acc.pr[p3][i][q]);
            }
            const int ps = first ? 0 : P - 1;
            acc.sg[p3][i] = fmaf(a[i], bb[i + ps + OFF], acc.sg[p3][i]);
          }
        }
      }
    }
  }
  cp_async_wait<0>();

  const int y = y0 + r, x = x0 + 4 * sx;
  float inv[4] = {1.f, 1.f, 1.f, 1.f};
  if (FUSE) {
    float ss[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int p3 = 0; p3 < 3; ++p3)
#pragma unroll
      for (int pw = 0; pw < P; ++pw)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float v = fmaxf(acc.at(p3, pw, i), 0.f);
          acc.at(p3, pw, i) = v;
          if ((P % 3 == 0) || 3 * g + p3 < P) ss[i] = fmaf(v, v, ss[i]);
        }
    __syncthreads();  // all stages consumed; ssq does not alias them but keep ordering simple
    *reinterpret_cast<float4*>(ssq + g * (LC_TH * LC_TW) + r * LC_TW + 4 * sx) = make_float4(ss[0], ss[1], ss[2], ss[3]);
    __syncthreads();
    float tot[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int gg = 0; gg < Cfg::NG; ++gg) {
      const float4 t = *reinterpret_cast<const float4*>(ssq + gg * (LC_TH * LC_TW) + r * LC_TW + 4 * sx);
      tot[0] += t.x; tot[1] += t.y; tot[2] += t.z; tot[3] += t.w;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) tot[i] = fmaxf(sqrtf(tot[i]), 1e-12f);
    if (norm_out != nullptr && g == 0 && y < H && x < W)
      *reinterpret_cast<float4*>(norm_out + (long)n * plane + (long)y * W + x) = make_float4(tot[0], tot[1], tot[2], tot[3]);
#pragma unroll
    for (int i = 0; i < 4; ++i) inv[i] = tot[i];
  }
  if (y < H && x < W) {
    float* o = out + (long)n * P * P * plane + (long)y * W + x;
#pragma unroll
    for (int p3 = 0; p3 < 3; ++p3) {
      const int ph = 3 * g + p3;
      if ((P % 3 == 0) || ph < P) {
#pragma unroll
        for (int pw = 0; pw < P; ++pw) {
          float4 v;
          if (FUSE) {
            v = make_float4(acc.at(p3, pw, 0) / inv[0], acc.at(p3, pw, 1) / inv[1], acc.at(p3, pw, 2) / inv[2],
                            acc.at(p3, pw, 3) / inv[3]);
          } else {
            v = make_float4(acc.at(p3, pw, 0), acc.at(p3, pw, 1), acc.at(p3, pw, 2), acc.at(p3, pw, 3));
          }
          *reinterpret_cast<float4*>(o + (long)(ph * P + pw) * plane) = v;  // stays L2-resident for the decoder that reads it next
        }
      }
    }
  }
}

template <int P>
static int launch_tiled(const float* in1, const float* in2, float* out, float* norm_out, int B, int C,
                        int H, int W, bool fuse, cudaStream_t st) {
  using Cfg = LcCfg<P>;
  dim3 grid((W + LC_TW - 1) / LC_TW, (H + LC_TH - 1) / LC_TH, B);
  if (fuse) {
    RF_CUDA(cudaFuncSetAttribute(local_corr_tiled_kernel<P, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    local_corr_tiled_kernel<P, true><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(in1, in2, out, norm_out, C, H, W);
  } else {
    RF_CUDA(cudaFuncSetAttribute(local_corr_tiled_kernel<P, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    local_corr_tiled_kernel<P, false><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(in1, in2, out, norm_out, C, H, W);
  }
  RF_CHECK_LAUNCH("local_corr_tiled_kernel");
  return RF_OK;
}

// ----------------------------------------------------------------------------
// generic forward: one thread per output element
// ----------------------------------------------------------------------------
struct LcGeom {
  int B, C, H, W, kH, kW, pH, pW, padH, padW, dilH, dilW, dpH, dpW, sH, sW, oH, oW;
};

__global__ void local_corr_generic_kernel(const float* __restrict__ in1, const float* __restrict__ in2,
                                          float* __restrict__ out, LcGeom g) {
  const long total = (long)g.B * g.pH * g.pW * g.oH * g.oW;
  const long plane = (long)g.H * g.W;
  const int radH = (g.pH - 1) / 2, radW = (g.pW - 1) / 2;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int x = idx % g.oW;
    const int y = (idx / g.oW) % g.oH;
    const int pw = (idx / ((long)g.oW * g.oH)) % g.pW;
    const int ph = (idx / ((long)g.oW * g.oH * g.pW)) % g.pH;
    const int n = idx / ((long)g.oW * g.oH * g.pW * g.pH);
    const int offy = (ph - radH) * g.dpH, offx = (pw - radW) * g.dpW;
    const int yb0 = y * g.sH - g.padH, xb0 = x * g.sW - g.padW;
    const float* a = in1 + (long)n * g.C * plane;
    const float* b = in2 + (long)n * g.C * plane;
    float acc = 0.f;
    for (int c = 0; c < g.C; ++c) {
      for (int i = 0; i < g.kH; ++i) {
        const int ya = yb0 + i * g.dilH, yb = ya + offy;
        if (ya < 0 || ya >= g.H || yb < 0 || yb >= g.H) continue;
        for (int j = 0; j < g.kW; ++j) {
          const int xa = xb0 + j * g.dilW, xb = xa + offx;
          if (xa < 0 || xa >= g.W || xb < 0 || xb >= g.W) continue;
          acc = fmaf(__ldg(a + c * plane + (long)ya * g.W + xa), __ldg(b + c * plane + (long)yb * g.W + xb), acc);
        }
      }
    }
    out[idx] = acc;
  }
}

// ----------------------------------------------------------------------------
// wide patches (9 < P <= 33, odd; kernel 1, stride 1, pad 0, dilation_patch 1): the sweep's stress points
// (max displacement 9 and 16).  One CTA = an 8 x 64 tile of target pixels x ONE displacement row ph; a thread owns
// two horizontally adjacent pixels and all P displacement columns (2 P accumulators): the source values
// s[x + pw] of pixel x and s[x + 1 + pw'] of its neighbour overlap, so one 8-byte shared-memory load feeds four
// FMAs (3.3 FMA per load at P = 19 against 0.95 for one pixel per thread).  Per 8-channel chunk the target tile and
// the 8 source rows (64 + 2 R columns, zero outside the image) are staged in shared memory.
// ----------------------------------------------------------------------------
constexpr int LW_CK = 8, LW_TH = 8, LW_TW = 64;

template <int PMAX>
__global__ void __launch_bounds__(256, 2)
local_corr_wide_kernel(const float* __restrict__ in1, const float* __restrict__ in2, float* __restrict__ out, int C,
                       int H, int W, int P) {
  constexpr int RMAX = (PMAX - 1) / 2;
  constexpr int SW = LW_TW + 2 * RMAX;     // staged source columns (even: rows stay 8-byte aligned)
  constexpr int NB2 = (PMAX + 2) / 2;      // float2 loads covering b[0 .. PMAX]
  __shared__ __align__(16) float s1[LW_CK][LW_TH][LW_TW];
  __shared__ __align__(16) float s2[LW_CK][LW_TH][SW];
  const int R = (P - 1) / 2;
  const int n = blockIdx.z / P, ph = blockIdx.z % P;
  const int y0 = blockIdx.y * LW_TH, x0 = blockIdx.x * LW_TW;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const long plane = (long)H * W;
  const float* a_n = in1 + (long)n * C * plane;
  const float* b_n = in2 + (long)n * C * plane;
  const int sw = LW_TW + 2 * R;            // source columns actually needed
  float acc0[PMAX], acc1[PMAX];
#pragma unroll
  for (int i = 0; i < PMAX; ++i) acc0[i] = acc1[i] = 0.f;

  for (int c0 = 0; c0 < C; c0 += LW_CK) {
    {  // staging: warp ty owns tile row ty of every channel of the chunk (no index divisions)
      const int yy1 = y0 + ty, yy2 = y0 + ty + ph - R;
      const bool y1ok = yy1 < H, y2ok = yy2 >= 0 && yy2 < H;
#pragma unroll
      for (int c = 0; c < LW_CK; ++c) {
        const bool cok = c0 + c < C;
        const float* pa = a_n + (long)(c0 + c) * plane + (long)yy1 * W;
        const float* pb = b_n + (long)(c0 + c) * plane + (long)yy2 * W;
#pragma unroll
        for (int x = tx; x < LW_TW; x += 32) {
          const int xx = x0 + x;
          s1[c][ty][x] = (cok && y1ok && xx < W) ? __ldg(pa + xx) : 0.f;
        }
        for (int x = tx; x < sw; x += 32) {
          const int xx = x0 + x - R;
          s2[c][ty][x] = (cok && y2ok && xx >= 0 && xx < W) ? __ldg(pb + xx) : 0.f;
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < LW_CK; ++c) {
      const float2 a = *reinterpret_cast<const float2*>(&s1[c][ty][2 * tx]);
      float b[2 * NB2];
#pragma unroll
      for (int k = 0; k < NB2; ++k)
        if (2 * k <= P) {   // b[pw] for pixel 0 and b[pw + 1] for pixel 1, pw < P
          const float2 v = *reinterpret_cast<const float2*>(&s2[c][ty][2 * tx + 2 * k]);
          b[2 * k] = v.x;
          b[2 * k + 1] = v.y;
        }
#pragma unroll
      for (int pw = 0; pw < PMAX; ++pw)
        if (pw < P) {
          acc0[pw] = fmaf(a.x, b[pw], acc0[pw]);
          acc1[pw] = fmaf(a.y, b[pw + 1], acc1[pw]);
        }
    }
    __syncthreads();
  }
  const int y = y0 + ty, x = x0 + 2 * tx;
  if (y < H && x < W) {
    float* o = out + (((long)n * P + ph) * P) * plane + (long)y * W + x;
    const bool two = x + 1 < W;
    const bool vec = two && ((W & 1) == 0) && ((reinterpret_cast<uintptr_t>(out) & 7) == 0);
#pragma unroll
    for (int pw = 0; pw < PMAX; ++pw)
      if (pw < P) {
        if (vec) {
          *reinterpret_cast<float2*>(o + (long)pw * plane) = make_float2(acc0[pw], acc1[pw]);
        } else {
          o[(long)pw * plane] = acc0[pw];
          if (two) o[(long)pw * plane + 1] = acc1[pw];
        }
      }
  }
}

// relu + l2norm over K channels of [B,K,HW] (generic companion of the fused epilogue)
// block = 32 pixels x 8 channel groups: with one thread per pixel a [2,361,128,128] volume had 32 k threads doing
// ~1 000 dependent strided accesses each (latency-bound, ~0.5 ms); here the channel loop is split 8 ways and the
// partial sums meet in shared memory
__global__ void __launch_bounds__(256)
relu_l2norm_kernel(float* __restrict__ x, float* __restrict__ norm_out, int K, long HW, long total) {
  __shared__ float part[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (long base = blockIdx.x * 32l; base < total; base += (long)gridDim.x * 32) {   // block-uniform trip count
    const long idx = base + tx;
    const bool ok = idx < total;
    const long b = ok ? idx / HW : 0, p = ok ? idx % HW : 0;
    float* px = x + b * K * HW + p;
    float ss = 0.f;
    if (ok)
      for (int k = ty; k < K; k += 8) {
        const float v = fmaxf(px[(long)k * HW], 0.f);
        ss = fmaf(v, v, ss);
      }
    part[ty][tx] = ss;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) tot += part[i][tx];
    const float nrm = fmaxf(sqrtf(tot), 1e-12f);
    if (ok) {
      if (norm_out && ty == 0) norm_out[idx] = nrm;
      for (int k = ty; k < K; k += 8) px[(long)k * HW] = fmaxf(px[(long)k * HW], 0.f) / nrm;
    }
    __syncthreads();
  }
}

// ----------------------------------------------------------------------------
// backward (gather form)
// ----------------------------------------------------------------------------
// grad_in1[n,c,ya,xa] = sum_{ph,pw,i,j} gout[n,ph,pw,y,x] * in2[n,c,ya+offy,xa+offx]
//   with y*sH - padH + i*dilH == ya  (likewise x)
// grad_in2[n,c,yb,xb] = sum_{ph,pw,i,j} gout[n,ph,pw,y,x] * in1[n,c,yb-offy,xb-offx]
template <bool FOR_IN2>
__global__ void local_corr_bwd_kernel(const float* __restrict__ other, const float* __restrict__ gout,
                                      float* __restrict__ gin, LcGeom g) {
  const long plane = (long)g.H * g.W;
  const long total = (long)g.B * g.C * plane;
  const long oplane = (long)g.oH * g.oW;
  const int radH = (g.pH - 1) / 2, radW = (g.pW - 1) / 2;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int xq = idx % g.W;
    const int yq = (idx / g.W) % g.H;
    const long nc = idx / plane;
    const int n = nc / g.C;
    const float* oth = other + nc * plane;
    const float* go = gout + (long)n * g.pH * g.pW * oplane;
    float acc = 0.f;
    for (int ph = 0; ph < g.pH; ++ph) {
      const int offy = (ph - radH) * g.dpH;
      const int ya = FOR_IN2 ? yq - offy : yq;      // coordinate in input1
      const int yo = FOR_IN2 ? ya : yq + offy;      // coordinate read from `other`
      if (ya < 0 || ya >= g.H || yo < 0 || yo >= g.H) continue;
      for (int pw = 0; pw < g.pW; ++pw) {
        const int offx = (pw - radW) * g.dpW;
        const int xa = FOR_IN2 ? xq - offx : xq;
        const int xo = FOR_IN2 ? xa : xq + offx;
        if (xa < 0 || xa >= g.W || xo < 0 || xo >= g.W) continue;
        const float ov = __ldg(oth + (long)yo * g.W + xo);
        const float* gp = go + (long)(ph * g.pW + pw) * oplane;
        for (int i = 0; i < g.kH; ++i) {
          const int ty = ya + g.padH - i * g.dilH;
          if (ty < 0 || ty % g.sH) continue;
          const int y = ty / g.sH;
          if (y >= g.oH) continue;
          for (int j = 0; j < g.kW; ++j) {
            const int tx = xa + g.padW - j * g.dilW;
            if (tx < 0 || tx % g.sW) continue;
            const int x = tx / g.sW;
            if (x >= g.oW) continue;
            acc = fmaf(__ldg(gp + (long)y * g.oW + x), ov, acc);
          }
        }
      }
    }
    gin[idx] = acc;
  }
}

// ----------------------------------------------------------------------------
// tiled backward (kernel 1, stride 1, pad 0, dilation_patch 1, odd P <= 9, W % 4 == 0)
// ----------------------------------------------------------------------------
// Both gradients are the same contraction
//     gin[n,c,y,x] = sum_{ph,pw} G[n,ph,pw,y,x] * S[n,c,y+ph-R,x+pw-R]            (S = 0 outside the image)
// grad_in1: G = grad_out,  S = in2.
// grad_in2: gin2[c,q] = sum_d gout[d, q-d] in1[c, q-d]; substituting d' = -d gives the same form with
//           G'[d'][q] = gout[-d'][q + d'] (zero outside) and S = in1 -- G' is built by local_corr_flip_shift_kernel
//           into the caller's scratch buffer (one streaming pass over the volume).
// CTA = 16 x 32 pixel tile, 128 threads (thread = 4 consecutive pixels of one row, all LC_CK channels of the
// current chunk in registers); the S halo tile streams through the same 3-stage cp.async pipeline as the forward;
// the thread's G values of one displacement row are read straight from L2 (coalesced 128-byte rows) and
// prefetched one displacement row ahead.
constexpr int LB_TH = 16;

template <int P>
struct LbCfg {
  static constexpr int R = (P - 1) / 2;
  static constexpr int THREADS = LB_TH * (LC_TW / 4);   // 128
  static constexpr int HROWS = LB_TH + P - 1;
  static constexpr int STAGE_FLOATS = LC_CK * HROWS * LC_HW;
  static constexpr int SMEM_BYTES = LC_STAGES * STAGE_FLOATS * 4;
};

template <int P>
__device__ __forceinline__ void lb_issue_stage(float* stage, const float* __restrict__ src, int c0, int C, int H,
                                               int W, int y0, int x0) {
  using Cfg = LbCfg<P>;
  const long plane = (long)H * W;
  constexpr int N2 = LC_CK * Cfg::HROWS * (LC_HW / 4);
  for (int i = threadIdx.x; i < N2; i += Cfg::THREADS) {
    const int q = i % (LC_HW / 4);
    const int r = (i / (LC_HW / 4)) % Cfg::HROWS;
    const int c = i / (LC_HW / 4 * Cfg::HROWS);
    const int y = y0 - Cfg::R + r, x = x0 - LC_R4 + 4 * q, cc = c0 + c;
    const bool ok = (cc < C) && (y >= 0) && (y < H) && (x >= 0) && (x < W);
    const float* g = ok ? src + (long)cc * plane + (long)y * W + x : src;
    cp_async16(stage + (c * Cfg::HROWS + r) * LC_HW + 4 * q, g, ok ? 16 : 0);
  }
}

template <int P>
__global__ void __launch_bounds__(LbCfg<P>::THREADS, 3)
local_corr_bwd_tiled_kernel(const float* __restrict__ G, const float* __restrict__ S, float* __restrict__ gin, int C,
                            int H, int W, int nsplit) {
  using Cfg = LbCfg<P>;
  extern __shared__ __align__(16) float smem[];
  // blockIdx.z = image * nsplit + channel split: every output channel is independent, so small maps fill the
  // machine by giving each CTA a range of channel chunks
  const int n = blockIdx.z / nsplit, split = blockIdx.z % nsplit;
  const int y0 = blockIdx.y * LB_TH, x0 = blockIdx.x * LC_TW;
  const long plane = (long)H * W;
  S += (long)n * C * plane;
  gin += (long)n * C * plane;
  G += (long)n * P * P * plane;
  const int sx = threadIdx.x & 7, r = threadIdx.x >> 3;
  const int y = y0 + r, x = x0 + 4 * sx;
  const bool live = y < H && x < W;
  const float4* gp = reinterpret_cast<const float4*>(G + (long)(live ? y : 0) * W + (live ? x : 0));
  const long gstride = plane / 4;   // float4 stride between displacement planes (W % 4 == 0)
  constexpr int OFF = LC_R4 - Cfg::R;

  const int allchunks = (C + LC_CK - 1) / LC_CK;
  const int per = (allchunks + nsplit - 1) / nsplit;
  const int kbeg = split * per;
  const int nchunks = min(per, allchunks - kbeg);   // chunks of this CTA (may be <= 0 for the last split)
#pragma unroll
  for (int s = 0; s < LC_STAGES - 1; ++s) {
    if (s < nchunks) lb_issue_stage<P>(smem + s * Cfg::STAGE_FLOATS, S, (kbeg + s) * LC_CK, C, H, W, y0, x0);
    cp_async_commit();
  }
  for (int kk = 0; kk < nchunks; ++kk) {
    const int k = kbeg + kk;
    cp_async_wait<LC_STAGES - 2>();
    __syncthreads();
    {
      const int kn = kk + LC_STAGES - 1;
      if (kn < nchunks)
        lb_issue_stage<P>(smem + (kn % LC_STAGES) * Cfg::STAGE_FLOATS, S, (kbeg + kn) * LC_CK, C, H, W, y0, x0);
      cp_async_commit();
    }
    const float* st = smem + (kk % LC_STAGES) * Cfg::STAGE_FLOATS;
    float acc[LC_CK][4];
#pragma unroll
    for (int c = 0; c < LC_CK; ++c)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[c][i] = 0.f;
    float4 gcur[P], gnext[P];
#pragma unroll
    for (int pw = 0; pw < P; ++pw) gcur[pw] = live ? __ldg(gp + (long)pw * gstride) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
    for (int ph = 0; ph < P; ++ph) {
      if (ph + 1 < P) {
#pragma unroll
        for (int pw = 0; pw < P; ++pw)
          gnext[pw] = live ? __ldg(gp + (long)((ph + 1) * P + pw) * gstride) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int c = 0; c < LC_CK; ++c) {
        const float4* brow = reinterpret_cast<const float4*>(st + (c * Cfg::HROWS + r + ph) * LC_HW + 4 * sx);
        const float4 b0 = brow[0], b1 = brow[1], b2 = brow[2];
        const float bb[12] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w, b2.x, b2.y, b2.z, b2.w};
#pragma unroll
        for (int pw = 0; pw < P; ++pw) {
          acc[c][0] = fmaf(gcur[pw].x, bb[0 + pw + OFF], acc[c][0]);
          acc[c][1] = fmaf(gcur[pw].y, bb[1 + pw + OFF], acc[c][1]);
          acc[c][2] = fmaf(gcur[pw].z, bb[2 + pw + OFF], acc[c][2]);
          acc[c][3] = fmaf(gcur[pw].w, bb[3 + pw + OFF], acc[c][3]);
        }
      }
#pragma unroll
      for (int pw = 0; pw < P; ++pw) gcur[pw] = gnext[pw];
    }
    if (live) {
#pragma unroll
      for (int c = 0; c < LC_CK; ++c) {
        const int cc = k * LC_CK + c;
        if (cc < C)
          *reinterpret_cast<float4*>(gin + (long)cc * plane + (long)y * W + x) =
              make_float4(acc[c][0], acc[c][1], acc[c][2], acc[c][3]);
      }
    }
  }
  cp_async_wait<0>();
}

// G'[n, ph', pw', y, x] = gout[n, P-1-ph', P-1-pw', y + ph' - R, x + pw' - R]  (0 outside the image)
__global__ void local_corr_flip_shift_kernel(const float* __restrict__ gout, float* __restrict__ gflip, int P, int H,
                                             int W, long total) {
  const int R = (P - 1) / 2;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int x = (int)(idx % W);
    long t = idx / W;
    const int y = (int)(t % H);
    t /= H;
    const int pw = (int)(t % P);
    t /= P;
    const int ph = (int)(t % P);
    const long n = t / P;
    const int ys = y + ph - R, xs = x + pw - R;
    float v = 0.f;
    if (ys >= 0 && ys < H && xs >= 0 && xs < W)
      v = __ldg(gout + ((n * P + (P - 1 - ph)) * P + (P - 1 - pw)) * (long)H * W + (long)ys * W + xs);
    gflip[idx] = v;
  }
}

template <int P>
static int launch_bwd_tiled(const float* G, const float* S, float* gin, int B, int C, int H, int W, cudaStream_t st) {
  using Cfg = LbCfg<P>;
  const long tiles = (long)((W + LC_TW - 1) / LC_TW) * ((H + LB_TH - 1) / LB_TH) * B;
  const int allchunks = (C + LC_CK - 1) / LC_CK;
  long nsplit = ((long)kNumSMs * 2 + tiles - 1) / tiles;     // >= 2 CTAs per SM when the channel count allows
  if (nsplit > allchunks) nsplit = allchunks;
  if (nsplit < 1) nsplit = 1;
  RF_REQUIRE((long)B * nsplit <= 65535, "rf_local_corr_bwd: batch too large");
  dim3 grid((W + LC_TW - 1) / LC_TW, (H + LB_TH - 1) / LB_TH, (unsigned)(B * nsplit));
  RF_CUDA(cudaFuncSetAttribute(local_corr_bwd_tiled_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  local_corr_bwd_tiled_kernel<P><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(G, S, gin, C, H, W, (int)nsplit);
  RF_CHECK_LAUNCH("local_corr_bwd_tiled_kernel");
  return RF_OK;
}

static int bwd_tiled_dispatch(int P, const float* G, const float* S, float* gin, int B, int C, int H, int W,
                              cudaStream_t st) {
  switch (P) {
    case 3: return launch_bwd_tiled<3>(G, S, gin, B, C, H, W, st);
    case 5: return launch_bwd_tiled<5>(G, S, gin, B, C, H, W, st);
    case 7: return launch_bwd_tiled<7>(G, S, gin, B, C, H, W, st);
    default: return launch_bwd_tiled<9>(G, S, gin, B, C, H, W, st);
  }
}

static bool bwd_tiled_ok(int H, int W, int kH, int kW, int pH, int pW, int padH, int padW, int dpH, int dpW, int sH,
                         int sW) {
  (void)H;
  return kH == 1 && kW == 1 && sH == 1 && sW == 1 && padH == 0 && padW == 0 && dpH == 1 && dpW == 1 && pH == pW &&
         (pH == 3 || pH == 5 || pH == 7 || pH == 9) && W % 4 == 0;
}

// backward of y = relu(c)/max(||relu(c)||,eps):  dc = relu'(c) * (gy - y * sum_k(gy*y)) / norm
__global__ void relu_l2norm_bwd_kernel(const float* __restrict__ y, const float* __restrict__ norm,
                                       const float* __restrict__ gy, float* __restrict__ gc, int K, long HW,
                                       long total) {
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const long b = idx / HW, p = idx % HW;
    const long base = b * K * HW + p;
    const float nrm = norm[idx];
    float dot = 0.f;
    for (int k = 0; k < K; ++k) dot = fmaf(gy[base + (long)k * HW], y[base + (long)k * HW], dot);
    // when the norm was clamped to eps the reference's normalize has d||x||/dx = 0 contribution
    const bool clamped = nrm <= 1e-12f;
    for (int k = 0; k < K; ++k) {
      const float yk = y[base + (long)k * HW];
      const float g0 = gy[base + (long)k * HW];
      const float v = clamped ? g0 / nrm : (g0 - yk * dot) / nrm;
      gc[base + (long)k * HW] = yk > 0.f ? v : 0.f;
    }
  }
}

static LcGeom make_geom(int B, int C, int H, int W, int kH, int kW, int pH, int pW, int padH, int padW, int dilH,
                        int dilW, int dpH, int dpW, int sH, int sW) {
  LcGeom g{B, C, H, W, kH, kW, pH, pW, padH, padW, dilH, dilW, dpH, dpW, sH, sW, 0, 0};
  g.oH = (H + 2 * padH - ((kH - 1) * dilH + 1)) / sH + 1;
  g.oW = (W + 2 * padW - ((kW - 1) * dilW + 1)) / sW + 1;
  return g;
}

static inline int grid_for(long total, int threads) {
  long b = (total + threads - 1) / threads;
  const long cap = (long)kNumSMs * 32;
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

}  // namespace rf

using namespace rf;

extern "C" int rf_local_corr_fwd(const float* in1, const float* in2, float* out, float* norm_out, int B, int C,
                                 int H, int W, int kH, int kW, int pH, int pW, int padH, int padW, int dilH,
                                 int dilW, int dpH, int dpW, int sH, int sW, int fuse_relu_l2norm, void* stream) {
  RF_REQUIRE(in1 && in2 && out, "rf_local_corr_fwd: null pointer");
  RF_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, "rf_local_corr_fwd: empty tensor (B=%d C=%d H=%d W=%d)", B, C, H, W);
  RF_REQUIRE(kH > 0 && kW > 0 && pH > 0 && pW > 0 && sH > 0 && sW > 0 && dilH > 0 && dilW > 0 && dpH > 0 && dpW > 0 && padH >= 0 && padW >= 0,
             "rf_local_corr_fwd: invalid geometry");
  LcGeom g = make_geom(B, C, H, W, kH, kW, pH, pW, padH, padW, dilH, dilW, dpH, dpW, sH, sW);
  RF_REQUIRE(g.oH > 0 && g.oW > 0, "rf_local_corr_fwd: empty output (%d x %d)", g.oH, g.oW);
  cudaStream_t st = (cudaStream_t)stream;
  const bool unit = kH == 1 && kW == 1 && sH == 1 && sW == 1 && padH == 0 && padW == 0 && dpH == 1 && dpW == 1;
  RF_REQUIRE(!fuse_relu_l2norm || unit, "rf_local_corr_fwd: fused relu+l2norm needs kernel 1, stride 1, pad 0, dilation_patch 1");
  const bool aligned = (W % 4 == 0) && (((uintptr_t)in1 | (uintptr_t)in2 | (uintptr_t)out | (uintptr_t)norm_out) % 16 == 0);
  if (unit && aligned && pH == 9 && pW == 9 && local_corr_tc_ok(C, H, W)) {
    // tensor-core banded GEMM (local_corr_tc.cu; bf16 hi/lo split, <= 2e-5 abs on unit-norm features).
    // RF_LOCAL_CORR_TC=0 keeps the exact-fp32 FFMA tiles (A/B switch, read per call).
    const char* e = getenv("RF_LOCAL_CORR_TC");
    if (!(e && e[0] == '0')) return local_corr_tc_launch(in1, in2, out, norm_out, B, C, H, W, fuse_relu_l2norm != 0, st);
  }
  if (unit && aligned && pH == pW && (pH == 3 || pH == 5 || pH == 7 || pH == 9)) {
    switch (pH) {
      case 3: return launch_tiled<3>(in1, in2, out, norm_out, B, C, H, W, fuse_relu_l2norm != 0, st);
      case 5: return launch_tiled<5>(in1, in2, out, norm_out, B, C, H, W, fuse_relu_l2norm != 0, st);
      case 7: return launch_tiled<7>(in1, in2, out, norm_out, B, C, H, W, fuse_relu_l2norm != 0, st);
      default: return launch_tiled<9>(in1, in2, out, norm_out, B, C, H, W, fuse_relu_l2norm != 0, st);
    }
  }
  if (unit && pH == pW && (pH & 1) && pH > 9 && pH <= 33 && (long)B * pH <= 65535) {   // wide odd patches: tiled, one ph per CTA
    dim3 grid((unsigned)ceil_div(W, LW_TW), (unsigned)ceil_div(H, LW_TH), (unsigned)(B * pH));
    if (pH <= 19)
      local_corr_wide_kernel<19><<<grid, 256, 0, st>>>(in1, in2, out, C, H, W, pH);
    else
      local_corr_wide_kernel<33><<<grid, 256, 0, st>>>(in1, in2, out, C, H, W, pH);
    RF_CHECK_LAUNCH("local_corr_wide_kernel");
  } else {
    const long total = (long)B * pH * pW * g.oH * g.oW;
    local_corr_generic_kernel<<<grid_for(total, 256), 256, 0, st>>>(in1, in2, out, g);
    RF_CHECK_LAUNCH("local_corr_generic_kernel");
  }
  if (fuse_relu_l2norm) {
    const long npix = (long)B * g.oH * g.oW;
    relu_l2norm_kernel<<<grid_for(npix, 32), 256, 0, st>>>(out, norm_out, pH * pW, (long)g.oH * g.oW, npix);
    RF_CHECK_LAUNCH("relu_l2norm_kernel");
  }
  return RF_OK;
}

extern "C" int64_t rf_local_corr_bwd_scratch_bytes(int B, int C, int H, int W, int kH, int kW, int pH, int pW, int padH,
                                                   int padW, int dilH, int dilW, int dpH, int dpW, int sH, int sW) {
  (void)C; (void)dilH; (void)dilW;
  // tiled path: the flipped / shifted copy of grad_out used for grad_in2; the generic gather path needs none
  if (bwd_tiled_ok(H, W, kH, kW, pH, pW, padH, padW, dpH, dpW, sH, sW))
    return (int64_t)sizeof(float) * B * pH * pW * H * W;
  return 0;
}

extern "C" int rf_local_corr_bwd(const float* in1, const float* in2, const float* grad_out, float* grad_in1,
                                 float* grad_in2, void* scratch, int B, int C, int H, int W, int kH, int kW, int pH,
                                 int pW, int padH, int padW, int dilH, int dilW, int dpH, int dpW, int sH, int sW,
                                 void* stream) {
  RF_REQUIRE(in1 && in2 && grad_out, "rf_local_corr_bwd: null pointer");
  RF_REQUIRE(grad_in1 || grad_in2, "rf_local_corr_bwd: no gradient requested");
  RF_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, "rf_local_corr_bwd: empty tensor");
  LcGeom g = make_geom(B, C, H, W, kH, kW, pH, pW, padH, padW, dilH, dilW, dpH, dpW, sH, sW);
  RF_REQUIRE(g.oH > 0 && g.oW > 0, "rf_local_corr_bwd: empty output");
  cudaStream_t st = (cudaStream_t)stream;
  const bool al = (((uintptr_t)in1 | (uintptr_t)in2 | (uintptr_t)grad_out | (uintptr_t)grad_in1 | (uintptr_t)grad_in2 |
                    (uintptr_t)scratch) % 16) == 0;
  if (al && bwd_tiled_ok(H, W, kH, kW, pH, pW, padH, padW, dpH, dpW, sH, sW) && (grad_in2 == nullptr || scratch != nullptr)) {
    int rc = RF_OK;
    if (grad_in1) rc = bwd_tiled_dispatch(pH, grad_out, in2, grad_in1, B, C, H, W, st);
    if (rc == RF_OK && grad_in2) {
      const long vol = (long)B * pH * pW * H * W;
      local_corr_flip_shift_kernel<<<grid_for(vol, 256), 256, 0, st>>>(grad_out, (float*)scratch, pH, H, W, vol);
      RF_CHECK_LAUNCH("local_corr_flip_shift_kernel");
      rc = bwd_tiled_dispatch(pH, (const float*)scratch, in1, grad_in2, B, C, H, W, st);
    }
    return rc;
  }
  const long total = (long)B * C * H * W;
  if (grad_in1) {
    local_corr_bwd_kernel<false><<<grid_for(total, 256), 256, 0, st>>>(in2, grad_out, grad_in1, g);
    RF_CHECK_LAUNCH("local_corr_bwd_kernel<in1>");
  }
  if (grad_in2) {
    local_corr_bwd_kernel<true><<<grid_for(total, 256), 256, 0, st>>>(in1, grad_out, grad_in2, g);
    RF_CHECK_LAUNCH("local_corr_bwd_kernel<in2>");
  }
  return RF_OK;
}

extern "C" int rf_relu_l2norm_bwd(const float* y, const float* norm, const float* grad_y, float* grad_c, int B, int K,
                                  int64_t HW, void* stream) {
  RF_REQUIRE(y && norm && grad_y && grad_c, "rf_relu_l2norm_bwd: null pointer");
  RF_REQUIRE(B > 0 && K > 0 && HW > 0, "rf_relu_l2norm_bwd: empty tensor");
  const long total = (long)B * HW;
  relu_l2norm_bwd_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(y, norm, grad_y, grad_c, K, HW, total);
  RF_CHECK_LAUNCH("relu_l2norm_bwd_kernel");
  return RF_OK;
}
