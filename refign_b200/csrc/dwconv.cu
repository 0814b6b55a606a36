// Depthwise 3x3 convolution on channels-last (NHWC == token [B, H*W, C]) tensors for sm_100a:
// forward (+ bias, + exact-erf GELU), input gradient, GELU-backward pre-pass and weight/bias gradient.
//
// Two reference call sites on the Refign hot path run through these kernels:
//   * Mix-FFN DWConv (+ nn.GELU)            /root/reference/models/backbones/mix_transformer.py:96-103,556-568
//     (dilation 1, bias, GELU fused; tokens stay [B,N,C] -- the reference transposes to NCHW and back per block);
//   * DAFormer ASPP depthwise branch         /root/reference/models/heads/daformer.py:26-35 via
//     DepthwiseSeparableConvModule           /root/reference/models/modules.py:29-36
//     (dilation 6/12/18, padding = dilation, no bias; BN + ReLU follow as separate ops).
//
// All four kernels are HBM/L2 streaming kernels: one thread owns 8 consecutive channels (one 16-byte
// bf16 vector, two for fp32) of one or more pixels, so a warp reads/writes 512 contiguous bytes per
// pixel; the nine taps of a pixel are re-read through L1/L2 (re-use distance (2*dil+1) rows << L2).
// Arithmetic is fp32; weights/bias are the fp32 parameters in their native [C,1,3,3] layout
// (72 contiguous floats per thread), so no per-call weight transposes or casts are launched.
#include <cuda_bf16.h>
#include <stdlib.h>

#include "rf_common.cuh"

namespace rf {

constexpr int DW_VEC = 8;  // channels per thread

template <typename T>
struct Vec8;
template <>
struct Vec8<__nv_bfloat16> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[8]) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __bfloat1622float2(h[i]);
      v[2 * i] = f.x;
      v[2 * i + 1] = f.y;
    }
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[8]) {
    uint4 u;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = u;
  }
};
template <>
struct Vec8<float> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[8]) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p));
    const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[8]) {
    reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
  }
};

// Exact-erf GELU (nn.GELU default, reference mix_transformer.py:85) with erf evaluated by the
// Abramowitz-Stegun 7.1.26 rational form (|error| <= 1.5e-7, i.e. fp32-erff accuracy) on one MUFU.RCP and
// one MUFU.EX2 instead of libdevice's branchy erff: these kernels are issue-bound, not HBM-bound.
//   erf(z) = 1 - (a1 t + a2 t^2 + a3 t^3 + a4 t^4 + a5 t^5) exp(-z^2),  t = 1 / (1 + p z),  z >= 0
// exp(-z^2) with z = |v| / sqrt(2) equals exp(-v^2 / 2), the Gaussian the derivative needs as well.
__device__ __forceinline__ void gelu_parts(float v, float& cdf, float& gauss) {
  const float z = fabsf(v) * 0.70710678118654752f;
  float t, den = fmaf(0.3275911f, z, 1.0f);
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(den));
  const float e = -0.72134752044448170f * v * v;     // log2(e) * (-v^2 / 2)
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(gauss) : "f"(e));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float erf_abs = fmaf(-poly * t, gauss, 1.0f);
  cdf = 0.5f * (1.0f + copysignf(erf_abs, v));
}
__device__ __forceinline__ float gelu_erf(float v) {
  float cdf, g;
  gelu_parts(v, cdf, g);
  return v * cdf;
}
__device__ __forceinline__ float gelu_erf_grad(float v) {
  float cdf, g;
  gelu_parts(v, cdf, g);
  return fmaf(v * 0.39894228040143268f, g, cdf);
}

// weights of 8 consecutive channels: native layout [C][9] -> w[k][tap]; 72 contiguous floats
__device__ __forceinline__ void load_w72(const float* __restrict__ w, int c0, float (&wr)[8][9]) {
  const float4* p = reinterpret_cast<const float4*>(w + (long)c0 * 9);
  float flat[72];
#pragma unroll
  for (int i = 0; i < 18; ++i) {
    const float4 q = __ldg(p + i);
    flat[4 * i] = q.x; flat[4 * i + 1] = q.y; flat[4 * i + 2] = q.z; flat[4 * i + 3] = q.w;
  }
#pragma unroll
  for (int k = 0; k < 8; ++k)
#pragma unroll
    for (int t = 0; t < 9; ++t) wr[k][t] = flat[k * 9 + t];
}

// MODE 0: y = conv(x) (+bias) (+GELU when act)           -- forward
// MODE 1: y = conv_flipped(x)                            -- input gradient of a plain depthwise conv
// MODE 2: y = aux * gelu'(conv(x) + bias)                -- GELU backward pre-pass (aux = dL/d(gelu out))
// PPT pixels per thread along x (weights stay in registers).
template <typename T, int MODE, int PPT>
__global__ void __launch_bounds__(256)
dwconv3x3_kernel(const T* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                 const T* __restrict__ aux, T* __restrict__ y, int B, int H, int W, int C, int dil, int act) {
  const int CG = C / DW_VEC;
  const int WX = (W + PPT - 1) / PPT;
  const long total = (long)B * H * WX * CG;
  const long gid = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (gid >= total) return;
  const int cg = (int)(gid % CG);
  long r = gid / CG;
  const int xs = (int)(r % WX) * PPT;
  r /= WX;
  const int yy = (int)(r % H);
  const int b = (int)(r / H);
  const int c0 = cg * DW_VEC;
  float wr[8][9];
  load_w72(w, c0, wr);
  float bv[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) bv[k] = (MODE != 1 && bias != nullptr) ? __ldg(bias + c0 + k) : 0.f;
  const T* xb = x + (long)b * H * W * C + c0;
#pragma unroll
  for (int p = 0; p < PPT; ++p) {
    const int xx = xs + p;
    if (xx >= W) break;
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = bv[k];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int iy = yy + (i - 1) * dil;
      if (iy < 0 || iy >= H) continue;
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const int ix = xx + (j - 1) * dil;
        if (ix < 0 || ix >= W) continue;
        float v[8];
        Vec8<T>::load(xb + ((long)iy * W + ix) * C, v);
        const int tap = (MODE == 1) ? (8 - (i * 3 + j)) : (i * 3 + j);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = fmaf(v[k], wr[k][tap], acc[k]);
      }
    }
    const long o = (((long)b * H + yy) * W + xx) * C + c0;
    if (MODE == 0) {
      if (act) {
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = gelu_erf(acc[k]);
      }
    } else if (MODE == 2) {
      float g[8];
      Vec8<T>::load(aux + o, g);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] = g[k] * gelu_erf_grad(acc[k]);
    }
    Vec8<T>::store(y + o, acc);
  }
}

// Dilation-1 specialisation: a thread owns 8 channels of PPT consecutive pixels of one row and slides a
// 3-row x (PPT+2)-column window over them, so every input vector is loaded once per thread
// (3*(PPT+2)/PPT = 3.75 loads per output at PPT = 8 instead of 9).
template <typename T, int MODE, int PPT>
__global__ void __launch_bounds__(128)
dwconv3x3_d1_kernel(const T* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                    const T* __restrict__ aux, T* __restrict__ y, int B, int H, int W, int C, int act) {
  const int CG = C / DW_VEC;
  const int WX = (W + PPT - 1) / PPT;
  const long total = (long)B * H * WX * CG;
  const long gid = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (gid >= total) return;
  const int cg = (int)(gid % CG);
  long r = gid / CG;
  const int xs = (int)(r % WX) * PPT;
  r /= WX;
  const int yy = (int)(r % H);
  const int b = (int)(r / H);
  const int c0 = cg * DW_VEC;
  float wr[8][9];
  load_w72(w, c0, wr);
  float acc[PPT][8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float bk = (MODE != 1 && bias != nullptr) ? __ldg(bias + c0 + k) : 0.f;
#pragma unroll
    for (int p = 0; p < PPT; ++p) acc[p][k] = bk;
  }
  const T* xb = x + (long)b * H * W * C + c0;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int iy = yy + i - 1;
    if (iy < 0 || iy >= H) continue;
    const T* row = xb + (long)iy * W * C;
#pragma unroll
    for (int q = 0; q < PPT + 2; ++q) {
      const int ix = xs + q - 1;
      if (ix < 0 || ix >= W) continue;
      float v[8];
      Vec8<T>::load(row + (long)ix * C, v);
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const int p = q - j;  // output pixel xs+p takes input column xs+p+j-1 == ix at tap j
        if (p < 0 || p >= PPT) continue;
        const int tap = (MODE == 1) ? (8 - (i * 3 + j)) : (i * 3 + j);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[p][k] = fmaf(v[k], wr[k][tap], acc[p][k]);
      }
    }
  }
#pragma unroll
  for (int p = 0; p < PPT; ++p) {
    const int xx = xs + p;
    if (xx >= W) break;
    const long o = (((long)b * H + yy) * W + xx) * C + c0;
    if (MODE == 0) {
      if (act) {
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[p][k] = gelu_erf(acc[p][k]);
      }
    } else if (MODE == 2) {
      float g[8];
      Vec8<T>::load(aux + o, g);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[p][k] = g[k] * gelu_erf_grad(acc[p][k]);
    }
    Vec8<T>::store(y + o, acc[p]);
  }
}

// dw[c][tap] += sum_pixels g[p,c] * x[p + off(tap), c];  db[c] += sum_pixels g[p,c]
// CTA = 32 channel groups (256 channels, lanes along channels -> 512 B per warp request) x 8 pixel
// lanes; each thread walks `strip`/8 pixels with fp32 register accumulators, the 8 pixel lanes are
// combined with shared-memory atomics and the CTA issues one red.global per (channel, tap).
constexpr int WG_CG = 32, WG_PL = 8;
template <typename T>
__global__ void __launch_bounds__(WG_CG * WG_PL)
dwconv3x3_wgrad_kernel(const T* __restrict__ x, const T* __restrict__ g, float* __restrict__ dw,
                       float* __restrict__ db, int B, int H, int W, int C, int dil, int strip) {
  __shared__ float red[80][WG_CG];  // 72 weight taps + 8 bias slots per channel group
  const int CG = C / DW_VEC;
  const long npix = (long)B * H * W;
  const int lane_cg = threadIdx.x % WG_CG, pl = threadIdx.x / WG_CG;
  const int cg = blockIdx.x * WG_CG + lane_cg;
  const bool live = cg < CG;
  const int c0 = cg * DW_VEC;
  for (int i = threadIdx.x; i < 80 * WG_CG; i += WG_CG * WG_PL) (&red[0][0])[i] = 0.f;
  __syncthreads();
  float acc[8][9];
  float accb[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    accb[k] = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) acc[k][t] = 0.f;
  }
  const long p0 = (long)blockIdx.y * strip;
  const long p1 = (p0 + strip < npix) ? p0 + strip : npix;
  if (live) {
    for (long p = p0 + pl; p < p1; p += WG_PL) {
      const int xx = (int)(p % W);
      const int yy = (int)((p / W) % H);
      const long bbase = (p / ((long)W * H)) * (long)H * W;
      float gv[8];
      Vec8<T>::load(g + p * C + c0, gv);
#pragma unroll
      for (int k = 0; k < 8; ++k) accb[k] += gv[k];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int iy = yy + (i - 1) * dil;
        if (iy < 0 || iy >= H) continue;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const int ix = xx + (j - 1) * dil;
          if (ix < 0 || ix >= W) continue;
          float v[8];
          Vec8<T>::load(x + (bbase + (long)iy * W + ix) * C + c0, v);
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[k][i * 3 + j] = fmaf(gv[k], v[k], acc[k][i * 3 + j]);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
#pragma unroll
      for (int t = 0; t < 9; ++t) atomicAdd(&red[k * 9 + t][lane_cg], acc[k][t]);
      atomicAdd(&red[72 + k][lane_cg], accb[k]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 80 * WG_CG; i += WG_CG * WG_PL) {
    const int e = i / WG_CG, l = i % WG_CG;
    const int gcg = blockIdx.x * WG_CG + l;
    if (gcg >= CG) continue;
    const float v = red[e][l];
    if (e < 72) {
      atomicAdd(dw + (long)(gcg * DW_VEC + e / 9) * 9 + e % 9, v);
    } else if (db != nullptr) {
      atomicAdd(db + gcg * DW_VEC + (e - 72), v);
    }
  }
}

// Dilated (ASPP, d = 6 / 12 / 18) variants.  Neighbouring pixels share no taps, but pixels d apart do:
// a thread owns 8 channels of a 4 x 2 block of outputs on the dilated lattice
//   (y0 + a d, x0 + b d), a < 4, b < 2,
// whose taps are the 6 x 4 lattice points (y0 + (a-1) d, x0 + (b-1) d): 24 vector loads for 8 outputs
// (3 per output instead of 9), which is what bounds these kernels (L2 -> SM traffic).
constexpr int DL_A = 4, DL_B = 2;
struct DilGeom {
  int yblocks, xblocks;   // lattice blocks per image: d * ceil(ceil(H/d) / DL_A), d * ceil(ceil(W/d) / DL_B)
};
__host__ __device__ inline DilGeom dil_geom(int H, int W, int d) {
  DilGeom g;
  g.yblocks = d * (((H + d - 1) / d + DL_A - 1) / DL_A);
  g.xblocks = d * (((W + d - 1) / d + DL_B - 1) / DL_B);
  return g;
}

template <typename T, int MODE>
__global__ void __launch_bounds__(128)
dwconv3x3_dil_kernel(const T* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                     T* __restrict__ y, int B, int H, int W, int C, int dil, int act) {
  const int CG = C / DW_VEC;
  const DilGeom g = dil_geom(H, W, dil);
  const long total = (long)B * g.yblocks * g.xblocks * CG;
  const long gid = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (gid >= total) return;
  const int cg = (int)(gid % CG);
  long r = gid / CG;
  const int xb = (int)(r % g.xblocks);
  r /= g.xblocks;
  const int yb = (int)(r % g.yblocks);
  const int b = (int)(r / g.yblocks);
  const int y0 = (yb % dil) + (yb / dil) * DL_A * dil;
  const int x0 = (xb % dil) + (xb / dil) * DL_B * dil;
  const int c0 = cg * DW_VEC;
  float wr[8][9];
  load_w72(w, c0, wr);
  float acc[DL_A][DL_B][8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float bk = (MODE != 1 && bias != nullptr) ? __ldg(bias + c0 + k) : 0.f;
#pragma unroll
    for (int a = 0; a < DL_A; ++a)
#pragma unroll
      for (int bb = 0; bb < DL_B; ++bb) acc[a][bb][k] = bk;
  }
  const T* xbase = x + (long)b * H * W * C + c0;
#pragma unroll
  for (int p = 0; p < DL_A + 2; ++p) {
    const int iy = y0 + (p - 1) * dil;
    if (iy < 0 || iy >= H) continue;
#pragma unroll
    for (int q = 0; q < DL_B + 2; ++q) {
      const int ix = x0 + (q - 1) * dil;
      if (ix < 0 || ix >= W) continue;
      float v[8];
      Vec8<T>::load(xbase + ((long)iy * W + ix) * C, v);
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int a = p - i;
        if (a < 0 || a >= DL_A) continue;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const int bb = q - j;
          if (bb < 0 || bb >= DL_B) continue;
          const int tap = (MODE == 1) ? (8 - (i * 3 + j)) : (i * 3 + j);
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[a][bb][k] = fmaf(v[k], wr[k][tap], acc[a][bb][k]);
        }
      }
    }
  }
#pragma unroll
  for (int a = 0; a < DL_A; ++a) {
    const int yy = y0 + a * dil;
    if (yy >= H) continue;
#pragma unroll
    for (int bb = 0; bb < DL_B; ++bb) {
      const int xx = x0 + bb * dil;
      if (xx >= W) continue;
      if (MODE == 0 && act) {
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[a][bb][k] = gelu_erf(acc[a][bb][k]);
      }
      Vec8<T>::store(y + (((long)b * H + yy) * W + xx) * C + c0, acc[a][bb]);
    }
  }
}

// Dilated weight gradient on the same 4 x 2 lattice blocks (24 + 8 loads per 8 pixels instead of 80);
// CTA = 32 channel groups x 8 block lanes, shared-memory + one red.global per (channel, tap) per CTA.
template <typename T>
__global__ void __launch_bounds__(WG_CG * WG_PL)
dwconv3x3_dil_wgrad_kernel(const T* __restrict__ x, const T* __restrict__ g, float* __restrict__ dw,
                           float* __restrict__ db, int B, int H, int W, int C, int dil, int blocks_per_cta) {
  __shared__ float red[80][WG_CG];
  const int CG = C / DW_VEC;
  const int lane_cg = threadIdx.x % WG_CG, pl = threadIdx.x / WG_CG;
  const int cg = blockIdx.x * WG_CG + lane_cg;
  const bool live = cg < CG;
  const int c0 = cg * DW_VEC;
  const DilGeom geo = dil_geom(H, W, dil);
  const long nblk = (long)B * geo.yblocks * geo.xblocks;
  for (int i = threadIdx.x; i < 80 * WG_CG; i += WG_CG * WG_PL) (&red[0][0])[i] = 0.f;
  __syncthreads();
  if (live) {
    float acc[8][9];
    float accb[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      accb[k] = 0.f;
#pragma unroll
      for (int t = 0; t < 9; ++t) acc[k][t] = 0.f;
    }
    const long n0 = (long)blockIdx.y * blocks_per_cta;
    const long n1 = (n0 + blocks_per_cta < nblk) ? n0 + blocks_per_cta : nblk;
    for (long n = n0 + pl; n < n1; n += WG_PL) {
      const int xb = (int)(n % geo.xblocks);
      const int yb = (int)((n / geo.xblocks) % geo.yblocks);
      const int b = (int)(n / ((long)geo.xblocks * geo.yblocks));
      const int y0 = (yb % dil) + (yb / dil) * DL_A * dil;
      const int x0 = (xb % dil) + (xb / dil) * DL_B * dil;
      const T* xbase = x + (long)b * H * W * C + c0;
      const T* gbase = g + (long)b * H * W * C + c0;
      float gv[DL_A][DL_B][8];
#pragma unroll
      for (int a = 0; a < DL_A; ++a)
#pragma unroll
        for (int bb = 0; bb < DL_B; ++bb) {
          const int yy = y0 + a * dil, xx = x0 + bb * dil;
          if (yy < H && xx < W) {
            Vec8<T>::load(gbase + ((long)yy * W + xx) * C, gv[a][bb]);
          } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) gv[a][bb][k] = 0.f;
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) accb[k] += gv[a][bb][k];
        }
#pragma unroll
      for (int p = 0; p < DL_A + 2; ++p) {
        const int iy = y0 + (p - 1) * dil;
        if (iy < 0 || iy >= H) continue;
#pragma unroll
        for (int q = 0; q < DL_B + 2; ++q) {
          const int ix = x0 + (q - 1) * dil;
          if (ix < 0 || ix >= W) continue;
          float v[8];
          Vec8<T>::load(xbase + ((long)iy * W + ix) * C, v);
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            const int a = p - i;
            if (a < 0 || a >= DL_A) continue;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              const int bb = q - j;
              if (bb < 0 || bb >= DL_B) continue;
#pragma unroll
              for (int k = 0; k < 8; ++k) acc[k][i * 3 + j] = fmaf(gv[a][bb][k], v[k], acc[k][i * 3 + j]);
            }
          }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
#pragma unroll
      for (int t = 0; t < 9; ++t) atomicAdd(&red[k * 9 + t][lane_cg], acc[k][t]);
      atomicAdd(&red[72 + k][lane_cg], accb[k]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 80 * WG_CG; i += WG_CG * WG_PL) {
    const int e = i / WG_CG, l = i % WG_CG;
    const int gcg = blockIdx.x * WG_CG + l;
    if (gcg >= CG) continue;
    const float v = red[e][l];
    if (e < 72) {
      atomicAdd(dw + (long)(gcg * DW_VEC + e / 9) * 9 + e % 9, v);
    } else if (db != nullptr) {
      atomicAdd(db + gcg * DW_VEC + (e - 72), v);
    }
  }
}

// Dilation-1 weight gradient: a thread owns 8 channels of PPT consecutive pixels of ROWS rows and slides
// the 3 x (PPT+2) input window like the forward, so each input vector is loaded once per thread
// ((3*(PPT+2) + PPT) / PPT = 4.75 loads per pixel instead of 10).  CTA = 32 channel groups x 8 pixel
// strips; partial sums are combined with shared-memory atomics, one red.global per (channel, tap) per CTA.
template <typename T, int PPT, int ROWS>
__global__ void __launch_bounds__(WG_CG * WG_PL)
dwconv3x3_d1_wgrad_kernel(const T* __restrict__ x, const T* __restrict__ g, float* __restrict__ dw,
                          float* __restrict__ db, int B, int H, int W, int C) {
  __shared__ float red[80][WG_CG];
  const int CG = C / DW_VEC;
  const int lane_cg = threadIdx.x % WG_CG, pl = threadIdx.x / WG_CG;
  const int cg = blockIdx.x * WG_CG + lane_cg;
  const bool live = cg < CG;
  const int c0 = cg * DW_VEC;
  const int WX = (W + PPT - 1) / PPT;          // pixel strips per row
  const int SG = (WX + WG_PL - 1) / WG_PL;     // strip groups per row (WG_PL strips per CTA)
  const int RG = (H + ROWS - 1) / ROWS;
  int t = blockIdx.y;
  const int sg = t % SG;
  t /= SG;
  const int rg = t % RG;
  const int b = t / RG;
  const int xs = (sg * WG_PL + pl) * PPT;
  for (int i = threadIdx.x; i < 80 * WG_CG; i += WG_CG * WG_PL) (&red[0][0])[i] = 0.f;
  __syncthreads();
  if (live && xs < W) {
    float acc[8][9];
    float accb[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      accb[k] = 0.f;
#pragma unroll
      for (int tt = 0; tt < 9; ++tt) acc[k][tt] = 0.f;
    }
    const T* xb = x + (long)b * H * W * C + c0;
    const T* gb = g + (long)b * H * W * C + c0;
    for (int r = 0; r < ROWS; ++r) {
      const int yy = rg * ROWS + r;
      if (yy >= H) break;
      float gv[PPT][8];
#pragma unroll
      for (int p = 0; p < PPT; ++p) {
        if (xs + p < W) {
          Vec8<T>::load(gb + ((long)yy * W + xs + p) * C, gv[p]);
        } else {
#pragma unroll
          for (int k = 0; k < 8; ++k) gv[p][k] = 0.f;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) accb[k] += gv[p][k];
      }
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int iy = yy + i - 1;
        if (iy < 0 || iy >= H) continue;
        const T* row = xb + (long)iy * W * C;
#pragma unroll
        for (int q = 0; q < PPT + 2; ++q) {
          const int ix = xs + q - 1;
          if (ix < 0 || ix >= W) continue;
          float v[8];
          Vec8<T>::load(row + (long)ix * C, v);
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            const int p = q - j;  // output pixel xs+p sees input column ix through tap j
            if (p < 0 || p >= PPT) continue;
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k][i * 3 + j] = fmaf(gv[p][k], v[k], acc[k][i * 3 + j]);
          }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
#pragma unroll
      for (int tt = 0; tt < 9; ++tt) atomicAdd(&red[k * 9 + tt][lane_cg], acc[k][tt]);
      atomicAdd(&red[72 + k][lane_cg], accb[k]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 80 * WG_CG; i += WG_CG * WG_PL) {
    const int e = i / WG_CG, l = i % WG_CG;
    const int gcg = blockIdx.x * WG_CG + l;
    if (gcg >= CG) continue;
    const float v = red[e][l];
    if (e < 72) {
      atomicAdd(dw + (long)(gcg * DW_VEC + e / 9) * 9 + e % 9, v);
    } else if (db != nullptr) {
      atomicAdd(db + gcg * DW_VEC + (e - 72), v);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Shared-memory tiled dilation-1 kernels for bf16 (the Mix-FFN DWConv of every MiT block: the largest
// HBM-streaming share of the train step).  These kernels are ISSUE-bound on B200, not bandwidth-bound: at
// 6.5 TB/s one SM has ~22 issue slots per bf16 element streamed in and out, and a 3x3 depthwise tap set plus an
// exact-erf GELU costs more than that in scalar fp32.  So the design minimises instructions per element:
//   * the (8+2) x (TW+2) x 64-channel input tile (128 B per pixel) is staged once with 16-byte cp.async
//     (zero-fill outside the image = the conv padding), which replaces the 3.75x re-read through L1 with
//     conflict-free LDS.128 and puts the whole tile's bytes in flight at once;
//   * all arithmetic is packed f32x2 (FFMA2/FMUL2/FADD2, sm_100): a bf16x2 word unpacks to one channel pair
//     and the weights are staged as [tap][channel] so that a pair of adjacent channels is one 64-bit operand;
//   * GELU for bf16 outputs uses Phi(-t) = 2^Q(t), t = min(|v|, 6), Q a degree-6 minimax fit of
//     log2(0.5 erfc(t / sqrt 2)): gelu(v) = relu(v) - t * 2^Q(t), one MUFU.EX2 per element and
//     |error| <= 7e-6 absolute / 5e-5 relative (bf16 rounding is 2e-3 relative); the fp32 instantiations used
//     by the fp32 parity runs keep the 1.5e-7-accurate erf above.
// Thread map (TW * 8 threads): chunk = tid & 7 owns 8 channels (16 bytes), run = tid >> 3 owns 8 consecutive
// output pixels of one tile row; a quarter-warp reads the 128 contiguous bytes of one pixel.
constexpr int DT_TH = 8;    // tile rows
constexpr int DT_TC = 64;   // channels per tile (bf16: 128 bytes per pixel)
constexpr int DT_RUN = 8;   // pixels per thread

__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 f2s(float a) { return make_float2(a, a); }

// bf16x2 word -> (lo, hi) channel pair in fp32: one shift and one mask
__device__ __forceinline__ float2 bf2_unpack(unsigned u) {
  return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}
__device__ __forceinline__ unsigned bf2_pack(float2 v) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(v.x, v.y);
  return *reinterpret_cast<const unsigned*>(&h);
}

// e = Phi(-t) = 0.5 erfc(t / sqrt 2) for the lane-wise clamped magnitudes t = min(|v|, 6)
__device__ __forceinline__ float2 normal_tail2(float2 t) {
  float2 q = f2s(2.2999249267741106e-05f);
  q = __ffma2_rn(q, t, f2s(-0.0006114901625551283f));
  q = __ffma2_rn(q, t, f2s(0.007200188934803009f));
  q = __ffma2_rn(q, t, f2s(-0.05120821297168732f));
  q = __ffma2_rn(q, t, f2s(-0.46122226119041443f));
  q = __ffma2_rn(q, t, f2s(-1.150214433670044f));
  q = __ffma2_rn(q, t, f2s(-1.000058889389038f));
  float2 e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(q.x));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(q.y));
  return e;
}
// gelu(v) = relu(v) - t * Phi(-t)
__device__ __forceinline__ float2 gelu_fast2(float2 v) {
  const float2 t = f2(fminf(fabsf(v.x), 6.0f), fminf(fabsf(v.y), 6.0f));
  const float2 e = normal_tail2(t);
  const float2 r = f2(fmaxf(v.x, 0.0f), fmaxf(v.y, 0.0f));
  return __ffma2_rn(f2(-t.x, -t.y), e, r);
}
// gelu'(v) = Phi(v) + v phi(v),  Phi(v) = 0.5 + copysign(0.5 - Phi(-t), v),  phi(v) = exp(-v^2/2) / sqrt(2 pi)
__device__ __forceinline__ float2 gelu_fast_grad2(float2 v) {
  const float2 t = f2(fminf(fabsf(v.x), 6.0f), fminf(fabsf(v.y), 6.0f));
  const float2 e = normal_tail2(t);
  const float2 h = __fadd2_rn(f2s(0.5f), f2(-e.x, -e.y));
  const float2 cdf = __fadd2_rn(f2s(0.5f), f2(copysignf(h.x, v.x), copysignf(h.y, v.y)));
  const float2 a = __fmul2_rn(__fmul2_rn(v, v), f2s(-0.72134752044448170f));
  float2 g;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(g.x) : "f"(a.x));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(g.y) : "f"(a.y));
  return __ffma2_rn(__fmul2_rn(v, f2s(0.39894228040143268f)), g, cdf);
}

// Stage ROWS rows x (TW + 2 * HALO) columns x 8 chunks of a channels-last bf16 image into shared memory with
// 16-byte cp.async, zero-filling outside the image.  Thread (chunk = tid & 7, slot = tid >> 3) copies tile column
// `slot` of every row (plus one of the 2 * HALO extra columns for the first slots): no div / mod in the loop.
template <int TW, int ROWS, int HALO>
__device__ __forceinline__ void stage_tile(uint4* tile, const __nv_bfloat16* __restrict__ img, int x0, int y0, int Hs,
                                           int Ws, long pix_pitch, long row_pitch, int tid) {
  constexpr int IW = TW + 2 * HALO;
  const int chunk = tid & 7, slot = tid >> 3;
  const int ix = x0 + slot - HALO, ix2 = x0 + TW + slot - HALO;
  const bool okx = ix >= 0 && ix < Ws, okx2 = slot < 2 * HALO && ix2 < Ws;
  const __nv_bfloat16* col = img + ix * pix_pitch + chunk * 8;
  const __nv_bfloat16* col2 = img + ix2 * pix_pitch + chunk * 8;
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    const int iy = y0 + r - HALO;
    const bool oky = iy >= 0 && iy < Hs;
    const long ro = iy * row_pitch;
    cp_async16(&tile[(r * IW + slot) * 8 + chunk], (oky && okx) ? col + ro : img, (oky && okx) ? 16 : 0);
    if (slot < 2 * HALO)
      cp_async16(&tile[(r * IW + TW + slot) * 8 + chunk], (oky && okx2) ? col2 + ro : img, (oky && okx2) ? 16 : 0);
  }
}

// A dilation-d 3x3 depthwise conv is d*d independent dilation-1 convs, one per residue class
// (y mod d, x mod d): the pixels of a class form a sub-image of ceil((H-ry)/d) x ceil((W-rx)/d) pixels with
// pixel pitch d*C and row pitch d*W*C.  The tile kernels below walk sub-images, so the ASPP branches
// (d = 6 / 12 / 18) run through the same shared-memory tiles as the Mix-FFN conv (d = 1: one class).
struct SubImage {
  int Hs, Ws;               // sub-image size
  long pix_pitch, row_pitch;
  long origin;              // element offset of sub-image pixel (0, 0) channel 0 within the batch image
};
__device__ __forceinline__ SubImage sub_image(int H, int W, int C, int dil, int ry, int rx) {
  SubImage s;
  s.Hs = (H - ry + dil - 1) / dil;
  s.Ws = (W - rx + dil - 1) / dil;
  s.pix_pitch = (long)dil * C;
  s.row_pitch = (long)dil * W * C;
  s.origin = ((long)ry * W + rx) * C;
  return s;
}

template <int MODE, int TW>
__global__ void __launch_bounds__(TW * 8, 64 / TW)
dwconv3x3_tile_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w,
                      const float* __restrict__ bias, const __nv_bfloat16* __restrict__ aux,
                      __nv_bfloat16* __restrict__ y, int B, int H, int W, int C, int dil, int act) {
  constexpr int NT = TW * 8;
  constexpr int IW = TW + 2, IH = DT_TH + 2;
  __shared__ __align__(16) uint4 tile[IH * IW * 8];      // [row][col][chunk], 16 bytes = 8 bf16 channels
  __shared__ __align__(16) float wsm[10][DT_TC];         // [tap][channel]; row 9 = bias
  const int tid = threadIdx.x;
  const int c0 = blockIdx.x * DT_TC;
  const int tiles_x = ((W + dil - 1) / dil + TW - 1) / TW, tiles_y = ((H + dil - 1) / dil + DT_TH - 1) / DT_TH;
  int t = blockIdx.y;
  const int tx = t % tiles_x;
  t /= tiles_x;
  const int ty = t % tiles_y;
  t /= tiles_y;
  const SubImage si = sub_image(H, W, C, dil, t / dil, t % dil);
  const int b = blockIdx.z;
  const int x0 = tx * TW, y0 = ty * DT_TH;
  const __nv_bfloat16* xb = x + (long)b * H * W * C + si.origin + c0;
  // ---- stage the input tile (zero-fill = padding) ------------------------------------------------------
  stage_tile<TW, IH, 1>(tile, xb, x0, y0, si.Hs, si.Ws, si.pix_pitch, si.row_pitch, tid);
  cp_async_commit();
  // ---- weights [C][9] -> [tap][channel] (taps flipped for the input gradient), bias -------------------
  for (int i = tid; i < DT_TC * 9; i += NT) {
    const int ch = i / 9, tap = i % 9;
    wsm[MODE == 1 ? 8 - tap : tap][ch] = __ldg(w + (long)c0 * 9 + i);
  }
  if (tid < DT_TC) wsm[9][tid] = (MODE != 1 && bias != nullptr) ? __ldg(bias + c0 + tid) : 0.f;
  cp_async_wait<0>();
  __syncthreads();
  // ---- compute -------------------------------------------------------------------------------------
  const int chunk = tid & 7, run = tid >> 3;
  const int row = run / (TW / DT_RUN), col0 = (run % (TW / DT_RUN)) * DT_RUN;
  float2 acc[DT_RUN][4];
  {
    const float4 b0 = *reinterpret_cast<const float4*>(&wsm[9][chunk * 8]);
    const float4 b1 = *reinterpret_cast<const float4*>(&wsm[9][chunk * 8 + 4]);
#pragma unroll
    for (int p = 0; p < DT_RUN; ++p) {
      acc[p][0] = f2(b0.x, b0.y); acc[p][1] = f2(b0.z, b0.w);
      acc[p][2] = f2(b1.x, b1.y); acc[p][3] = f2(b1.z, b1.w);
    }
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    float2 wr[3][4];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float4 a0 = *reinterpret_cast<const float4*>(&wsm[i * 3 + j][chunk * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&wsm[i * 3 + j][chunk * 8 + 4]);
      wr[j][0] = f2(a0.x, a0.y); wr[j][1] = f2(a0.z, a0.w);
      wr[j][2] = f2(a1.x, a1.y); wr[j][3] = f2(a1.z, a1.w);
    }
    const uint4* trow = &tile[((row + i) * IW + col0) * 8 + chunk];
#pragma unroll
    for (int q = 0; q < DT_RUN + 2; ++q) {
      const uint4 u = trow[q * 8];
      const float2 v0 = bf2_unpack(u.x), v1 = bf2_unpack(u.y), v2 = bf2_unpack(u.z), v3 = bf2_unpack(u.w);
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const int p = q - j;  // output pixel col0+p reads input column col0+p+j-1, tile column col0+p+j == col0+q
        if (p < 0 || p >= DT_RUN) continue;
        acc[p][0] = __ffma2_rn(v0, wr[j][0], acc[p][0]);
        acc[p][1] = __ffma2_rn(v1, wr[j][1], acc[p][1]);
        acc[p][2] = __ffma2_rn(v2, wr[j][2], acc[p][2]);
        acc[p][3] = __ffma2_rn(v3, wr[j][3], acc[p][3]);
      }
    }
  }
  // ---- epilogue ------------------------------------------------------------------------------------
  const int oy = y0 + row;
  if (oy >= si.Hs) return;
  const long obase = (long)b * H * W * C + si.origin + oy * si.row_pitch + (x0 + col0) * si.pix_pitch + c0 + chunk * 8;
#pragma unroll
  for (int p = 0; p < DT_RUN; ++p) {
    if (x0 + col0 + p >= si.Ws) break;
    if (MODE == 0) {
      if (act) {
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[p][k] = gelu_fast2(acc[p][k]);
      }
    } else if (MODE == 2) {
      const uint4 g = __ldg(reinterpret_cast<const uint4*>(aux + obase + p * si.pix_pitch));
      acc[p][0] = __fmul2_rn(bf2_unpack(g.x), gelu_fast_grad2(acc[p][0]));
      acc[p][1] = __fmul2_rn(bf2_unpack(g.y), gelu_fast_grad2(acc[p][1]));
      acc[p][2] = __fmul2_rn(bf2_unpack(g.z), gelu_fast_grad2(acc[p][2]));
      acc[p][3] = __fmul2_rn(bf2_unpack(g.w), gelu_fast_grad2(acc[p][3]));
    }
    uint4 o;
    o.x = bf2_pack(acc[p][0]); o.y = bf2_pack(acc[p][1]); o.z = bf2_pack(acc[p][2]); o.w = bf2_pack(acc[p][3]);
    *reinterpret_cast<uint4*>(y + obase + p * si.pix_pitch) = o;
  }
}

template <int MODE>
static int launch_dw_tile(const void* x, const float* w, const float* bias, const void* aux, void* y, int B, int H,
                          int W, int C, int dil, int act, cudaStream_t st, const char* name) {
  const int Ws = (W + dil - 1) / dil, Hs = (H + dil - 1) / dil;   // largest residue-class sub-image
  const bool narrow = dil > 1 || ((Ws % 32 != 0) && (Ws <= 16 || Ws % 32 <= 16));
  const int TW = narrow ? 16 : 32;
  const long tiles = (long)((Ws + TW - 1) / TW) * ((Hs + DT_TH - 1) / DT_TH) * dil * dil;
  RF_REQUIRE(tiles <= 65535 && B <= 65535, "%s: grid too large", name);
  dim3 grid((unsigned)(C / DT_TC), (unsigned)tiles, (unsigned)B);
  if (narrow)
    dwconv3x3_tile_kernel<MODE, 16><<<grid, 128, 0, st>>>((const __nv_bfloat16*)x, w, bias, (const __nv_bfloat16*)aux,
                                                         (__nv_bfloat16*)y, B, H, W, C, dil, act);
  else
    dwconv3x3_tile_kernel<MODE, 32><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, w, bias, (const __nv_bfloat16*)aux,
                                                         (__nv_bfloat16*)y, B, H, W, C, dil, act);
  RF_CHECK_LAUNCH(name);
  return RF_OK;
}

// Weight / bias gradient on the same tiles: dw[c][tap] = sum_p g[p,c] x[p + off(tap), c], db[c] = sum_p g[p,c].
// One CTA (128 threads = 8 chunks x 16 runs, tile 8 x 16 pixels x 64 channels) walks a strided list of tiles of
// one 64-channel block with its 9 x 8 + 8 partial sums in registers (packed f32x2), staging the x tile (with
// halo) and the g tile with cp.async; 3 CTAs per SM overlap each other's staging.  At the end the 16 runs are
// combined by warp shuffles + shared memory and the CTA issues one red.global per (channel, tap).
constexpr int WT_TW = 16;
__global__ void __launch_bounds__(128, 3)
dwconv3x3_tile_wgrad_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ g,
                            float* __restrict__ dw, float* __restrict__ db, int B, int H, int W, int C, int dil,
                            int tiles_x, int tiles_y) {
  constexpr int TW = WT_TW;
  constexpr int IW = TW + 2, IH = DT_TH + 2;
  __shared__ __align__(16) uint4 xt[IH * IW * 8];
  __shared__ __align__(16) uint4 gt[DT_TH * TW * 8];
  float (*red)[80][8] = reinterpret_cast<float (*)[80][8]>(xt);   // [warp][tap*8 + k | 72 + k][chunk], reuses xt
  const int tid = threadIdx.x;
  const int c0 = blockIdx.x * DT_TC;
  const int chunk = tid & 7, run = tid >> 3;
  const int row = run / (TW / DT_RUN), col0 = (run % (TW / DT_RUN)) * DT_RUN;
  float2 wacc[9][4], bacc[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    bacc[k] = f2s(0.f);
#pragma unroll
    for (int t = 0; t < 9; ++t) wacc[t][k] = f2s(0.f);
  }
  const int ntiles = tiles_x * tiles_y * dil * dil * B;
  for (int t = blockIdx.y; t < ntiles; t += gridDim.y) {
    int u = t;
    const int tx = u % tiles_x;
    u /= tiles_x;
    const int ty = u % tiles_y;
    u /= tiles_y;
    const int res = u % (dil * dil), b = u / (dil * dil);
    const SubImage si = sub_image(H, W, C, dil, res / dil, res % dil);
    const int x0 = tx * TW, y0 = ty * DT_TH;
    const __nv_bfloat16* xb = x + (long)b * H * W * C + si.origin + c0;
    const __nv_bfloat16* gb = g + (long)b * H * W * C + si.origin + c0;
    __syncthreads();   // the previous tile's readers are done
    stage_tile<TW, IH, 1>(xt, xb, x0, y0, si.Hs, si.Ws, si.pix_pitch, si.row_pitch, tid);
    stage_tile<TW, DT_TH, 0>(gt, gb, x0, y0, si.Hs, si.Ws, si.pix_pitch, si.row_pitch, tid);   // outside: g = 0
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    float2 gv[DT_RUN][4];
#pragma unroll
    for (int p = 0; p < DT_RUN; ++p) {
      const uint4 u = gt[((row * TW) + col0 + p) * 8 + chunk];
      gv[p][0] = bf2_unpack(u.x); gv[p][1] = bf2_unpack(u.y); gv[p][2] = bf2_unpack(u.z); gv[p][3] = bf2_unpack(u.w);
#pragma unroll
      for (int k = 0; k < 4; ++k) bacc[k] = __fadd2_rn(bacc[k], gv[p][k]);
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const uint4* trow = &xt[((row + i) * IW + col0) * 8 + chunk];
#pragma unroll
      for (int q = 0; q < DT_RUN + 2; ++q) {
        const uint4 u = trow[q * 8];
        const float2 v0 = bf2_unpack(u.x), v1 = bf2_unpack(u.y), v2 = bf2_unpack(u.z), v3 = bf2_unpack(u.w);
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const int p = q - j;
          if (p < 0 || p >= DT_RUN) continue;
          wacc[i * 3 + j][0] = __ffma2_rn(gv[p][0], v0, wacc[i * 3 + j][0]);
          wacc[i * 3 + j][1] = __ffma2_rn(gv[p][1], v1, wacc[i * 3 + j][1]);
          wacc[i * 3 + j][2] = __ffma2_rn(gv[p][2], v2, wacc[i * 3 + j][2]);
          wacc[i * 3 + j][3] = __ffma2_rn(gv[p][3], v3, wacc[i * 3 + j][3]);
        }
      }
    }
  }
  // ---- combine the 16 runs: lanes (chunk, run & 3) -> shuffle over the run bits, then the 4 warps in smem ----
  __syncthreads();   // xt is reused as the reduction buffer
  const int warp = tid >> 5, lane = tid & 31;
#pragma unroll
  for (int t = 0; t < 10; ++t) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float2 v = (t < 9) ? wacc[t][k] : bacc[k];
      v.x += __shfl_xor_sync(0xffffffffu, v.x, 8);
      v.y += __shfl_xor_sync(0xffffffffu, v.y, 8);
      v.x += __shfl_xor_sync(0xffffffffu, v.x, 16);
      v.y += __shfl_xor_sync(0xffffffffu, v.y, 16);
      if (lane < 8) {
        red[warp][t * 8 + 2 * k][lane] = v.x;
        red[warp][t * 8 + 2 * k + 1][lane] = v.y;
      }
    }
  }
  __syncthreads();
  for (int i = tid; i < 80 * 8; i += 128) {
    const int e = i >> 3, ch = i & 7;        // e = tap * 8 + k  (or 72 + k), channel = c0 + ch * 8 + k
    const float v = red[0][e][ch] + red[1][e][ch] + red[2][e][ch] + red[3][e][ch];
    const int k = e & 7, tap = e >> 3;
    const int c = c0 + ch * 8 + k;
    if (tap < 9) {
      atomicAdd(dw + (long)c * 9 + tap, v);
    } else if (db != nullptr) {
      atomicAdd(db + c, v);
    }
  }
}

// RF_DWCONV_IMPL=direct selects the register-window kernels for A/B measurements (read once)
static bool dw_use_tile() {
  static const bool v = [] {
    const char* e = getenv("RF_DWCONV_IMPL");
    return !(e && e[0] == 'd');
  }();
  return v;
}

template <typename T, int MODE>
static int launch_dw(const void* x, const float* w, const float* bias, const void* aux, void* y, int B, int H, int W,
                     int C, int dil, int act, cudaStream_t st, const char* name) {
  if (sizeof(T) == 2 && C % DT_TC == 0 && dw_use_tile() && dil <= H && dil <= W &&
      (long)dil * dil * (((W + dil - 1) / dil + 15) / 16) * (((H + dil - 1) / dil + 7) / 8) <= 65535)
    return launch_dw_tile<MODE>(x, w, bias, aux, y, B, H, W, C, dil, act, st, name);
  if (dil == 1) {
    constexpr int PPT = 8;
    const long total = (long)B * H * ((W + PPT - 1) / PPT) * (C / DW_VEC);
    const long blocks = (total + 127) / 128;
    RF_REQUIRE(blocks < (1l << 31), "%s: grid too large", name);
    dwconv3x3_d1_kernel<T, MODE, PPT><<<(unsigned)blocks, 128, 0, st>>>((const T*)x, w, bias, (const T*)aux, (T*)y, B, H,
                                                                          W, C, act);
    RF_CHECK_LAUNCH(name);
    return RF_OK;
  }
  if (MODE != 2 && dil >= 2) {
    const DilGeom g = dil_geom(H, W, dil);
    const long total = (long)B * g.yblocks * g.xblocks * (C / DW_VEC);
    const long blocks = (total + 127) / 128;
    RF_REQUIRE(blocks < (1l << 31), "%s: grid too large", name);
    dwconv3x3_dil_kernel<T, MODE><<<(unsigned)blocks, 128, 0, st>>>((const T*)x, w, bias, (T*)y, B, H, W, C, dil, act);
    RF_CHECK_LAUNCH(name);
    return RF_OK;
  }
  constexpr int PPT = 4;
  const long total = (long)B * H * ((W + PPT - 1) / PPT) * (C / DW_VEC);
  const long blocks = (total + 255) / 256;
  RF_REQUIRE(blocks < (1l << 31), "%s: grid too large", name);
  dwconv3x3_kernel<T, MODE, PPT><<<(unsigned)blocks, 256, 0, st>>>((const T*)x, w, bias, (const T*)aux, (T*)y, B, H, W,
                                                                     C, dil, act);
  RF_CHECK_LAUNCH(name);
  return RF_OK;
}

static int check_dw(const void* x, const void* w, const void* y, int B, int H, int W, int C, int dil, int dtype,
                    const char* name) {
  RF_REQUIRE(x && w && y, "%s: null pointer", name);
  RF_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && dil > 0, "%s: bad shape", name);
  RF_REQUIRE(C % DW_VEC == 0, "%s: C=%d must be a multiple of %d", name, C, DW_VEC);
  RF_REQUIRE(dtype == 0 || dtype == 1, "%s: dtype must be 0 (f32) or 1 (bf16)", name);
  RF_REQUIRE(((uintptr_t)x % 16 == 0) && ((uintptr_t)y % 16 == 0) && ((uintptr_t)w % 16 == 0),
             "%s: pointers must be 16-byte aligned", name);
  return RF_OK;
}

}  // namespace rf

using namespace rf;

extern "C" int rf_dwconv3x3_nhwc_fwd(const void* x, const float* weight, const float* bias, void* y, int B, int H,
                                     int W, int C, int dilation, int gelu, int dtype, void* stream) {
  int rc = check_dw(x, weight, y, B, H, W, C, dilation, dtype, "rf_dwconv3x3_nhwc_fwd");
  if (rc != RF_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == 1)
    return launch_dw<__nv_bfloat16, 0>(x, weight, bias, nullptr, y, B, H, W, C, dilation, gelu, st,
                                       "dwconv3x3_fwd<bf16>");
  return launch_dw<float, 0>(x, weight, bias, nullptr, y, B, H, W, C, dilation, gelu, st, "dwconv3x3_fwd<f32>");
}

extern "C" int rf_dwconv3x3_nhwc_bwd_input(const void* grad_y, const float* weight, void* grad_x, int B, int H, int W,
                                           int C, int dilation, int dtype, void* stream) {
  int rc = check_dw(grad_y, weight, grad_x, B, H, W, C, dilation, dtype, "rf_dwconv3x3_nhwc_bwd_input");
  if (rc != RF_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == 1)
    return launch_dw<__nv_bfloat16, 1>(grad_y, weight, nullptr, nullptr, grad_x, B, H, W, C, dilation, 0, st,
                                       "dwconv3x3_bwd_input<bf16>");
  return launch_dw<float, 1>(grad_y, weight, nullptr, nullptr, grad_x, B, H, W, C, dilation, 0, st,
                             "dwconv3x3_bwd_input<f32>");
}

extern "C" int rf_dwconv3x3_gelu_bwd_pre(const void* x, const float* weight, const float* bias, const void* grad_out,
                                         void* grad_pre, int B, int H, int W, int C, int dilation, int dtype,
                                         void* stream) {
  int rc = check_dw(x, weight, grad_pre, B, H, W, C, dilation, dtype, "rf_dwconv3x3_gelu_bwd_pre");
  if (rc != RF_OK) return rc;
  RF_REQUIRE(grad_out != nullptr, "rf_dwconv3x3_gelu_bwd_pre: null grad_out");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == 1)
    return launch_dw<__nv_bfloat16, 2>(x, weight, bias, grad_out, grad_pre, B, H, W, C, dilation, 1, st,
                                       "dwconv3x3_gelu_bwd_pre<bf16>");
  return launch_dw<float, 2>(x, weight, bias, grad_out, grad_pre, B, H, W, C, dilation, 1, st,
                             "dwconv3x3_gelu_bwd_pre<f32>");
}

extern "C" int rf_dwconv3x3_nhwc_bwd_weight(const void* x, const void* grad_pre, float* grad_weight, float* grad_bias,
                                            int B, int H, int W, int C, int dilation, int dtype, int accumulate,
                                            void* stream) {
  int rc = check_dw(x, grad_pre, grad_weight, B, H, W, C, dilation, dtype, "rf_dwconv3x3_nhwc_bwd_weight");
  if (rc != RF_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (!accumulate) {
    RF_CUDA(cudaMemsetAsync(grad_weight, 0, sizeof(float) * 9 * (size_t)C, st));
    if (grad_bias) RF_CUDA(cudaMemsetAsync(grad_bias, 0, sizeof(float) * (size_t)C, st));
  }
  const long npix = (long)B * H * W;
  const int CG = C / DW_VEC;
  const int gx = (CG + WG_CG - 1) / WG_CG;
  if (dtype == 1 && C % DT_TC == 0 && dw_use_tile() && dilation <= H && dilation <= W) {
    const int tiles_x = ((W + dilation - 1) / dilation + WT_TW - 1) / WT_TW;
    const int tiles_y = ((H + dilation - 1) / dilation + DT_TH - 1) / DT_TH;
    const long ntiles = (long)tiles_x * tiles_y * dilation * dilation * B;
    RF_REQUIRE(ntiles < (1l << 31), "rf_dwconv3x3_nhwc_bwd_weight: too many tiles");
    const long cblocks = C / DT_TC;
    long gy = ((long)kNumSMs * 3 + cblocks - 1) / cblocks;   // ~3 CTAs per SM over all channel blocks
    if (gy > ntiles) gy = ntiles;
    if (gy < 1) gy = 1;
    dim3 grid((unsigned)cblocks, (unsigned)gy);
    dwconv3x3_tile_wgrad_kernel<<<grid, 128, 0, st>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)grad_pre,
                                                       grad_weight, grad_bias, B, H, W, C, dilation, tiles_x, tiles_y);
    RF_CHECK_LAUNCH("dwconv3x3_tile_wgrad_kernel");
    return RF_OK;
  }
  if (dilation == 1) {
    constexpr int PPT = 8, ROWS = 4;
    const int WX = (W + PPT - 1) / PPT, SG = (WX + WG_PL - 1) / WG_PL, RG = (H + ROWS - 1) / ROWS;
    const long gy = (long)B * RG * SG;
    RF_REQUIRE(gy <= 65535, "rf_dwconv3x3_nhwc_bwd_weight: too many tiles (%ld)", gy);
    dim3 grid((unsigned)gx, (unsigned)gy);
    if (dtype == 1)
      dwconv3x3_d1_wgrad_kernel<__nv_bfloat16, PPT, ROWS><<<grid, WG_CG * WG_PL, 0, st>>>(
          (const __nv_bfloat16*)x, (const __nv_bfloat16*)grad_pre, grad_weight, grad_bias, B, H, W, C);
    else
      dwconv3x3_d1_wgrad_kernel<float, PPT, ROWS><<<grid, WG_CG * WG_PL, 0, st>>>(
          (const float*)x, (const float*)grad_pre, grad_weight, grad_bias, B, H, W, C);
    RF_CHECK_LAUNCH("dwconv3x3_d1_wgrad_kernel");
    return RF_OK;
  }
  if (dilation >= 2) {
    const DilGeom geo = dil_geom(H, W, dilation);
    const long nblk = (long)B * geo.yblocks * geo.xblocks;
    long per = nblk * gx / ((long)kNumSMs * 8);   // lattice blocks per CTA: ~8 CTAs per SM
    if (per < 16) per = 16;
    if (per > 512) per = 512;
    const long gy = (nblk + per - 1) / per;
    RF_REQUIRE(gy <= 65535, "rf_dwconv3x3_nhwc_bwd_weight: too many block strips");
    dim3 grid((unsigned)gx, (unsigned)gy);
    if (dtype == 1)
      dwconv3x3_dil_wgrad_kernel<__nv_bfloat16><<<grid, WG_CG * WG_PL, 0, st>>>(
          (const __nv_bfloat16*)x, (const __nv_bfloat16*)grad_pre, grad_weight, grad_bias, B, H, W, C, dilation, (int)per);
    else
      dwconv3x3_dil_wgrad_kernel<float><<<grid, WG_CG * WG_PL, 0, st>>>(
          (const float*)x, (const float*)grad_pre, grad_weight, grad_bias, B, H, W, C, dilation, (int)per);
    RF_CHECK_LAUNCH("dwconv3x3_dil_wgrad_kernel");
    return RF_OK;
  }
  // pixel strips per CTA: ~8 CTAs per SM over the machine, 64..1024 pixels each
  long strip = npix * gx / ((long)kNumSMs * 8);
  if (strip < 64) strip = 64;
  if (strip > 1024) strip = 1024;
  const long gy = (npix + strip - 1) / strip;
  RF_REQUIRE(gy <= 65535, "rf_dwconv3x3_nhwc_bwd_weight: too many pixel strips");
  dim3 grid((unsigned)gx, (unsigned)gy);
  if (dtype == 1)
    dwconv3x3_wgrad_kernel<__nv_bfloat16><<<grid, WG_CG * WG_PL, 0, st>>>(
        (const __nv_bfloat16*)x, (const __nv_bfloat16*)grad_pre, grad_weight, grad_bias, B, H, W, C, dilation,
        (int)strip);
  else
    dwconv3x3_wgrad_kernel<float><<<grid, WG_CG * WG_PL, 0, st>>>((const float*)x, (const float*)grad_pre, grad_weight,
                                                                   grad_bias, B, H, W, C, dilation, (int)strip);
  RF_CHECK_LAUNCH("dwconv3x3_wgrad_kernel");
  return RF_OK;
}
