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

template <typename T, int MODE>
static int launch_dw(const void* x, const float* w, const float* bias, const void* aux, void* y, int B, int H, int W,
                     int C, int dil, int act, cudaStream_t st, const char* name) {
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
                                            int B, int H, int W, int C, int dilation, int dtype, void* stream) {
  int rc = check_dw(x, grad_pre, grad_weight, B, H, W, C, dilation, dtype, "rf_dwconv3x3_nhwc_bwd_weight");
  if (rc != RF_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  RF_CUDA(cudaMemsetAsync(grad_weight, 0, sizeof(float) * 9 * (size_t)C, st));
  if (grad_bias) RF_CUDA(cudaMemsetAsync(grad_bias, 0, sizeof(float) * (size_t)C, st));
  const long npix = (long)B * H * W;
  const int CG = C / DW_VEC;
  const int gx = (CG + WG_CG - 1) / WG_CG;
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
