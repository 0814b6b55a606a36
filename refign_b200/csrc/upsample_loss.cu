// Loss tail of the train step for sm_100a: bilinear up-sampling of the [B,K,h,w] logits to the label resolution
// fused with the pixel-weighted cross-entropy, and a plain fast up-sampling for the teacher logits.
//
// Restates, for every student forward of DomainAdaptationSegmentationModel.training_step
// (/root/reference/models/segmentation_model.py:160-170,228-240):
//     logits = F.interpolate(logits, size=(H, W), mode='bilinear', align_corners=False)
//     loss   = PixelWeightedCrossEntropyLoss(ignore_index=255)(logits, target, pixel_weight)   (models/losses.py:10-22)
//            = mean over ALL B*H*W pixels of  w_i * (logsumexp_k z_ik - z_i,target_i)   (0 at ignored pixels)
// The library path materialises the fp32 [B,K,H,W] tensor (159 MB at K=19, 1024^2) and passes over it ~8 times
// (up-sample, log-softmax, nll, their backward passes); here the K interpolated logits of a pixel live in
// registers: the forward reads the low-resolution logits (L2-resident) + labels and writes one scalar, the
// backward is a gather per low-resolution pixel over the (2s)^2 full-resolution pixels that read it (their
// softmax is recomputed; no atomics on the gradient).
#include <stdlib.h>

#include "rf_common.cuh"

namespace rf {

constexpr int UL_MAXK = 32;

// ATen area_pixel_compute_source_index (align_corners = false) + clamped neighbours
__device__ __forceinline__ void ul_coord(int dst, int in, float scale, int& i0, int& i1, float& l1) {
  float src = ((float)dst + 0.5f) * scale - 0.5f;
  src = src < 0.f ? 0.f : src;
  i0 = (int)src;
  if (i0 > in - 1) i0 = in - 1;
  i1 = i0 + (i0 < in - 1 ? 1 : 0);
  l1 = src - (float)i0;
}

// interpolated logits of full-resolution pixel (y, x): z[k], and their log-sum-exp
template <int K>
__device__ __forceinline__ float ul_logits(const float* __restrict__ low, int KK, int h, int w, float sy, float sx,
                                           int y, int x, float (&z)[K > 0 ? K : UL_MAXK]) {
  int y0, y1, x0, x1;
  float ly, lx;
  ul_coord(y, h, sy, y0, y1, ly);
  ul_coord(x, w, sx, x0, x1, lx);
  const long plane = (long)h * w;
  const float* p00 = low + (long)y0 * w + x0;
  const float* p01 = low + (long)y0 * w + x1;
  const float* p10 = low + (long)y1 * w + x0;
  const float* p11 = low + (long)y1 * w + x1;
  float mx = -INFINITY;
#pragma unroll
  for (int k = 0; k < (K > 0 ? K : UL_MAXK); ++k)
    if (K > 0 || k < KK) {
      // same association as ATen's kernel: h0 * (w0 a + w1 b) + h1 * (w0 c + w1 d)
      z[k] = (1.f - ly) * ((1.f - lx) * __ldg(p00 + k * plane) + lx * __ldg(p01 + k * plane)) +
             ly * ((1.f - lx) * __ldg(p10 + k * plane) + lx * __ldg(p11 + k * plane));
      mx = fmaxf(mx, z[k]);
    }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < (K > 0 ? K : UL_MAXK); ++k)
    if (K > 0 || k < KK) s += expf(z[k] - mx);
  return mx + logf(s);
}

// softmax of the interpolated logits of full-resolution pixel (y, x) (backward: one exponential per class)
template <int K>
__device__ __forceinline__ void ul_softmax(const float* __restrict__ low, int KK, int h, int w, float sy, float sx,
                                           int y, int x, float (&p)[K > 0 ? K : UL_MAXK]) {
  int y0, y1, x0, x1;
  float ly, lx;
  ul_coord(y, h, sy, y0, y1, ly);
  ul_coord(x, w, sx, x0, x1, lx);
  const long plane = (long)h * w;
  const float* p00 = low + (long)y0 * w + x0;
  const float* p01 = low + (long)y0 * w + x1;
  const float* p10 = low + (long)y1 * w + x0;
  const float* p11 = low + (long)y1 * w + x1;
  float mx = -INFINITY;
#pragma unroll
  for (int k = 0; k < (K > 0 ? K : UL_MAXK); ++k)
    if (K > 0 || k < KK) {
      p[k] = (1.f - ly) * ((1.f - lx) * __ldg(p00 + k * plane) + lx * __ldg(p01 + k * plane)) +
             ly * ((1.f - lx) * __ldg(p10 + k * plane) + lx * __ldg(p11 + k * plane));
      mx = fmaxf(mx, p[k]);
    }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < (K > 0 ? K : UL_MAXK); ++k)
    if (K > 0 || k < KK) {
      p[k] = __expf(p[k] - mx);
      s += p[k];
    }
  const float inv = 1.f / s;
#pragma unroll
  for (int k = 0; k < (K > 0 ? K : UL_MAXK); ++k)
    if (K > 0 || k < KK) p[k] *= inv;
}

template <int K>
__global__ void __launch_bounds__(256)
upsample_ce_fwd_kernel(const float* __restrict__ low, const long long* __restrict__ target,
                       const float* __restrict__ weight, float* __restrict__ loss_sum, int B, int KK, int h, int w,
                       int H, int W, int ignore_index) {
  const long total = (long)B * H * W;
  const float sy = (float)h / (float)H, sx = (float)w / (float)W;
  float local = 0.f;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int x = (int)(idx % W);
    const int y = (int)((idx / W) % H);
    const int b = (int)(idx / ((long)W * H));
    const long long t = target[idx];
    if (t == ignore_index || t < 0 || t >= KK) continue;
    float z[K > 0 ? K : UL_MAXK];
    const float* lowb = low + (long)b * KK * h * w;
    const float lse = ul_logits<K>(lowb, KK, h, w, sy, sx, y, x, z);
    // the target logit is interpolated again from its own plane (same operations, same bits as z[t]): a dynamic
    // index into z[] would push the array to local memory
    float zt;
    {
      int y0, y1, x0, x1;
      float ly, lx;
      ul_coord(y, h, sy, y0, y1, ly);
      ul_coord(x, w, sx, x0, x1, lx);
      const float* pt = lowb + (long)t * h * w;
      zt = (1.f - ly) * ((1.f - lx) * __ldg(pt + (long)y0 * w + x0) + lx * __ldg(pt + (long)y0 * w + x1)) +
           ly * ((1.f - lx) * __ldg(pt + (long)y1 * w + x0) + lx * __ldg(pt + (long)y1 * w + x1));
    }
    const float wi = weight ? __ldg(weight + idx) : 1.f;
    local += wi * (lse - zt);
  }
  local = warp_sum(local);
  __shared__ float part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tsum = 0.f;
    for (int i = 0; i < 8; ++i) tsum += part[i];
    atomicAdd(loss_sum, tsum);
  }
}

// grad_low[b,k,py,px] = g * sum over the full-resolution pixels (y,x) whose bilinear window contains (py,px) of
//   tap(y,x -> py,px) * w_yx * (softmax_k(z_yx) - [k == target_yx])
// FOUR lanes per low-resolution pixel (lane q takes the full-resolution rows ya+q, ya+q+4, ...; the K partial sums
// are combined with two xor-shuffles), so a 1024^2 / 256^2 launch has 524 k threads instead of 131 k.
template <int K>
__global__ void __launch_bounds__(128)
upsample_ce_bwd_kernel(const float* __restrict__ low, const long long* __restrict__ target,
                       const float* __restrict__ weight, const float* __restrict__ grad_loss,
                       float* __restrict__ grad_low, int B, int KK, int h, int w, int H, int W, int ignore_index,
                       float inv_count) {
  constexpr int KR = K > 0 ? K : UL_MAXK;
  const long total = (long)B * h * w;
  const long gidx = blockIdx.x * (long)blockDim.x + threadIdx.x;
  const long idx = gidx >> 2;
  const int q = (int)(gidx & 3);
  const bool active = idx < total;   // whole groups of 4 lanes; inactive lanes still take part in the shuffles
  const long cidx = active ? idx : 0;
  const int px = (int)(cidx % w);
  const int py = (int)((cidx / w) % h);
  const int b = (int)(cidx / ((long)w * h));
  const float sy = (float)h / (float)H, sx = (float)w / (float)W;
  const float ry = (float)H / (float)h, rx = (float)W / (float)w;
  // conservative candidate range (one extra pixel each side: the taps below decide exactly)
  int ya = (int)floorf(((float)py - 0.5f) * ry - 0.5f) - 1, yb = (int)ceilf(((float)py + 1.5f) * ry - 0.5f) + 1;
  int xa = (int)floorf(((float)px - 0.5f) * rx - 0.5f) - 1, xb = (int)ceilf(((float)px + 1.5f) * rx - 0.5f) + 1;
  ya = ya < 0 ? 0 : ya;
  xa = xa < 0 ? 0 : xa;
  yb = yb > H - 1 ? H - 1 : yb;
  xb = xb > W - 1 ? W - 1 : xb;
  const float* lowb = low + (long)b * KK * h * w;
  float acc[KR];
#pragma unroll
  for (int k = 0; k < KR; ++k) acc[k] = 0.f;
  if (active) {
    for (int y = ya + q; y <= yb; y += 4) {
      int y0, y1;
      float ly;
      ul_coord(y, h, sy, y0, y1, ly);
      const float wy = (y0 == py ? 1.f - ly : 0.f) + (y1 == py ? ly : 0.f);
      if (wy == 0.f) continue;
      for (int x = xa; x <= xb; ++x) {
        int x0, x1;
        float lx;
        ul_coord(x, w, sx, x0, x1, lx);
        const float wx = (x0 == px ? 1.f - lx : 0.f) + (x1 == px ? lx : 0.f);
        if (wx == 0.f) continue;
        const long fi = ((long)b * H + y) * W + x;
        const long long t = target[fi];
        if (t == ignore_index || t < 0 || t >= KK) continue;
        float pr[KR];
        ul_softmax<K>(lowb, KK, h, w, sy, sx, y, x, pr);
        const float c = wy * wx * (weight ? __ldg(weight + fi) : 1.f);
#pragma unroll
        for (int k = 0; k < KR; ++k)
          if (K > 0 || k < KK) acc[k] = fmaf(c, pr[k] - (k == (int)t ? 1.f : 0.f), acc[k]);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < KR; ++k) {
    acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], 1);
    acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], 2);
  }
  if (!active || q != 0) return;
  const float g = __ldg(grad_loss) * inv_count;
  float* o = grad_low + (long)b * KK * h * w + (long)py * w + px;
#pragma unroll
  for (int k = 0; k < KR; ++k)
    if (K > 0 || k < KK) o[(long)k * h * w] = acc[k] * g;
}

// Tiled form of the same gradient (round 2): the gather kernel above evaluates the softmax of every full-resolution pixel
// once per low-resolution pixel that reads it -- four times.  Here one CTA owns an 8 x 64 tile of full-resolution pixels:
//   1. every pixel's softmax is evaluated ONCE and  c_yx (softmax_k - [k == target])  goes to shared memory [k][y][x];
//   2. the bilinear adjoint is separable: a reduction along x into the tile's low-resolution columns
//      (R[k][y][j] = sum_x wx(x, j) C[k][y][x]; the x that touch column j form one contiguous run), then along y;
//   3. the tile's partial sums are added to grad_low with red.global (a low-resolution pixel collects from at most four
//      tiles), so grad_low must be zero on entry.
// Needs the up-sampling factor to be >= 2 in both directions (the tile then covers <= 6 x 34 low-resolution pixels).
constexpr int UT_TH = 8, UT_TW = 64, UT_LH = 6, UT_LW = 34;   // 60 KB of shared memory at K = 19: three CTAs per SM

template <int K>
__global__ void __launch_bounds__(256)
upsample_ce_bwd_tile_kernel(const float* __restrict__ low, const long long* __restrict__ target,
                            const float* __restrict__ weight, const float* __restrict__ grad_loss,
                            float* __restrict__ grad_low, int B, int KK, int h, int w, int H, int W, int ignore_index,
                            float inv_count) {
  constexpr int KR = K > 0 ? K : UL_MAXK;
  extern __shared__ float ut_smem[];
  float* C = ut_smem;                                   // [KK][UT_TH][UT_TW]
  float* R = C + (size_t)KK * UT_TH * UT_TW;            // [KK][UT_TH][UT_LW]
  __shared__ int cx0[UT_TW], cx1[UT_TW], ry0[UT_TH], ry1[UT_TH];
  __shared__ float clx[UT_TW], rly[UT_TH];
  __shared__ int xs[UT_LW], xe[UT_LW], ys[UT_LH], ye[UT_LH];
  const int tid = threadIdx.x;
  const int tx0 = blockIdx.x * UT_TW, ty0 = blockIdx.y * UT_TH, b = blockIdx.z;
  const float sy = (float)h / (float)H, sx = (float)w / (float)W;
  // ---- taps of the tile's columns / rows (outside the image: weight 0 through an impossible cell index)
  if (tid < UT_TW) {
    const int x = tx0 + tid;
    int a = -1, c = -1;
    float l = 0.f;
    if (x < W) ul_coord(x, w, sx, a, c, l);
    cx0[tid] = a; cx1[tid] = c; clx[tid] = l;
  } else if (tid < UT_TW + UT_TH) {
    const int i = tid - UT_TW, y = ty0 + i;
    int a = -1, c = -1;
    float l = 0.f;
    if (y < H) ul_coord(y, h, sy, a, c, l);
    ry0[i] = a; ry1[i] = c; rly[i] = l;
  }
  __syncthreads();
  const int pxa = cx0[0], pya = ry0[0];                 // first low-resolution column / row the tile touches
  // the run of tile columns (rows) that touch low-resolution column j (row i); x0 / x1 are non-decreasing in x
  if (tid < UT_LW) {
    int lo = UT_TW, hi = 0;
    for (int x = 0; x < UT_TW; ++x)
      if (cx0[x] == pxa + tid || cx1[x] == pxa + tid) { lo = min(lo, x); hi = x + 1; }
    xs[tid] = lo; xe[tid] = hi;
  } else if (tid >= 64 && tid < 64 + UT_LH) {
    const int i = tid - 64;
    int lo = UT_TH, hi = 0;
    for (int y = 0; y < UT_TH; ++y)
      if (ry0[y] == pya + i || ry1[y] == pya + i) { lo = min(lo, y); hi = y + 1; }
    ys[i] = lo; ye[i] = hi;
  }
  for (int i = tid; i < KK * UT_TH * UT_LW; i += 256) R[i] = 0.f;      // (visible after the barrier that follows step 1)
  // ---- 1. one softmax per pixel -> C
  const float* lowb = low + (long)b * KK * h * w;
  for (int i = tid; i < UT_TH * UT_TW; i += 256) {
    const int yy = i / UT_TW, xx = i % UT_TW;
    const int y = ty0 + yy, x = tx0 + xx;
    bool live = y < H && x < W;
    long long t = 0;
    long fi = 0;
    if (live) {
      fi = ((long)b * H + y) * W + x;
      t = target[fi];
      live = !(t == ignore_index || t < 0 || t >= KK);
    }
    if (live) {
      float pr[KR];
      ul_softmax<K>(lowb, KK, h, w, sy, sx, y, x, pr);
      const float c = weight ? __ldg(weight + fi) : 1.f;
#pragma unroll
      for (int k = 0; k < KR; ++k)
        if (K > 0 || k < KK) C[(k * UT_TH + yy) * UT_TW + xx] = c * (pr[k] - (k == (int)t ? 1.f : 0.f));
    } else {
      for (int k = 0; k < KK; ++k) C[(k * UT_TH + yy) * UT_TW + xx] = 0.f;
    }
  }
  __syncthreads();
  // ---- 2a. along x: one thread per (class, tile row) walks its 64 columns once.  The low-resolution column x0(x) is
  // non-decreasing and advances by at most one per x (factor >= 2), x1 is x0 or x0 + 1: two running sums, flushed to R
  // when x0 moves on -- no atomics, no searches.
  const int ncol = min(UT_LW, w - pxa), nrow = min(UT_LH, h - pya);
  // (four 16-column segments per row keep all 256 threads busy: with one thread per row 5 of the 8 warps walked 64
  //  dependent shared-memory loads while the rest sat at the barrier -- 35 % of the stall samples; the cells at segment
  //  borders are shared, so the flushes are shared-memory adds into the zeroed R)
  for (int it = tid; it < KK * UT_TH * 4; it += 256) {
    const int seg = it & 3, rid = it >> 2;
    const float* row = C + rid * UT_TW;
    float* rrow = R + rid * UT_LW;
    int cur = cx0[seg * 16];
    if (cur < 0) continue;                             // the whole segment is beyond the image
    float a0 = 0.f, a1 = 0.f;
#pragma unroll 4
    for (int x = seg * 16; x < seg * 16 + 16; ++x) {
      const int c0 = cx0[x];
      if (c0 < 0) break;
      if (c0 != cur) {
        atomicAdd(&rrow[cur - pxa], a0);
        a0 = a1;
        a1 = 0.f;
        cur = c0;
      }
      const float c = row[x], l = clx[x];
      a0 = fmaf(1.f - l, c, a0);
      if (cx1[x] == cur) a0 = fmaf(l, c, a0); else a1 = fmaf(l, c, a1);
    }
    atomicAdd(&rrow[cur - pxa], a0);
    if (cur + 1 - pxa < UT_LW) atomicAdd(&rrow[cur + 1 - pxa], a1);
  }
  __syncthreads();
  // ---- 2b. along y (same walk over the 8 tile rows, one thread per (class, low-resolution column)), 3. add to the gradient
  const float g = __ldg(grad_loss) * inv_count;
  for (int it = tid; it < KK * UT_LW; it += 256) {
    const int j = it % UT_LW, k = it / UT_LW;
    if (j >= ncol || xs[j] >= xe[j]) continue;         // this low-resolution column gets nothing from the tile
    float* gcol = grad_low + ((long)b * KK + k) * h * w + pxa + j;
    int cur = pya;
    float a0 = 0.f, a1 = 0.f;
    for (int yy = 0; yy < UT_TH; ++yy) {
      const int r0 = ry0[yy];
      if (r0 < 0) break;
      if (r0 != cur) {
        atomicAdd(gcol + (long)cur * w, a0 * g);
        a0 = a1;
        a1 = 0.f;
        cur = r0;
      }
      const float c = R[(k * UT_TH + yy) * UT_LW + j], l = rly[yy];
      a0 = fmaf(1.f - l, c, a0);
      if (ry1[yy] == cur) a0 = fmaf(l, c, a0); else a1 = fmaf(l, c, a1);
    }
    atomicAdd(gcol + (long)cur * w, a0 * g);
    if (cur + 1 < h && cur + 1 - pya < nrow) atomicAdd(gcol + (long)(cur + 1) * w, a1 * g);
  }
}

// plain bilinear up-sampling of an fp32 NCHW tensor (align_corners = false): one thread per 4 consecutive x and
// FOUR output rows; 3-D grid (x, row group, plane) -- the flat 1-D form spent 344 instructions per thread on 64-bit
// index divisions (issue slots 87 % busy at 23 % of the HBM roof); here the x taps are computed once for four rows.
__global__ void __launch_bounds__(128)
upsample_bilinear_f32_kernel(const float* __restrict__ in, float* __restrict__ out, int h, int w, int H, int W) {
  const int x4 = blockIdx.x * blockDim.x + threadIdx.x;
  if (x4 * 4 >= W) return;
  const int yg = blockIdx.y * 4;
  const long pl = blockIdx.z;
  const float sy = (float)h / (float)H, sx = (float)w / (float)W;
  int x0[4], x1[4];
  float lx[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) ul_coord(x4 * 4 + i, w, sx, x0[i], x1[i], lx[i]);
  const float* plane = in + pl * (long)h * w;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int y = yg + r;
    if (y >= H) break;
    int y0, y1;
    float ly;
    ul_coord(y, h, sy, y0, y1, ly);
    const float* r0 = plane + (long)y0 * w;
    const float* r1 = plane + (long)y1 * w;
    float o[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
      o[i] = (1.f - ly) * ((1.f - lx[i]) * __ldg(r0 + x0[i]) + lx[i] * __ldg(r0 + x1[i])) +
             ly * ((1.f - lx[i]) * __ldg(r1 + x0[i]) + lx[i] * __ldg(r1 + x1[i]));
    st_cs_f4(out + (pl * H + y) * W + x4 * 4, make_float4(o[0], o[1], o[2], o[3]));
  }
}

}  // namespace rf

using namespace rf;

extern "C" int rf_upsample_ce_fwd(const float* logits, const int64_t* target, const float* pixel_weight,
                                  float* loss_sum, int B, int K, int h, int w, int H, int W, int ignore_index,
                                  void* stream) {
  RF_REQUIRE(logits && target && loss_sum, "rf_upsample_ce_fwd: null pointer");
  RF_REQUIRE(B > 0 && K >= 2 && K <= UL_MAXK && h > 0 && w > 0 && H >= h && W >= w, "rf_upsample_ce_fwd: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  RF_CUDA(cudaMemsetAsync(loss_sum, 0, sizeof(float), st));
  const long total = (long)B * H * W;
  long blocks = (total + 255) / 256;
  if (blocks > (long)kNumSMs * 16) blocks = (long)kNumSMs * 16;
  if (K == 19)
    upsample_ce_fwd_kernel<19><<<(unsigned)blocks, 256, 0, st>>>(logits, (const long long*)target, pixel_weight, loss_sum,
                                                                 B, K, h, w, H, W, ignore_index);
  else
    upsample_ce_fwd_kernel<0><<<(unsigned)blocks, 256, 0, st>>>(logits, (const long long*)target, pixel_weight, loss_sum,
                                                                B, K, h, w, H, W, ignore_index);
  RF_CHECK_LAUNCH("upsample_ce_fwd_kernel");
  return RF_OK;
}

extern "C" int rf_upsample_ce_bwd(const float* logits, const int64_t* target, const float* pixel_weight,
                                  const float* grad_loss, float* grad_logits, int B, int K, int h, int w, int H, int W,
                                  int ignore_index, void* stream) {
  RF_REQUIRE(logits && target && grad_loss && grad_logits, "rf_upsample_ce_bwd: null pointer");
  RF_REQUIRE(B > 0 && K >= 2 && K <= UL_MAXK && h > 0 && w > 0 && H >= h && W >= w, "rf_upsample_ce_bwd: bad shape");
  const long total = (long)B * h * w * 4;   // four lanes per low-resolution pixel
  const long blocks = (total + 127) / 128;
  RF_REQUIRE(blocks < (1l << 31), "rf_upsample_ce_bwd: tensor too large");
  const float inv_count = 1.0f / (float)((double)B * H * W);
  cudaStream_t st = (cudaStream_t)stream;
  // tiled scatter form (one softmax per full-resolution pixel) for up-sampling factors >= 2; RF_UPSAMPLE_CE_BWD=gather
  // keeps the gather kernel (A/B switch)
  static const bool tiled_on = [] { const char* e = getenv("RF_UPSAMPLE_CE_BWD"); return !(e && e[0] == 'g'); }();
  const size_t tile_smem = sizeof(float) * (size_t)K * UT_TH * (UT_TW + UT_LW);
  if (tiled_on && H >= 2 * h && W >= 2 * w && tile_smem <= 200 * 1024 && B <= 65535 && (H + UT_TH - 1) / UT_TH <= 65535) {
    static bool attr = false;
    if (!attr) {
      RF_CUDA(cudaFuncSetAttribute(upsample_ce_bwd_tile_kernel<19>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      RF_CUDA(cudaFuncSetAttribute(upsample_ce_bwd_tile_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      attr = true;
    }
    RF_CUDA(cudaMemsetAsync(grad_logits, 0, sizeof(float) * (size_t)B * K * h * w, st));
    dim3 grid((unsigned)((W + UT_TW - 1) / UT_TW), (unsigned)((H + UT_TH - 1) / UT_TH), (unsigned)B);
    if (K == 19)
      upsample_ce_bwd_tile_kernel<19><<<grid, 256, tile_smem, st>>>(logits, (const long long*)target, pixel_weight, grad_loss,
                                                                   grad_logits, B, K, h, w, H, W, ignore_index, inv_count);
    else
      upsample_ce_bwd_tile_kernel<0><<<grid, 256, tile_smem, st>>>(logits, (const long long*)target, pixel_weight, grad_loss,
                                                                  grad_logits, B, K, h, w, H, W, ignore_index, inv_count);
    RF_CHECK_LAUNCH("upsample_ce_bwd_tile_kernel");
    return RF_OK;
  }
  if (K == 19)
    upsample_ce_bwd_kernel<19><<<(unsigned)blocks, 128, 0, st>>>(logits, (const long long*)target, pixel_weight, grad_loss,
                                                                 grad_logits, B, K, h, w, H, W, ignore_index, inv_count);
  else
    upsample_ce_bwd_kernel<0><<<(unsigned)blocks, 128, 0, st>>>(logits, (const long long*)target, pixel_weight, grad_loss,
                                                                grad_logits, B, K, h, w, H, W, ignore_index, inv_count);
  RF_CHECK_LAUNCH("upsample_ce_bwd_kernel");
  return RF_OK;
}

extern "C" int rf_upsample_bilinear_f32(const float* in, float* out, int64_t planes, int h, int w, int H, int W,
                                        void* stream) {
  RF_REQUIRE(in && out && planes > 0 && h > 0 && w > 0 && H >= h && W >= w, "rf_upsample_bilinear_f32: bad argument");
  RF_REQUIRE(W % 4 == 0 && ((uintptr_t)out & 15) == 0, "rf_upsample_bilinear_f32: W %% 4 == 0 and a 16-byte aligned output");
  RF_REQUIRE(planes <= 65535 && (H + 3) / 4 <= 65535, "rf_upsample_bilinear_f32: tensor too large");
  dim3 grid((unsigned)((W / 4 + 127) / 128), (unsigned)((H + 3) / 4), (unsigned)planes);
  upsample_bilinear_f32_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(in, out, h, w, H, W);
  RF_CHECK_LAUNCH("upsample_bilinear_f32_kernel");
  return RF_OK;
}
