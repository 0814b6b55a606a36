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
