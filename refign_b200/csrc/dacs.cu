// Fused DACS strong transform between `refine` and the student's mixed forward
// (reference models/segmentation_model.py:525-582 get_dacs_mix + helpers/dacs_transforms.py:14-112, which runs a
// per-image Python loop of ~40 small torch / kornia launches and two host synchronisations):
//   rf_dacs_count : #pixels whose pseudo-label confidence reaches the threshold (segmentation_model.py:552-556)
//   rf_dacs_mix   : class-mix of (source, target) image / label / weight by the per-image class subset
//                   (dacs_transforms.py:81-112) + kornia 0.5.8 ColorJitter on the de-normalised mixed image
//                   (dacs_transforms.py:42-59): brightness (additive), contrast (multiplicative), saturation and hue
//                   (through HSV) in a per-image random order, all clamped to [0, 1] -- one pass, every pixel once
//   rf_dacs_blur  : kornia GaussianBlur2d (reflect border, separable normalised gaussian; dacs_transforms.py:62-78)
//                   as two 1-D passes, in place
// All random draws arrive in a small per-image DEVICE parameter block, so the whole transform is capturable in a
// CUDA graph and replayable with new draws (flags switch the jitter / blur off without changing the launch sequence).
// The arithmetic mirrors oracle/kornia_058.py operation by operation (fp32).
#include <math.h>

#include "rf_common.cuh"

namespace rf {

// per-image parameter block (floats): see refign_b200/dacs_transforms.py: pack_strong_params
constexpr int DP_STRIDE = 64;
constexpr int DP_JITTER = 0, DP_ORDER = 1, DP_BSHIFT = 5, DP_CONTRAST = 6, DP_SAT = 7, DP_HSHIFT = 8;
constexpr int DP_BLUR = 9, DP_RY = 10, DP_RX = 11, DP_WY = 12, DP_WX = 29;   // half kernels: centre + 16 taps
constexpr int DP_MAXR = 16;

__global__ void __launch_bounds__(256) dacs_count_kernel(const float* __restrict__ prob, long n, float thr,
                                                         unsigned long long* __restrict__ count) {
  long c = 0;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    c += __ldg(prob + i) >= thr ? 1 : 0;
  c = warp_sum_i64(c);
  __shared__ long long part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long t = 0;
    for (int w = 0; w < 8; ++w) t += part[w];
    if (t) atomicAdd(count, (unsigned long long)t);
  }
}

__device__ __forceinline__ float clamp01(float x) { return fminf(fmaxf(x, 0.f), 1.f); }
__device__ __forceinline__ float rem_pos(float a, float b) { return a - b * floorf(a / b); }   // torch.remainder

// kornia.color.rgb_to_hsv (0.5.8): h in [0, 2 pi)
__device__ __forceinline__ void rgb_to_hsv(float r, float g, float b, float& h, float& s, float& v) {
  const float mx = fmaxf(r, fmaxf(g, b)), mn = fminf(r, fminf(g, b));
  float dc = mx - mn;
  v = mx;
  s = dc / (mx + 1e-6f);
  if (dc == 0.f) dc = 1.f;
  const float rc = mx - r, gc = mx - g, bc = mx - b;
  float hh;
  if (r >= g && r >= b) hh = (bc - gc) / dc;
  else if (g >= b) hh = ((rc - bc) + 2.0f * dc) / dc;
  else hh = ((gc - rc) + 4.0f * dc) / dc;
  hh = rem_pos(hh / 6.0f, 1.0f);
  h = 6.2831855f * hh;
}
__device__ __forceinline__ void hsv_to_rgb(float h, float s, float v, float& r, float& g, float& b) {
  const float hn = h / 6.2831855f;
  const float h6 = hn * 6.f;
  const float hi = rem_pos(floorf(h6), 6.f);
  const float f = rem_pos(h6, 6.f) - hi;
  const float p = v * (1.f - s), q = v * (1.f - f * s), t = v * (1.f - (1.f - f) * s);
  switch ((int)hi) {
    case 0: r = v; g = t; b = p; break;
    case 1: r = q; g = v; b = p; break;
    case 2: r = p; g = v; b = t; break;
    case 3: r = p; g = q; b = v; break;
    case 4: r = t; g = p; b = v; break;
    default: r = v; g = p; b = q; break;
  }
}

__global__ void __launch_bounds__(256)
dacs_mix_kernel(const float* __restrict__ src, const float* __restrict__ trg, const long long* __restrict__ gt,
                const long long* __restrict__ pl, const unsigned long long* __restrict__ count,
                const unsigned char* __restrict__ mask, const float* __restrict__ params, float* __restrict__ oimg,
                long long* __restrict__ olbl, float* __restrict__ ow, int B, int H, int W, int ignore_top, int ignore_bottom,
                float numel) {
  const long hw = (long)H * W;
  const float frac = (float)(*count) / numel;
  const float mean[3] = {0.485f, 0.456f, 0.406f}, sd[3] = {0.229f, 0.224f, 0.225f};
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < (long)B * hw; i += (long)gridDim.x * blockDim.x) {
    const int b = (int)(i / hw);
    const long pix = i - b * hw;
    const int y = (int)(pix / W);
    const float* P = params + b * DP_STRIDE;
    const long long g = __ldg(gt + i);
    const bool m = __ldg(mask + i) != 0;                // 1 = take the source pixel (its class is in the image's subset)
    olbl[i] = m ? g : __ldg(pl + i);
    const float pw = (y < ignore_top || y >= H - ignore_bottom) ? 0.f : frac;
    ow[i] = m ? 1.f : pw;
    const float* im = (m ? src : trg) + (long)b * 3 * hw + pix;
    float c[3] = {__ldg(im), __ldg(im + hw), __ldg(im + 2 * hw)};
    if (P[DP_JITTER] != 0.f) {
#pragma unroll
      for (int k = 0; k < 3; ++k) c[k] = c[k] * sd[k] + mean[k];        // denorm (dacs_transforms.py:30-36)
#pragma unroll 1
      for (int t = 0; t < 4; ++t) {
        const int op = (int)P[DP_ORDER + t];
        if (op == 0) {
          const float sft = P[DP_BSHIFT];
#pragma unroll
          for (int k = 0; k < 3; ++k) c[k] = clamp01(c[k] + sft);
        } else if (op == 1) {
          const float f = P[DP_CONTRAST];
#pragma unroll
          for (int k = 0; k < 3; ++k) c[k] = clamp01(c[k] * f);
        } else {
          float h, s, v;
          rgb_to_hsv(c[0], c[1], c[2], h, s, v);
          if (op == 2) s = clamp01(s * P[DP_SAT]);
          else h = fmodf(h + P[DP_HSHIFT], 6.2831855f);
          hsv_to_rgb(h, s, v, c[0], c[1], c[2]);
        }
      }
#pragma unroll
      for (int k = 0; k < 3; ++k) c[k] = (c[k] - mean[k]) / sd[k];      // renorm
    }
    float* o = oimg + (long)b * 3 * hw + pix;
    o[0] = c[0];
    o[hw] = c[1];
    o[2 * hw] = c[2];
  }
}

__device__ __forceinline__ int reflect(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i); }

// one 1-D pass of the separable gaussian over [B*3, H, W] planes; AXIS 0 = along x, 1 = along y
template <int AXIS>
__global__ void __launch_bounds__(256)
dacs_blur_kernel(const float* __restrict__ in, float* __restrict__ out, const float* __restrict__ params, int B, int H, int W) {
  const long hw = (long)H * W;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < (long)B * 3 * hw; i += (long)gridDim.x * blockDim.x) {
    const int b = (int)(i / (3 * hw));
    const float* P = params + b * DP_STRIDE;
    if (P[DP_BLUR] == 0.f) continue;                    // flag off: the image stays as the mix kernel wrote it
    const long pix = i % hw;
    const int y = (int)(pix / W), x = (int)(pix - (long)y * W);
    const float* plane = in + (i - pix);
    const int r = (int)P[AXIS == 0 ? DP_RX : DP_RY];
    const float* w = P + (AXIS == 0 ? DP_WX : DP_WY);
    float acc = 0.f;
    for (int k = -r; k <= r; ++k) {                      // ascending tap order, as a correlation with the full kernel
      const float v = AXIS == 0 ? __ldg(plane + (long)y * W + reflect(x + k, W)) : __ldg(plane + (long)reflect(y + k, H) * W + x);
      acc = fmaf(w[k < 0 ? -k : k], v, acc);
    }
    out[i] = acc;
  }
}

}  // namespace rf

using namespace rf;

extern "C" int rf_dacs_count(const float* prob, int64_t n, float threshold, void* count_u64, void* stream) {
  RF_REQUIRE(prob && count_u64 && n > 0, "rf_dacs_count: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  RF_CUDA(cudaMemsetAsync(count_u64, 0, 8, st));
  const int blocks = (int)(ceil_div(n, 256 * 8) < 4 * kNumSMs ? ceil_div(n, 256 * 8) : 4 * kNumSMs);
  dacs_count_kernel<<<blocks, 256, 0, st>>>(prob, n, threshold, (unsigned long long*)count_u64);
  RF_CHECK_LAUNCH("dacs_count_kernel");
  return RF_OK;
}

extern "C" int rf_dacs_mix(const float* img_src, const float* img_trg, const int64_t* gt_src, const int64_t* pseudo_label,
                           const void* count_u64, const uint8_t* mix_mask, const float* params, float* out_img,
                           int64_t* out_label, float* out_weight, int B, int H, int W, int ignore_top, int ignore_bottom,
                           void* stream) {
  RF_REQUIRE(img_src && img_trg && gt_src && pseudo_label && count_u64 && mix_mask && params && out_img && out_label && out_weight,
             "rf_dacs_mix: null pointer");
  RF_REQUIRE(B > 0 && H > 0 && W > 0 && ignore_top >= 0 && ignore_bottom >= 0, "rf_dacs_mix: bad shape");
  const long n = (long)B * H * W;
  const int blocks = (int)(ceil_div(n, 256) < 8 * kNumSMs ? ceil_div(n, 256) : 8 * kNumSMs);
  dacs_mix_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(img_src, img_trg, (const long long*)gt_src,
                                                            (const long long*)pseudo_label, (const unsigned long long*)count_u64,
                                                            mix_mask, params, out_img, (long long*)out_label, out_weight, B, H, W,
                                                            ignore_top, ignore_bottom, (float)n);
  RF_CHECK_LAUNCH("dacs_mix_kernel");
  return RF_OK;
}

extern "C" int rf_dacs_blur(float* img, float* tmp, const float* params, int B, int H, int W, void* stream) {
  RF_REQUIRE(img && tmp && params && B > 0 && H > DP_MAXR && W > DP_MAXR, "rf_dacs_blur: bad arguments (image sides must exceed %d)", DP_MAXR);
  const long n = (long)B * 3 * H * W;
  const int blocks = (int)(ceil_div(n, 256) < 8 * kNumSMs ? ceil_div(n, 256) : 8 * kNumSMs);
  cudaStream_t st = (cudaStream_t)stream;
  dacs_blur_kernel<0><<<blocks, 256, 0, st>>>(img, tmp, params, B, H, W);
  RF_CHECK_LAUNCH("dacs_blur_kernel<0>");
  dacs_blur_kernel<1><<<blocks, 256, 0, st>>>(tmp, img, params, B, H, W);
  RF_CHECK_LAUNCH("dacs_blur_kernel<1>");
  return RF_OK;
}
