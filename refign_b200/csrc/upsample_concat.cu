// Fused bilinear up-sampling + channel concatenation of the DAFormer head for sm_100a.
//
// Restates the feature fusion of DAFormerHead.forward (/root/reference/models/heads/daformer.py:203-221):
// every embedded stage feature [B, h_i, w_i, E] is resized to the stage-1 resolution with
// F.interpolate(mode='bilinear', align_corners=False) and the four results are concatenated along the channel
// dimension.  The reference (and the library formulation) runs three up-sampling kernels that each write a
// full-resolution tensor and a concatenation that copies all four again; here one kernel writes the
// concatenated channels-last tensor [B, H, W, sum E] once (the low-resolution sources stay L2-resident), and
// one kernel per source computes its gradient with a gather over the (2s)^2 output pixels that read it
// (no atomics).  HBM-bound: bytes = elt * B * H * W * sum E written (forward) / read (backward).
//
// Source coordinate (ATen area_pixel_compute_source_index, align_corners = false):
//   src = max((dst + 0.5) * in / out - 0.5, 0);  i0 = floor(src);  i1 = min(i0 + 1, in - 1);  l1 = src - i0.
#include <cuda_bf16.h>

#include "rf_common.cuh"

namespace rf {

constexpr int UC_MAXSRC = 4;

struct UcSources {
  const void* ptr[UC_MAXSRC];   // [B, h, w, E] channels-last
  int h[UC_MAXSRC], w[UC_MAXSRC];
  int chunk_begin[UC_MAXSRC + 1];   // first 8-channel chunk of each source inside the concatenated pixel
  int n;
};

__device__ __forceinline__ void uc_unpack(const uint4& u, float (&v)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ uint4 uc_pack(const float (&v)[8]) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  return u;
}

__device__ __forceinline__ void uc_coord(int dst, int in, int out, int& i0, int& i1, float& l1) {
  const float scale = (float)in / (float)out;
  float src = ((float)dst + 0.5f) * scale - 0.5f;
  src = src < 0.f ? 0.f : src;
  i0 = (int)src;
  if (i0 > in - 1) i0 = in - 1;
  i1 = i0 + (i0 < in - 1 ? 1 : 0);
  l1 = src - (float)i0;
}

// one thread = one 16-byte chunk (8 bf16 channels) of one output pixel
__global__ void __launch_bounds__(256)
upsample_concat_fwd_kernel(UcSources S, __nv_bfloat16* __restrict__ y, int B, int H, int W, int chunks) {
  const long idx = blockIdx.x * (long)blockDim.x + threadIdx.x;
  const long total = (long)B * H * W * chunks;
  if (idx >= total) return;
  const int ck = (int)(idx % chunks);
  long p = idx / chunks;
  const int x = (int)(p % W);
  p /= W;
  const int yy = (int)(p % H);
  const int b = (int)(p / H);
  int s = 0;
#pragma unroll
  for (int i = 1; i < UC_MAXSRC; ++i)
    if (i < S.n && ck >= S.chunk_begin[i]) s = i;
  const int lc = ck - S.chunk_begin[s];
  const int E8 = S.chunk_begin[s + 1] - S.chunk_begin[s];
  const int h = S.h[s], w = S.w[s];
  const uint4* src = reinterpret_cast<const uint4*>(S.ptr[s]) + (long)b * h * w * E8 + lc;
  uint4 out;
  if (h == H && w == W) {
    out = __ldg(src + ((long)yy * w + x) * E8);
  } else {
    int y0, y1, x0, x1;
    float ly, lx;
    uc_coord(yy, h, H, y0, y1, ly);
    uc_coord(x, w, W, x0, x1, lx);
    float a[8], bq[8], c[8], d[8], o[8];
    uc_unpack(__ldg(src + ((long)y0 * w + x0) * E8), a);
    uc_unpack(__ldg(src + ((long)y0 * w + x1) * E8), bq);
    uc_unpack(__ldg(src + ((long)y1 * w + x0) * E8), c);
    uc_unpack(__ldg(src + ((long)y1 * w + x1) * E8), d);
    const float hy = 1.f - ly, hx = 1.f - lx;
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = hy * (hx * a[k] + lx * bq[k]) + ly * (hx * c[k] + lx * d[k]);
    out = uc_pack(o);
  }
  *reinterpret_cast<uint4*>(y + idx * 8) = out;
}

// Gradient of one source: thread = one 16-byte chunk of one low-resolution pixel; gathers the output pixels whose
// interpolation window contains it.  gy: [B, H, W, chunks*8]; gsrc: [B, h, w, E8*8].
__global__ void __launch_bounds__(256)
upsample_concat_bwd_kernel(const __nv_bfloat16* __restrict__ gy, __nv_bfloat16* __restrict__ gsrc, int B, int H, int W,
                           int chunks, int chunk_begin, int E8, int h, int w) {
  const long idx = blockIdx.x * (long)blockDim.x + threadIdx.x;
  const long total = (long)B * h * w * E8;
  if (idx >= total) return;
  const int lc = (int)(idx % E8);
  long p = idx / E8;
  const int sx = (int)(p % w);
  p /= w;
  const int sy = (int)(p % h);
  const int b = (int)(p / h);
  const uint4* g = reinterpret_cast<const uint4*>(gy) + (long)b * H * W * chunks + chunk_begin + lc;
  if (h == H && w == W) {
    *reinterpret_cast<uint4*>(gsrc + idx * 8) = __ldg(g + ((long)sy * W + sx) * chunks);
    return;
  }
  // candidate output range: dst with src in (s - 1, s + 1)  =>  dst in ((s - 0.5) * r - 0.5, (s + 1.5) * r - 0.5)
  const float ry = (float)H / (float)h, rx = (float)W / (float)w;
  int ya = (int)floorf(((float)sy - 0.5f) * ry - 0.5f), yb = (int)ceilf(((float)sy + 1.5f) * ry - 0.5f);
  int xa = (int)floorf(((float)sx - 0.5f) * rx - 0.5f), xb = (int)ceilf(((float)sx + 1.5f) * rx - 0.5f);
  ya = ya < 0 ? 0 : ya;
  xa = xa < 0 ? 0 : xa;
  yb = yb > H - 1 ? H - 1 : yb;
  xb = xb > W - 1 ? W - 1 : xb;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int yy = ya; yy <= yb; ++yy) {
    int y0, y1;
    float ly;
    uc_coord(yy, h, H, y0, y1, ly);
    const float wy = (y0 == sy ? 1.f - ly : 0.f) + (y1 == sy ? ly : 0.f);
    if (wy == 0.f) continue;
    for (int xx = xa; xx <= xb; ++xx) {
      int x0, x1;
      float lx;
      uc_coord(xx, w, W, x0, x1, lx);
      const float wx = (x0 == sx ? 1.f - lx : 0.f) + (x1 == sx ? lx : 0.f);
      if (wx == 0.f) continue;
      float v[8];
      uc_unpack(__ldg(g + ((long)yy * W + xx) * chunks), v);
      const float wgt = wy * wx;
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] = fmaf(wgt, v[k], acc[k]);
    }
  }
  *reinterpret_cast<uint4*>(gsrc + idx * 8) = uc_pack(acc);
}

static int uc_check(int n, const int* h, const int* w, const int* E, int B, int H, int W, const char* name) {
  RF_REQUIRE(n >= 1 && n <= UC_MAXSRC, "%s: 1..%d sources", name, UC_MAXSRC);
  RF_REQUIRE(B > 0 && H > 0 && W > 0, "%s: bad output shape", name);
  for (int i = 0; i < n; ++i) {
    RF_REQUIRE(h[i] > 0 && w[i] > 0 && h[i] <= H && w[i] <= W, "%s: source %d is larger than the output", name, i);
    RF_REQUIRE(E[i] > 0 && E[i] % 8 == 0, "%s: source %d has %d channels (must be a multiple of 8)", name, i, E[i]);
  }
  return RF_OK;
}

}  // namespace rf

using namespace rf;

extern "C" int rf_upsample_concat_fwd(const void* const* src, const int* h, const int* w, const int* E, int n, void* y,
                                      int B, int H, int W, void* stream) {
  RF_REQUIRE(src && h && w && E && y, "rf_upsample_concat_fwd: null pointer");
  int rc = uc_check(n, h, w, E, B, H, W, "rf_upsample_concat_fwd");
  if (rc != RF_OK) return rc;
  UcSources S;
  S.n = n;
  int cb = 0;
  for (int i = 0; i < UC_MAXSRC; ++i) {
    S.ptr[i] = i < n ? src[i] : nullptr;
    S.h[i] = i < n ? h[i] : 1;
    S.w[i] = i < n ? w[i] : 1;
    S.chunk_begin[i] = cb;
    if (i < n) {
      RF_REQUIRE(src[i] && ((uintptr_t)src[i] & 15) == 0, "rf_upsample_concat_fwd: source %d null / unaligned", i);
      cb += E[i] / 8;
    }
  }
  S.chunk_begin[UC_MAXSRC] = cb;
  for (int i = n; i <= UC_MAXSRC; ++i) S.chunk_begin[i] = cb;
  RF_REQUIRE(((uintptr_t)y & 15) == 0, "rf_upsample_concat_fwd: output must be 16-byte aligned");
  const long total = (long)B * H * W * cb;
  const long blocks = (total + 255) / 256;
  RF_REQUIRE(blocks < (1l << 31), "rf_upsample_concat_fwd: tensor too large");
  upsample_concat_fwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(S, (__nv_bfloat16*)y, B, H, W, cb);
  RF_CHECK_LAUNCH("upsample_concat_fwd_kernel");
  return RF_OK;
}

extern "C" int rf_upsample_concat_bwd(const void* grad_y, void* const* grad_src, const int* h, const int* w,
                                      const int* E, int n, int B, int H, int W, void* stream) {
  RF_REQUIRE(grad_y && grad_src && h && w && E, "rf_upsample_concat_bwd: null pointer");
  int rc = uc_check(n, h, w, E, B, H, W, "rf_upsample_concat_bwd");
  if (rc != RF_OK) return rc;
  int chunks = 0;
  for (int i = 0; i < n; ++i) chunks += E[i] / 8;
  int cb = 0;
  for (int i = 0; i < n; ++i) {
    const int E8 = E[i] / 8;
    if (grad_src[i] != nullptr) {
      RF_REQUIRE(((uintptr_t)grad_src[i] & 15) == 0, "rf_upsample_concat_bwd: gradient %d unaligned", i);
      const long total = (long)B * h[i] * w[i] * E8;
      const long blocks = (total + 255) / 256;
      upsample_concat_bwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
          (const __nv_bfloat16*)grad_y, (__nv_bfloat16*)grad_src[i], B, H, W, chunks, cb, E8, h[i], w[i]);
      RF_CHECK_LAUNCH("upsample_concat_bwd_kernel");
    }
    cb += E8;
  }
  return RF_OK;
}
