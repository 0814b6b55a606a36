// Bilinear warp (+ validity mask) for sm_100a.
//
// Restates helpers.matching_utils.warp
// (/root/reference/helpers/matching_utils.py:11-49): the flow is turned into a
// normalised sampling grid (2*v/max(size-1,1) - 1), sampled by
// grid_sample(bilinear, zeros, align_corners=True), and the mask is the strict
// inequality test on the normalised fp32 grid (:45-47).  The reference builds
// meshgrid / repeat / cat / permute temporaries (~12 kernels) and synchronises
// the host on torch.all(flo == 0); here it is one kernel and the zero-flow early
// exit is a device flag.
//
// The coordinate arithmetic uses explicit _rn intrinsics in the same order as
// the reference's tensor ops so the boolean mask is bit-identical to it.
//
// Mapping: one thread per output pixel and per chunk of `cpt` channels (the whole
// channel range when the pixel grid alone fills the machine, so the coordinate
// arithmetic with its four IEEE divisions is done once per pixel); consecutive
// threads = consecutive x, so the flow reads and the output writes are coalesced
// and the four gathers of neighbouring pixels hit the same lines.  Eight channels
// (32 gathers) are in flight per thread.  (A shared-memory staged variant -- bounding box of a tile's
// samples copied with coalesced row loads, gathers from shared memory -- was measured on B200 and was 2-3x
// SLOWER for both smooth and noisy flows: the per-channel-batch barriers serialise the copy latency, while
// the direct gathers of a dense-correspondence flow already hit L1 lines shared by neighbouring pixels.)
#include "rf_common.cuh"

namespace rf {

constexpr int WARP_UNROLL = 8;  // channels in flight per thread

struct WarpCoord {
  float w00, w01, w10, w11;
  int x0, y0;
  bool vx0, vx1, vy0, vy1;
  bool inside;
};

__device__ __forceinline__ WarpCoord warp_coord(float fx, float fy, int xx, int yy, int H, int W) {
  const float dw = (float)max(W - 1, 1), dh = (float)max(H - 1, 1);
  float gx = __fadd_rn((float)xx, fx);
  float gy = __fadd_rn((float)yy, fy);
  gx = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, gx), dw), 1.0f);
  gy = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, gy), dh), 1.0f);
  WarpCoord c;
  c.inside = (gx > -1.0f) && (gy > -1.0f) && (gx < 1.0f) && (gy < 1.0f);
  // grid_sample un-normalisation, align_corners=True: ((g+1)/2)*(size-1)
  const float ix = __fmul_rn(__fdiv_rn(__fadd_rn(gx, 1.0f), 2.0f), (float)(W - 1));
  const float iy = __fmul_rn(__fdiv_rn(__fadd_rn(gy, 1.0f), 2.0f), (float)(H - 1));
  const float x0f = floorf(ix), y0f = floorf(iy);
  const float wx1 = __fsub_rn(ix, x0f), wy1 = __fsub_rn(iy, y0f);
  const float wx0 = __fsub_rn(__fadd_rn(x0f, 1.0f), ix), wy0 = __fsub_rn(__fadd_rn(y0f, 1.0f), iy);
  const bool ok = (ix > -2.0f) && (ix < (float)W + 1.0f) && (iy > -2.0f) && (iy < (float)H + 1.0f);
  c.x0 = ok ? (int)x0f : -5;
  c.y0 = ok ? (int)y0f : -5;
  c.vx0 = c.x0 >= 0 && c.x0 < W;
  c.vx1 = c.x0 + 1 >= 0 && c.x0 + 1 < W;
  c.vy0 = c.y0 >= 0 && c.y0 < H;
  c.vy1 = c.y0 + 1 >= 0 && c.y0 + 1 < H;
  c.w00 = __fmul_rn(wx0, wy0);
  c.w01 = __fmul_rn(wx1, wy0);
  c.w10 = __fmul_rn(wx0, wy1);
  c.w11 = __fmul_rn(wx1, wy1);
  return c;
}

__global__ void __launch_bounds__(256)
warp_bilinear_fwd_kernel(const float* __restrict__ x, const float* __restrict__ flow, float* __restrict__ out,
                         uint8_t* __restrict__ mask, const int32_t* __restrict__ zero_flag, int C, int H, int W,
                         int cpt) {
  const long plane = (long)H * W;
  const long pix = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (pix >= plane) return;
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * cpt;
  const int c1 = min(c0 + cpt, C);
  const float* xb = x + (long)b * C * plane;
  float* ob = out + (long)b * C * plane;
  if (zero_flag != nullptr && *zero_flag != 0) {  // matching_utils.py:19-22
    for (int c = c0; c < c1; ++c) ob[c * plane + pix] = xb[c * plane + pix];
    if (mask != nullptr && c0 == 0) mask[(long)b * plane + pix] = 1;
    return;
  }
  const int yy = pix / W, xx = pix % W;
  const float fx = flow[((long)b * 2 + 0) * plane + pix];
  const float fy = flow[((long)b * 2 + 1) * plane + pix];
  const WarpCoord k = warp_coord(fx, fy, xx, yy, H, W);
  if (mask != nullptr && c0 == 0) mask[(long)b * plane + pix] = k.inside ? 1 : 0;
  const long o00 = (long)k.y0 * W + k.x0;
  const bool v00 = k.vy0 && k.vx0, v01 = k.vy0 && k.vx1, v10 = k.vy1 && k.vx0, v11 = k.vy1 && k.vx1;
  const float* p = xb + (long)c0 * plane + o00;
  float* q = ob + (long)c0 * plane + pix;
  int c = c0;
  for (; c + WARP_UNROLL <= c1; c += WARP_UNROLL) {
    float t00[WARP_UNROLL], t01[WARP_UNROLL], t10[WARP_UNROLL], t11[WARP_UNROLL];
#pragma unroll
    for (int u = 0; u < WARP_UNROLL; ++u) {
      const float* pu = p + u * plane;
      t00[u] = v00 ? __ldg(pu) : 0.f;
      t01[u] = v01 ? __ldg(pu + 1) : 0.f;
      t10[u] = v10 ? __ldg(pu + W) : 0.f;
      t11[u] = v11 ? __ldg(pu + W + 1) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < WARP_UNROLL; ++u) {
      // an invalid corner contributes nothing: the reference's zero padding adds 0 * w = +0, which leaves
      // the running sum unchanged bit for bit (the sum of the earlier terms is never -0 unless all are)
      float acc = 0.f;
      if (v00) acc = __fadd_rn(acc, __fmul_rn(t00[u], k.w00));
      if (v01) acc = __fadd_rn(acc, __fmul_rn(t01[u], k.w01));
      if (v10) acc = __fadd_rn(acc, __fmul_rn(t10[u], k.w10));
      if (v11) acc = __fadd_rn(acc, __fmul_rn(t11[u], k.w11));
      __stcs(q + u * plane, acc);
    }
    p += WARP_UNROLL * plane;
    q += WARP_UNROLL * plane;
  }
  for (; c < c1; ++c) {
    float acc = 0.f;
    if (v00) acc = __fadd_rn(acc, __fmul_rn(__ldg(p), k.w00));
    if (v01) acc = __fadd_rn(acc, __fmul_rn(__ldg(p + 1), k.w01));
    if (v10) acc = __fadd_rn(acc, __fmul_rn(__ldg(p + W), k.w10));
    if (v11) acc = __fadd_rn(acc, __fmul_rn(__ldg(p + W + 1), k.w11));
    __stcs(q, acc);
    p += plane;
    q += plane;
  }
}

// Backward: grad wrt x is a scatter-add of the four bilinear weights; grad wrt
// the flow is d(out)/d(ix) * d(ix)/d(flow_x) with d(ix)/d(flow_x) = 1 (the
// normalise / un-normalise pair cancels for align_corners=True when W > 1).
__global__ void __launch_bounds__(256)
warp_bilinear_bwd_kernel(const float* __restrict__ x, const float* __restrict__ flow, const float* __restrict__ gout,
                         float* __restrict__ gx, float* __restrict__ gflow, const int32_t* __restrict__ zero_flag,
                         int C, int H, int W) {
  const long plane = (long)H * W;
  const long pix = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (pix >= plane) return;
  const int b = blockIdx.z;
  const float* xb = x + (long)b * C * plane;
  const float* gb = gout + (long)b * C * plane;
  float* gxb = gx ? gx + (long)b * C * plane : nullptr;
  if (zero_flag != nullptr && *zero_flag != 0) {
    if (gxb) for (int c = 0; c < C; ++c) atomicAdd(gxb + c * plane + pix, gb[c * plane + pix]);
    if (gflow) { gflow[((long)b * 2 + 0) * plane + pix] = 0.f; gflow[((long)b * 2 + 1) * plane + pix] = 0.f; }
    return;
  }
  const int yy = pix / W, xx = pix % W;
  const float fx = flow[((long)b * 2 + 0) * plane + pix];
  const float fy = flow[((long)b * 2 + 1) * plane + pix];
  const WarpCoord k = warp_coord(fx, fy, xx, yy, H, W);
  // recover the 1-D weights from the products' factors
  const float dw = (float)max(W - 1, 1), dh = (float)max(H - 1, 1);
  float gxn = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, __fadd_rn((float)xx, fx)), dw), 1.0f);
  float gyn = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, __fadd_rn((float)yy, fy)), dh), 1.0f);
  const float ix = __fmul_rn(__fdiv_rn(__fadd_rn(gxn, 1.0f), 2.0f), (float)(W - 1));
  const float iy = __fmul_rn(__fdiv_rn(__fadd_rn(gyn, 1.0f), 2.0f), (float)(H - 1));
  const float x0f = floorf(ix), y0f = floorf(iy);
  const float wx1 = ix - x0f, wy1 = iy - y0f, wx0 = (x0f + 1.f) - ix, wy0 = (y0f + 1.f) - iy;
  const long o00 = (long)k.y0 * W + k.x0;
  const bool v00 = k.vy0 && k.vx0, v01 = k.vy0 && k.vx1, v10 = k.vy1 && k.vx0, v11 = k.vy1 && k.vx1;
  float dix = 0.f, diy = 0.f;
  for (int c = 0; c < C; ++c) {
    const float g = gb[c * plane + pix];
    const float* p = xb + c * plane;
    const float p00 = v00 ? __ldg(p + o00) : 0.f, p01 = v01 ? __ldg(p + o00 + 1) : 0.f;
    const float p10 = v10 ? __ldg(p + o00 + W) : 0.f, p11 = v11 ? __ldg(p + o00 + W + 1) : 0.f;
    if (gxb) {
      float* q = gxb + c * plane;
      if (v00) atomicAdd(q + o00, g * k.w00);
      if (v01) atomicAdd(q + o00 + 1, g * k.w01);
      if (v10) atomicAdd(q + o00 + W, g * k.w10);
      if (v11) atomicAdd(q + o00 + W + 1, g * k.w11);
    }
    dix += g * ((p01 - p00) * wy0 + (p11 - p10) * wy1);
    diy += g * ((p10 - p00) * wx0 + (p11 - p01) * wx1);
  }
  if (gflow) {
    // d ix / d flow_x = (W-1)/max(W-1,1): 1 unless W == 1
    gflow[((long)b * 2 + 0) * plane + pix] = W > 1 ? dix : 0.f;
    gflow[((long)b * 2 + 1) * plane + pix] = H > 1 ? diy : 0.f;
  }
}

__global__ void flow_is_zero_kernel(const float* __restrict__ flow, long n, int32_t* flag) {
  bool nz = false;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    nz |= (flow[i] != 0.0f);
  if (__any_sync(0xffffffffu, nz) && (threadIdx.x & 31) == 0) atomicAnd(flag, 0);
}
__global__ void set_flag_kernel(int32_t* flag, int v) { *flag = v; }

}  // namespace rf

using namespace rf;

extern "C" int rf_flow_is_zero(const float* flow, int64_t n, int32_t* flag, void* stream) {
  RF_REQUIRE(flow && flag && n > 0, "rf_flow_is_zero: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  set_flag_kernel<<<1, 1, 0, st>>>(flag, 1);
  long blocks = (n + 255) / 256;
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  flow_is_zero_kernel<<<(int)blocks, 256, 0, st>>>(flow, n, flag);
  RF_CHECK_LAUNCH("flow_is_zero_kernel");
  return RF_OK;
}

extern "C" int rf_warp_bilinear_fwd(const float* x, const float* flow, float* out, uint8_t* mask,
                                    const int32_t* all_zero_flag, int B, int C, int H, int W, void* stream) {
  RF_REQUIRE(x && flow && out, "rf_warp_bilinear_fwd: null pointer");
  RF_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, "rf_warp_bilinear_fwd: empty tensor");
  RF_REQUIRE(B <= 65535, "rf_warp_bilinear_fwd: batch too large");
  const long plane = (long)H * W;
  const long pblocks = (plane + 255) / 256;
  // channel chunks: as few as keep >= 8 CTAs per SM in flight (coordinates are recomputed per chunk)
  long chunks = ((long)kNumSMs * 8 + pblocks * B - 1) / (pblocks * B);
  if (chunks < 1) chunks = 1;
  int cpt = (int)((C + chunks - 1) / chunks);
  cpt = ((cpt + WARP_UNROLL - 1) / WARP_UNROLL) * WARP_UNROLL;
  if (cpt > C) cpt = C;
  dim3 grid((unsigned)pblocks, (unsigned)((C + cpt - 1) / cpt), B);
  warp_bilinear_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, flow, out, mask, all_zero_flag, C, H, W, cpt);
  RF_CHECK_LAUNCH("warp_bilinear_fwd_kernel");
  return RF_OK;
}

extern "C" int rf_warp_bilinear_bwd(const float* x, const float* flow, const float* grad_out, float* grad_x,
                                    float* grad_flow, const int32_t* all_zero_flag, int B, int C, int H, int W,
                                    void* stream) {
  RF_REQUIRE(x && flow && grad_out, "rf_warp_bilinear_bwd: null pointer");
  RF_REQUIRE(grad_x || grad_flow, "rf_warp_bilinear_bwd: no gradient requested");
  RF_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && B <= 65535, "rf_warp_bilinear_bwd: bad shape");
  const long plane = (long)H * W;
  dim3 grid((unsigned)((plane + 255) / 256), 1, B);
  warp_bilinear_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, flow, grad_out, grad_x, grad_flow, all_zero_flag, C, H, W);
  RF_CHECK_LAUNCH("warp_bilinear_bwd_kernel");
  return RF_OK;
}
