// Optimiser-side multi-tensor ops on FLAT fp32 buffers.
//
// The reference updates the EMA teacher with a Python loop over ~1086 tensors
// x 3 kernels (/root/reference/models/segmentation_model.py:680-689) and steps
// torch.optim.AdamW over four param groups (:390-419, configs/cityscapes_acdc/
// refign_daformer.yaml:149-153).  Here all trainable parameters, their EMA
// copies, gradients and Adam moments live in single flat buffers (the modules
// hold views), so each update is ONE vectorised, HBM-bound launch:
//   EMA  : 3 x 4 bytes per parameter, AdamW : 7 x 4 bytes per parameter.
#include "rf_common.cuh"

namespace rf {

__global__ void __launch_bounds__(256)
ema_kernel(float* __restrict__ ema, const float* __restrict__ live, long n, float m, float om,
           const float* __restrict__ hyper) {
  if (hyper != nullptr) {  // step-dependent scalars from device memory (CUDA-graph replay)
    m = __ldg(hyper + 10);
    om = __ldg(hyper + 11);
  }
  const long n4 = n >> 2;
  float4* e4 = reinterpret_cast<float4*>(ema);
  const float4* l4 = reinterpret_cast<const float4*>(live);
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
    float4 e = e4[i];
    const float4 l = l4[i];
    // param_m * m + param * (1 - m), same association as the reference
    e.x = __fadd_rn(__fmul_rn(e.x, m), __fmul_rn(l.x, om));
    e.y = __fadd_rn(__fmul_rn(e.y, m), __fmul_rn(l.y, om));
    e.z = __fadd_rn(__fmul_rn(e.z, m), __fmul_rn(l.z, om));
    e.w = __fadd_rn(__fmul_rn(e.w, m), __fmul_rn(l.w, om));
    e4[i] = e;
  }
  for (long i = (n4 << 2) + blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    ema[i] = __fadd_rn(__fmul_rn(ema[i], m), __fmul_rn(live[i], om));
}

constexpr int ADAM_MAX_SEG = 8;
struct AdamSegs {
  long end[ADAM_MAX_SEG];
  float lr[ADAM_MAX_SEG];
  float wd[ADAM_MAX_SEG];
  int n;
};

// torch.optim.AdamW (single-tensor reference semantics):
//   p *= 1 - lr*wd ; m = b1*m + (1-b1)*g ; v = b2*v + (1-b2)*g*g
//   p -= (lr/bc1) * m / (sqrt(v)/sqrt(bc2) + eps)
__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
             long n, AdamSegs segs, float b1, float b2, float eps, float bc1, float sqrt_bc2, float gscale,
             const float* __restrict__ hyper) {
  if (hyper != nullptr) {  // per-segment lr and the bias corrections from device memory (CUDA-graph replay)
#pragma unroll
    for (int k = 0; k < ADAM_MAX_SEG; ++k) segs.lr[k] = __ldg(hyper + k);
    bc1 = __ldg(hyper + 8);
    sqrt_bc2 = __ldg(hyper + 9);
  }
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    int s = 0;
#pragma unroll
    for (int k = 0; k < ADAM_MAX_SEG - 1; ++k) s += (k < segs.n - 1 && i >= segs.end[k]) ? 1 : 0;
    const float lr = segs.lr[s], wd = segs.wd[s];
    const float gi = g[i] * gscale;
    float pi = p[i] * (1.f - lr * wd);
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    const float denom = sqrtf(vi) / sqrt_bc2 + eps;
    pi -= (lr / bc1) * (mi / denom);
    p[i] = pi;
    m[i] = mi;
    v[i] = vi;
  }
}

}  // namespace rf

using namespace rf;

static int ema_launch(float* ema, const float* live, int64_t n, double momentum, const float* hyper, void* stream) {
  RF_REQUIRE(ema && live && n > 0, "rf_ema_update: bad argument");
  RF_REQUIRE((((uintptr_t)ema | (uintptr_t)live) & 15) == 0, "rf_ema_update: buffers must be 16-byte aligned");
  long blocks = ceil_div(n / 4 + 1, 256);
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  ema_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(ema, live, n, (float)momentum, (float)(1.0 - momentum),
                                                            hyper);
  RF_CHECK_LAUNCH("ema_kernel");
  return RF_OK;
}

extern "C" int rf_ema_update(float* ema, const float* live, int64_t n, double momentum, void* stream) {
  return ema_launch(ema, live, n, momentum, nullptr, stream);
}

extern "C" int rf_ema_update_dev(float* ema, const float* live, int64_t n, const float* hyper, void* stream) {
  RF_REQUIRE(hyper != nullptr, "rf_ema_update_dev: null hyper-parameter block");
  return ema_launch(ema, live, n, 0.0, hyper, stream);
}

static int adamw_launch(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, int nseg,
                        const int64_t* seg_end, const float* seg_lr, const float* seg_wd, float beta1, float beta2,
                        float eps, int step, float grad_scale, const float* hyper, void* stream) {
  RF_REQUIRE(param && grad && exp_avg && exp_avg_sq && n > 0, "rf_adamw_step: bad argument");
  RF_REQUIRE(nseg >= 1 && nseg <= ADAM_MAX_SEG && seg_end && seg_lr && seg_wd, "rf_adamw_step: 1..%d segments", ADAM_MAX_SEG);
  RF_REQUIRE(step >= 1, "rf_adamw_step: step counts from 1");
  AdamSegs segs;
  segs.n = nseg;
  for (int i = 0; i < ADAM_MAX_SEG; ++i) {
    segs.end[i] = i < nseg ? seg_end[i] : n;
    segs.lr[i] = i < nseg ? seg_lr[i] : 0.f;
    segs.wd[i] = i < nseg ? seg_wd[i] : 0.f;
  }
  RF_REQUIRE(segs.end[nseg - 1] == n, "rf_adamw_step: last segment must end at n");
  const double bc1 = 1.0 - pow((double)beta1, step), bc2 = 1.0 - pow((double)beta2, step);
  long blocks = ceil_div(n, 256);
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  adamw_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n, segs, beta1, beta2, eps,
                                                              (float)bc1, (float)sqrt(bc2), grad_scale, hyper);
  RF_CHECK_LAUNCH("adamw_kernel");
  return RF_OK;
}

extern "C" int rf_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, int nseg,
                             const int64_t* seg_end, const float* seg_lr, const float* seg_wd, float beta1,
                             float beta2, float eps, int step, float grad_scale, void* stream) {
  return adamw_launch(param, grad, exp_avg, exp_avg_sq, n, nseg, seg_end, seg_lr, seg_wd, beta1, beta2, eps, step,
                      grad_scale, nullptr, stream);
}

extern "C" int rf_adamw_step_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                                 int nseg, const int64_t* seg_end, const float* seg_wd, float beta1, float beta2,
                                 float eps, float grad_scale, const float* hyper, void* stream) {
  RF_REQUIRE(hyper != nullptr, "rf_adamw_step_dev: null hyper-parameter block");
  const float zero[ADAM_MAX_SEG] = {0, 0, 0, 0, 0, 0, 0, 0};
  return adamw_launch(param, grad, exp_avg, exp_avg_sq, n, nseg, seg_end, zero, seg_wd, beta1, beta2, eps, 1,
                      grad_scale, hyper, stream);
}
