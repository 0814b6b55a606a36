// Adaptive label refinement (+ confidence, + pseudo-label) for sm_100a.
//
// Restates DomainAdaptationSegmentationModel.refine / eta
// (/root/reference/models/segmentation_model.py:438-491), the confidence map
// of helpers/matching_utils.py:52-57 and the torch.max of get_dacs_mix
// (segmentation_model.py:551).  The reference runs ~30 elementwise / reduction
// kernels over [B,19,H,W] tensors (re-reading them ~15x) plus a Python loop over
// 8 classes; here it is three launches:
//   1. refine_entropy_kernel : per-pixel softmax entropy of the target logits,
//      summed per image with warp-shuffle + one 64-bit atomic per CTA, in 2^-40
//      fixed point so the sum does not depend on the reduction order;
//   2. refine_trust_kernel   : s_b = mean^gamma (B threads);
//   3. refine_mix_kernel     : both softmaxes, both argmaxes, static-class mask,
//      eps = s*max(P,M) gated by the warp mask, convex mix, max/argmax of the
//      result -> probabilities, int64 pseudo-label, max-probability, one pass.
//
// Bit-exactness: the integer label must equal the CPU oracle's.  All arithmetic
// that feeds the argmax is a fixed sequence of correctly rounded binary32
// operations (__fmul_rn/__fadd_rn/__fdiv_rn/fmaf, no libdevice exp/log, no FMA
// contraction), mirrored operation for operation in oracle/refign_oracle.c.
//
// Layout: class planes are strided by H*W, so one thread per pixel gives fully
// coalesced 128-byte requests per class plane; each thread keeps the 2x19
// probabilities in registers.
#include "rf_common.cuh"

namespace rf {

constexpr int RFN_MAXK = 32;
constexpr double RFN_ENT_SCALE = 1099511627776.0;  // 2^40
// classes {0,1,2,3,4,8,9,10}: the static classes of segmentation_model.py:452-461
constexpr unsigned long long kRefignStaticMask = 0x71Full;

__device__ __forceinline__ float exact_expf(float x) {
  if (!(x > -103.0f)) return 0.0f;
  if (x > 88.72f) return __int_as_float(0x7f800000);
  const float t = __fmul_rn(x, 1.44269504088896341f);
  const float n = rintf(t);
  float r = fmaf(n, -0.693359375f, x);
  r = fmaf(n, 2.12194440e-4f, r);
  float p = 1.9875691500e-4f;
  p = fmaf(p, r, 1.3981999507e-3f);
  p = fmaf(p, r, 8.3334519073e-3f);
  p = fmaf(p, r, 4.1665795894e-2f);
  p = fmaf(p, r, 1.6666665459e-1f);
  p = fmaf(p, r, 5.0000001201e-1f);
  const float r2 = __fmul_rn(r, r);
  float y = fmaf(p, r2, r);
  y = __fadd_rn(y, 1.0f);
  int e = (int)n;
  if (e < -126) {
    y = __fmul_rn(y, __int_as_float((127 - 100) << 23));
    e += 100;
  } else if (e > 127) {
    y = __fmul_rn(y, 2.0f);
    e -= 1;
  }
  return __fmul_rn(y, __int_as_float((e + 127) << 23));
}

__device__ __forceinline__ float exact_logf(float x) {  // normal x > 0
  const unsigned u = __float_as_uint(x);
  int e = (int)(u >> 23) - 126;
  float m = __uint_as_float((u & 0x007fffffu) | 0x3f000000u);
  if (m < 0.707106781186547524f) {
    e -= 1;
    m = __fadd_rn(m, m);
  }
  m = __fsub_rn(m, 1.0f);
  const float z = __fmul_rn(m, m);
  float p = 7.0376836292e-2f;
  p = fmaf(p, m, -1.1514610310e-1f);
  p = fmaf(p, m, 1.1676998740e-1f);
  p = fmaf(p, m, -1.2420140846e-1f);
  p = fmaf(p, m, 1.4249322787e-1f);
  p = fmaf(p, m, -1.6668057665e-1f);
  p = fmaf(p, m, 2.0000714765e-1f);
  p = fmaf(p, m, -2.4999993993e-1f);
  p = fmaf(p, m, 3.3333331174e-1f);
  float y = __fmul_rn(__fmul_rn(m, z), p);
  const float fe = (float)e;
  y = fmaf(fe, -2.12194440e-4f, y);
  y = fmaf(z, -0.5f, y);
  float r = __fadd_rn(m, y);
  r = fmaf(fe, 0.693359375f, r);
  return r;
}


// exp(max(x, -86)) for x <= 0 (softmax arguments after the max subtraction); the clamp (error < 5e-38
// absolute) keeps the result a normal number and 2^n can be applied by an integer add on the exponent field.
// Round-to-nearest-integer uses the 1.5*2^23 magic constant (two FADDs) instead of FRND/F2I, which
// run on the quarter-rate XU pipe; every step is one correctly rounded binary32 operation and
// oracle/refign_oracle.c (orc_exp_nonpos) mirrors it operation for operation.
__device__ __forceinline__ float exact_exp_nonpos(float x) {
  x = fmaxf(x, -86.0f);
  const float t = __fmul_rn(x, 1.44269504088896341f);
  const float m = __fadd_rn(t, 12582912.0f);
  const float n = __fsub_rn(m, 12582912.0f);
  float r = fmaf(n, -0.693359375f, x);
  r = fmaf(n, 2.12194440e-4f, r);
  float p = 1.9875691500e-4f;
  p = fmaf(p, r, 1.3981999507e-3f);
  p = fmaf(p, r, 8.3334519073e-3f);
  p = fmaf(p, r, 4.1665795894e-2f);
  p = fmaf(p, r, 1.6666665459e-1f);
  p = fmaf(p, r, 5.0000001201e-1f);
  const float r2 = __fmul_rn(r, r);
  float y = fmaf(p, r2, r);
  y = __fadd_rn(y, 1.0f);
  return __uint_as_float(__float_as_uint(y) + (__float_as_uint(m) << 23));
}

// a / b with r = RN(1/b) precomputed: q0 = RN(a r), rem = a - q0 b (exact in the FMA), q = RN(q0 + rem r).
// Three FMA-pipe operations per quotient instead of the ~9-instruction IEEE division sequence; with a
// correctly rounded reciprocal this is the correctly rounded quotient (Markstein) for the normal-range
// softmax operands, and the oracle evaluates the identical sequence (orc_div_by).
__device__ __forceinline__ float exact_div_by(float a, float b, float r) {
  const float q0 = __fmul_rn(a, r);
  const float rem = fmaf(-q0, b, a);
  return fmaf(rem, r, q0);
}

// Packed (FFMA2 / FADD2 / FMUL2, sm_100 f32x2) forms of the two helpers above: the refine kernel runs the
// target and the reference softmax in the two lanes of one 64-bit register pair, which halves the
// floating-point instruction count of this issue-bound kernel.  Each lane is the same correctly rounded
// binary32 operation as in the scalar helper, with two re-associations the oracle mirrors: x - max and
// m - magic are formed as fma(max, -1, x) / (m + (-magic)) -- both exact restatements of the subtraction.
__device__ __forceinline__ float2 splat2(float v) { return make_float2(v, v); }

// exp(x - mx) lane-wise for x - mx <= 0, flushed at -86 (see exact_exp_nonpos)
__device__ __forceinline__ float2 exact_exp_nonpos2(float2 x, float2 mx, float2* vout = nullptr) {
  float2 v = __ffma2_rn(mx, splat2(-1.0f), x);           // RN(x - mx)
  v.x = fmaxf(v.x, -86.0f);
  v.y = fmaxf(v.y, -86.0f);
  if (vout) *vout = v;
  const float2 t = __fmul2_rn(v, splat2(1.44269504088896341f));
  const float2 m = __fadd2_rn(t, splat2(12582912.0f));
  const float2 n = __fadd2_rn(m, splat2(-12582912.0f));
  float2 r = __ffma2_rn(n, splat2(-0.693359375f), v);
  r = __ffma2_rn(n, splat2(2.12194440e-4f), r);
  float2 p = splat2(1.9875691500e-4f);
  p = __ffma2_rn(p, r, splat2(1.3981999507e-3f));
  p = __ffma2_rn(p, r, splat2(8.3334519073e-3f));
  p = __ffma2_rn(p, r, splat2(4.1665795894e-2f));
  p = __ffma2_rn(p, r, splat2(1.6666665459e-1f));
  p = __ffma2_rn(p, r, splat2(5.0000001201e-1f));
  const float2 r2 = __fmul2_rn(r, r);
  float2 y = __ffma2_rn(p, r2, r);
  y = __fadd2_rn(y, splat2(1.0f));
  y.x = __uint_as_float(__float_as_uint(y.x) + (__float_as_uint(m.x) << 23));
  y.y = __uint_as_float(__float_as_uint(y.y) + (__float_as_uint(m.y) << 23));
  return y;
}
__device__ __forceinline__ float2 exact_div_by2(float2 a, float2 negb, float2 r) {
  const float2 q0 = __fmul2_rn(a, r);
  const float2 rem = __ffma2_rn(q0, negb, a);             // a - q0 b, exact
  return __ffma2_rn(rem, r, q0);
}

// normalised entropy of one pixel from its softmax statistics: H / log K = (lse - dot / sum) / log K
__device__ __forceinline__ long long entropy_fix(float sum, float dot, float inv_logk) {
  const float lse = exact_logf(sum);
  float ent = __fsub_rn(lse, __fdiv_rn(dot, sum));
  ent = __fmul_rn(ent, inv_logk);
  return __double2ll_rn((double)ent * RFN_ENT_SCALE);
}

// Two horizontally adjacent pixels per thread (one 8-byte load per class plane), one pixel per f32x2 lane.
// H = -sum p_k log p_k = lse - (sum_k e_k v_k) / sum,  v_k = max(x_k - max, -86), e_k = exp(v_k): one
// division per pixel instead of one per class.
template <int K>
__global__ void __launch_bounds__(256)
refine_entropy_kernel(const float* __restrict__ logits, unsigned long long* __restrict__ ent_fix, long HW, int kdyn) {
  const int KK = K > 0 ? K : kdyn;
  constexpr int KA = K > 0 ? K : RFN_MAXK;
  const int b = blockIdx.y;
  const float inv_logk = __fdiv_rn(1.0f, exact_logf((float)KK));
  const float* base = logits + (long)b * KK * HW;
  long long local = 0;
  const bool vec = (HW % 2 == 0) && ((reinterpret_cast<uintptr_t>(base) & 7) == 0);
  const long npair = vec ? HW / 2 : 0;
  for (long j = blockIdx.x * (long)blockDim.x + threadIdx.x; j < npair; j += (long)gridDim.x * blockDim.x) {
    float2 v[KA];
    float2 mx;
#pragma unroll
    for (int k = 0; k < KA; ++k)
      if (k < KK) {
        v[k] = __ldg(reinterpret_cast<const float2*>(base + (long)k * HW) + j);
        mx.x = (k == 0) ? v[0].x : fmaxf(v[k].x, mx.x);
        mx.y = (k == 0) ? v[0].y : fmaxf(v[k].y, mx.y);
      }
    float2 sum = splat2(0.0f), dot = splat2(0.0f);
#pragma unroll
    for (int k = 0; k < KA; ++k)
      if (k < KK) {
        float2 vk;
        const float2 ek = exact_exp_nonpos2(v[k], mx, &vk);
        sum = __fadd2_rn(sum, ek);
        dot = __ffma2_rn(ek, vk, dot);
      }
    local += entropy_fix(sum.x, dot.x, inv_logk);
    local += entropy_fix(sum.y, dot.y, inv_logk);
  }
  // scalar path: odd HW / unaligned planes
  for (long i = 2 * npair + blockIdx.x * (long)blockDim.x + threadIdx.x; i < HW; i += (long)gridDim.x * blockDim.x) {
    float v[KA];
    float mx;
#pragma unroll
    for (int k = 0; k < KA; ++k)
      if (k < KK) {
        v[k] = __ldg(base + (long)k * HW + i);
        mx = (k == 0) ? v[0] : (v[k] > mx ? v[k] : mx);
      }
    float sum = 0.0f, dot = 0.0f;
#pragma unroll
    for (int k = 0; k < KA; ++k)
      if (k < KK) {
        const float vk = fmaxf(__fsub_rn(v[k], mx), -86.0f);
        const float ek = exact_exp_nonpos(vk);
        sum = __fadd_rn(sum, ek);
        dot = fmaf(ek, vk, dot);
      }
    local += entropy_fix(sum, dot, inv_logk);
  }
  local = warp_sum_i64(local);
  __shared__ long long part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long t = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += part[w];
    atomicAdd(ent_fix + b, (unsigned long long)t);
  }
}

__global__ void refine_trust_kernel(const long long* __restrict__ ent_fix, float* __restrict__ trust, int B, long HW,
                                    float gamma) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const double mean = ((double)ent_fix[b] / RFN_ENT_SCALE) / (double)HW;
  trust[b] = (float)pow(mean, (double)gamma);
}

// SMASK != 0: the static-class set is a compile-time constant (the Refign default {0,1,2,3,4,8,9,10}), which
// removes the per-class mask test and weight selects from the unrolled class loops.
template <int K, unsigned long long SMASK>
__global__ void __launch_bounds__(256)
refine_mix_kernel(const float* __restrict__ lt, const float* __restrict__ lr, const float* __restrict__ certs,
                  const float* __restrict__ logvar, const uint8_t* __restrict__ wmask,
                  const float* __restrict__ trust, float* __restrict__ probs, long long* __restrict__ label,
                  float* __restrict__ maxprob, long HW, int kdyn, unsigned long long smask_dyn, int flags) {
  const unsigned long long smask = SMASK != 0 ? SMASK : smask_dyn;
  const int KK = K > 0 ? K : kdyn;
  constexpr int KA = K > 0 ? K : RFN_MAXK;
  const int b = blockIdx.y;
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= HW) return;
  const float* pt = lt + (long)b * KK * HW + i;
  const float* pr = lr + (long)b * KK * HW + i;
  const unsigned hw = (unsigned)HW;  // K * HW < 2^31 (checked by the launcher): 32-bit plane offsets
  float2 e[KA];  // lane x: target, lane y: warped reference
  float2 mx;
#pragma unroll
  for (int k = 0; k < KA; ++k)
    if (k < KK) {
      e[k].x = __ldg(pt + k * hw);
      e[k].y = __ldg(pr + k * hw);
      mx.x = (k == 0) ? e[0].x : fmaxf(e[k].x, mx.x);
      mx.y = (k == 0) ? e[0].y : fmaxf(e[k].y, mx.y);
    }
  float2 sum = splat2(0.0f);
#pragma unroll
  for (int k = 0; k < KA; ++k)
    if (k < KK) {
      e[k] = exact_exp_nonpos2(e[k], mx);
      sum = __fadd2_rn(sum, e[k]);
    }
  const float2 rcp = make_float2(__frcp_rn(sum.x), __frcp_rn(sum.y));
  const float2 nsum = make_float2(-sum.x, -sum.y);
  int at = 0, ar = 0;
  float bt = -1.0f, br = -1.0f;
#pragma unroll
  for (int k = 0; k < KA; ++k)
    if (k < KK) {
      e[k] = exact_div_by2(e[k], nsum, rcp);
      if (e[k].x > bt) { bt = e[k].x; at = k; }
      if (e[k].y > br) { br = e[k].y; ar = k; }
    }
  float P = 0.5f;
  if (!(flags & 2)) {
    if (certs != nullptr) {
      P = __ldg(certs + (long)b * HW + i);
    } else if (logvar != nullptr) {
      const float var = exact_expf(__ldg(logvar + (long)b * HW + i));
      P = __fsub_rn(1.0f, exact_expf(__fdiv_rn(-1.0f, __fmul_rn(2.0f, var))));
    }
  }
  const bool pairS = !(flags & 1) && ((smask >> at) & 1ull) && ((smask >> ar) & 1ull);
  const bool inside = wmask ? (wmask[(long)b * HW + i] != 0) : true;
  const float s = __ldg(trust + b);
  // eps_k = s * max(P, M_k) takes two values per pixel (M_k is 0 or 1): hoist them and their complements
  float eps0 = __fmul_rn(s, (P > 0.0f ? P : 0.0f));
  float eps1 = __fmul_rn(s, (P > 1.0f ? P : 1.0f));
  if (!pairS) eps1 = eps0;
  if (!inside) { eps0 = 0.0f; eps1 = 0.0f; }
  const float2 w0 = make_float2(__fsub_rn(1.0f, eps0), eps0), w1 = make_float2(__fsub_rn(1.0f, eps1), eps1);
  int best = 0;
  float bestv = -__int_as_float(0x7f800000);
  float* po = probs + (long)b * KK * HW + i;
#pragma unroll
  for (int k = 0; k < KA; ++k)
    if (k < KK) {
      const bool inS = (smask >> k) & 1ull;
      const float2 ac = __fmul2_rn(inS ? w1 : w0, e[k]);   // ((1 - eps) p_t, eps p_r)
      const float v = __fadd_rn(ac.x, ac.y);
      __stcs(po + k * hw, v);
      if (v > bestv) { bestv = v; best = k; }
    }
  if (label) label[(long)b * HW + i] = best;
  if (maxprob) maxprob[(long)b * HW + i] = bestv;
}

__global__ void cert_kernel(const float* __restrict__ u, float* __restrict__ out, long n) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const float var = exact_expf(u[i]);
    out[i] = __fsub_rn(1.0f, exact_expf(__fdiv_rn(-1.0f, __fmul_rn(2.0f, var))));
  }
}

}  // namespace rf

using namespace rf;

extern "C" int rf_cert_fwd(const float* logvar, float* cert, int64_t n, void* stream) {
  RF_REQUIRE(logvar && cert && n > 0, "rf_cert_fwd: bad argument");
  long blocks = (n + 255) / 256;
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  cert_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(logvar, cert, n);
  RF_CHECK_LAUNCH("cert_kernel");
  return RF_OK;
}

extern "C" int rf_refine_fwd(const float* logits_trg, const float* logits_ref, const float* certs, const float* logvar,
                             const uint8_t* warp_mask, int64_t* ent_fix, float* trust, float* probs_out,
                             int64_t* label_out, float* maxprob_out, int B, int K, int64_t HW, float gamma,
                             uint64_t static_mask, int flags, void* stream) {
  RF_REQUIRE(logits_trg && logits_ref && ent_fix && trust && probs_out, "rf_refine_fwd: null pointer");
  RF_REQUIRE(B > 0 && B <= 65535 && HW > 0, "rf_refine_fwd: bad shape");
  RF_REQUIRE(K >= 2 && K <= RFN_MAXK, "rf_refine_fwd: K=%d outside [2,%d]", K, RFN_MAXK);
  RF_REQUIRE((int64_t)K * HW < (1ll << 31), "rf_refine_fwd: K*HW=%lld exceeds 2^31", (long long)K * HW);
  cudaStream_t st = (cudaStream_t)stream;
  RF_CUDA(cudaMemsetAsync(ent_fix, 0, sizeof(int64_t) * B, st));
  {
    long bx = (HW + 255) / 256;
    const long cap = (long)kNumSMs * 8 / B + 1;
    if (bx > cap) bx = cap;
    dim3 grid((unsigned)bx, B);
    if (K == 19)
      refine_entropy_kernel<19><<<grid, 256, 0, st>>>(logits_trg, (unsigned long long*)ent_fix, HW, K);
    else
      refine_entropy_kernel<0><<<grid, 256, 0, st>>>(logits_trg, (unsigned long long*)ent_fix, HW, K);
    RF_CHECK_LAUNCH("refine_entropy_kernel");
  }
  refine_trust_kernel<<<(B + 63) / 64, 64, 0, st>>>((const long long*)ent_fix, trust, B, HW, gamma);
  RF_CHECK_LAUNCH("refine_trust_kernel");
  {
    dim3 grid((unsigned)((HW + 255) / 256), B);
    if (K == 19 && static_mask == kRefignStaticMask)
      refine_mix_kernel<19, kRefignStaticMask><<<grid, 256, 0, st>>>(
          logits_trg, logits_ref, certs, logvar, warp_mask, trust, probs_out, (long long*)label_out, maxprob_out, HW, K,
          static_mask, flags);
    else if (K == 19)
      refine_mix_kernel<19, 0><<<grid, 256, 0, st>>>(logits_trg, logits_ref, certs, logvar, warp_mask, trust, probs_out,
                                                      (long long*)label_out, maxprob_out, HW, K, static_mask, flags);
    else
      refine_mix_kernel<0, 0><<<grid, 256, 0, st>>>(logits_trg, logits_ref, certs, logvar, warp_mask, trust, probs_out,
                                                     (long long*)label_out, maxprob_out, HW, K, static_mask, flags);
    RF_CHECK_LAUNCH("refine_mix_kernel");
  }
  return RF_OK;
}
