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

template <int K>
__global__ void __launch_bounds__(256)
refine_entropy_kernel(const float* __restrict__ logits, unsigned long long* __restrict__ ent_fix, long HW, int kdyn) {
  const int KK = K > 0 ? K : kdyn;
  const int b = blockIdx.y;
  const float inv_logk = __fdiv_rn(1.0f, exact_logf((float)KK));
  const float* base = logits + (long)b * KK * HW;
  long long local = 0;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < HW; i += (long)gridDim.x * blockDim.x) {
    float v[K > 0 ? K : RFN_MAXK];
    float mx;
#pragma unroll
    for (int k = 0; k < (K > 0 ? K : RFN_MAXK); ++k)
      if (k < KK) {
        v[k] = __ldg(base + (long)k * HW + i);
        mx = (k == 0) ? v[0] : (v[k] > mx ? v[k] : mx);
      }
    float sum = 0.0f;
#pragma unroll
    for (int k = 0; k < (K > 0 ? K : RFN_MAXK); ++k)
      if (k < KK) {
        v[k] = __fsub_rn(v[k], mx);
        sum = __fadd_rn(sum, exact_expf(v[k]));
      }
    const float lse = exact_logf(sum);
    float ent = 0.0f;
#pragma unroll
    for (int k = 0; k < (K > 0 ? K : RFN_MAXK); ++k)
      if (k < KK) {
        const float pk = __fdiv_rn(exact_expf(v[k]), sum);
        const float lp = __fsub_rn(v[k], lse);
        ent = __fsub_rn(ent, __fmul_rn(pk, lp));
      }
    ent = __fmul_rn(ent, inv_logk);
    local += __double2ll_rn((double)ent * RFN_ENT_SCALE);
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

template <int K>
__global__ void __launch_bounds__(256)
refine_mix_kernel(const float* __restrict__ lt, const float* __restrict__ lr, const float* __restrict__ certs,
                  const float* __restrict__ logvar, const uint8_t* __restrict__ wmask,
                  const float* __restrict__ trust, float* __restrict__ probs, long long* __restrict__ label,
                  float* __restrict__ maxprob, long HW, int kdyn, unsigned long long smask, int flags) {
  const int KK = K > 0 ? K : kdyn;
  constexpr int KA = K > 0 ? K : RFN_MAXK;
  const int b = blockIdx.y;
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= HW) return;
  const float* pt = lt + (long)b * KK * HW + i;
  const float* pr = lr + (long)b * KK * HW + i;
  float et[KA], er[KA];
  float mt, mr;
#pragma unroll
  for (int k = 0; k < KA; ++k)
    if (k < KK) {
      et[k] = __ldg(pt + (long)k * HW);
      er[k] = __ldg(pr + (long)k * HW);
      mt = (k == 0) ? et[0] : (et[k] > mt ? et[k] : mt);
      mr = (k == 0) ? er[0] : (er[k] > mr ? er[k] : mr);
    }
  float st = 0.0f, sr = 0.0f;
#pragma unroll
  for (int k = 0; k < KA; ++k)
    if (k < KK) {
      et[k] = exact_expf(__fsub_rn(et[k], mt));
      st = __fadd_rn(st, et[k]);
      er[k] = exact_expf(__fsub_rn(er[k], mr));
      sr = __fadd_rn(sr, er[k]);
    }
  int at = 0, ar = 0;
  float bt = -1.0f, br = -1.0f;
#pragma unroll
  for (int k = 0; k < KA; ++k)
    if (k < KK) {
      et[k] = __fdiv_rn(et[k], st);
      er[k] = __fdiv_rn(er[k], sr);
      if (et[k] > bt) { bt = et[k]; at = k; }
      if (er[k] > br) { br = er[k]; ar = k; }
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
  int best = 0;
  float bestv = -__int_as_float(0x7f800000);
  float* po = probs + (long)b * KK * HW + i;
#pragma unroll
  for (int k = 0; k < KA; ++k)
    if (k < KK) {
      const float Mk = (pairS && ((smask >> k) & 1ull)) ? 1.0f : 0.0f;
      float eps = __fmul_rn(s, (P > Mk ? P : Mk));
      if (!inside) eps = 0.0f;
      const float a = __fmul_rn(__fsub_rn(1.0f, eps), et[k]);
      const float c = __fmul_rn(eps, er[k]);
      const float v = __fadd_rn(a, c);
      po[(long)k * HW] = v;
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
    if (K == 19)
      refine_mix_kernel<19><<<grid, 256, 0, st>>>(logits_trg, logits_ref, certs, logvar, warp_mask, trust, probs_out,
                                                   (long long*)label_out, maxprob_out, HW, K, static_mask, flags);
    else
      refine_mix_kernel<0><<<grid, 256, 0, st>>>(logits_trg, logits_ref, certs, logvar, warp_mask, trust, probs_out,
                                                  (long long*)label_out, maxprob_out, HW, K, static_mask, flags);
    RF_CHECK_LAUNCH("refine_mix_kernel");
  }
  return RF_OK;
}
