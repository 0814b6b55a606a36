/*
 * refign_oracle.c -- CPU restatement of the Refign hot-path operators.
 *
 * THIS FILE IS TEST INFRASTRUCTURE.  It is the checker the CUDA kernels in
 * refign_b200/csrc are compared against; nothing under refign_b200/ may link,
 * import or call it.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py load it.
 *
 * Parity pinning: every function here is checked against outputs of the
 * reference's own code (tests/golden/ fixtures, produced by
 * tests/golden/make_golden.py which imports /root/reference and builds its
 * correlation extension) by tests/test_oracle_golden.py.
 *
 * Each function cites the reference file:line it restates
 * (paths relative to the reference repository root).
 *
 * Plain C99 + OpenMP.  Build: see oracle/Makefile (-O2 -ffp-contract=off so
 * that the "exact" section below evaluates the same IEEE-754 operation
 * sequence as the CUDA kernel does with explicit _rn intrinsics).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void orc_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* ------------------------------------------------------------------------
 * Local (windowed) correlation, forward.
 * Restates models/correlation_ops/correlation.cpp:14-42 (per-output patch
 * product) and :80-129 (output geometry, loop nest).  Same summation order
 * (channel outermost, float accumulator) so the result is bit-identical to
 * the reference CPU extension.
 *   in1,in2 : [B,C,H,W]   out : [B,pH,pW,oH,oW]
 * ---------------------------------------------------------------------- */
void orc_local_corr_fwd(const float *in1, const float *in2, float *out,
                        int B, int C, int H, int W,
                        int kH, int kW, int pH, int pW, int padH, int padW,
                        int dilH, int dilW, int dpH, int dpW, int sH, int sW) {
  const int radH = (pH - 1) / 2, radW = (pW - 1) / 2;
  const int oH = (H + 2 * padH - ((kH - 1) * dilH + 1)) / sH + 1;
  const int oW = (W + 2 * padW - ((kW - 1) * dilW + 1)) / sW + 1;
  const long plane = (long)H * W;
  long job;
  const long njobs = (long)B * pH * pW;
#pragma omp parallel for schedule(dynamic, 1)
  for (job = 0; job < njobs; ++job) {
    const int n = (int)(job / (pH * pW));
    const int ph = (int)((job / pW) % pH);
    const int pw = (int)(job % pW);
    const int offy = (ph - radH) * dpH, offx = (pw - radW) * dpW;
    const float *a = in1 + (long)n * C * plane;
    const float *b = in2 + (long)n * C * plane;
    float *o = out + (((long)n * pH + ph) * pW + pw) * oH * oW;
    for (int y = 0; y < oH; ++y) {
      for (int x = 0; x < oW; ++x) {
        float acc = 0.0f;
        const int y0 = y * sH - padH, x0 = x * sW - padW;
        for (int c = 0; c < C; ++c) {
          for (int i = 0; i < kH; ++i) {
            const int ya = y0 + i * dilH, yb = ya + offy;
            if (ya < 0 || ya >= H || yb < 0 || yb >= H) continue;
            for (int j = 0; j < kW; ++j) {
              const int xa = x0 + j * dilW, xb = xa + offx;
              if (xa < 0 || xa >= W || xb < 0 || xb >= W) continue;
              acc += a[c * plane + (long)ya * W + xa] *
                     b[c * plane + (long)yb * W + xb];
            }
          }
        }
        o[(long)y * oW + x] = acc;
      }
    }
  }
}

/* ------------------------------------------------------------------------
 * Local correlation, backward (both inputs).
 * Restates correlation.cpp:45-78 and :131-183: every (ph,pw,y,x,c) term
 * adds gOut*in2 into gIn1 and gOut*in1 into gIn2.  Parallel over (n,c)
 * here (the reference parallelises over n only); the per-element summation
 * order (ph,pw,y,x) is kept.
 * ---------------------------------------------------------------------- */
void orc_local_corr_bwd(const float *in1, const float *in2, const float *gout,
                        float *gin1, float *gin2,
                        int B, int C, int H, int W,
                        int kH, int kW, int pH, int pW, int padH, int padW,
                        int dilH, int dilW, int dpH, int dpW, int sH, int sW) {
  const int radH = (pH - 1) / 2, radW = (pW - 1) / 2;
  const int oH = (H + 2 * padH - ((kH - 1) * dilH + 1)) / sH + 1;
  const int oW = (W + 2 * padW - ((kW - 1) * dilW + 1)) / sW + 1;
  const long plane = (long)H * W;
  memset(gin1, 0, sizeof(float) * (size_t)B * C * plane);
  memset(gin2, 0, sizeof(float) * (size_t)B * C * plane);
  long job;
  const long njobs = (long)B * C;
#pragma omp parallel for schedule(dynamic, 1)
  for (job = 0; job < njobs; ++job) {
    const int n = (int)(job / C);
    const float *a = in1 + job * plane;
    const float *b = in2 + job * plane;
    float *ga = gin1 + job * plane;
    float *gb = gin2 + job * plane;
    for (int ph = 0; ph < pH; ++ph) {
      for (int pw = 0; pw < pW; ++pw) {
        const int offy = (ph - radH) * dpH, offx = (pw - radW) * dpW;
        const float *g = gout + (((long)n * pH + ph) * pW + pw) * oH * oW;
        for (int y = 0; y < oH; ++y) {
          for (int x = 0; x < oW; ++x) {
            const float go = g[(long)y * oW + x];
            const int y0 = y * sH - padH, x0 = x * sW - padW;
            for (int i = 0; i < kH; ++i) {
              const int ya = y0 + i * dilH, yb = ya + offy;
              if (ya < 0 || ya >= H || yb < 0 || yb >= H) continue;
              for (int j = 0; j < kW; ++j) {
                const int xa = x0 + j * dilW, xb = xa + offx;
                if (xa < 0 || xa >= W || xb < 0 || xb >= W) continue;
                gb[(long)yb * W + xb] += go * a[(long)ya * W + xa];
                ga[(long)ya * W + xa] += go * b[(long)yb * W + xb];
              }
            }
          }
        }
      }
    }
  }
}

/* ------------------------------------------------------------------------
 * ReLU + L2 normalisation over the channel dimension of [B,K,HW].
 * Restates models/modules.py:273 and :307
 * (F.normalize(F.relu(corr), p=2, dim=1), eps = 1e-12).
 * ---------------------------------------------------------------------- */
void orc_relu_l2norm(float *x, int B, long K, long HW) {
  long job;
#pragma omp parallel for
  for (job = 0; job < (long)B * HW; ++job) {
    float *p = x + (job / HW) * K * HW + (job % HW);
    float ss = 0.0f;
    for (long k = 0; k < K; ++k) {
      float v = p[k * HW];
      v = v > 0.0f ? v : 0.0f;
      p[k * HW] = v;
      ss += v * v;
    }
    float nrm = sqrtf(ss);
    if (nrm < 1e-12f) nrm = 1e-12f;
    for (long k = 0; k < K; ++k) p[k * HW] = p[k * HW] / nrm;
  }
}

/* ------------------------------------------------------------------------
 * Global correlation volume with mutual matching, ReLU and L2 norm.
 * Restates models/modules.py:294-308 (forward), :310-333 (mutual matching,
 * eps 1e-5, parenthesisation corr*(A*B)) and :362-374 (bmm; source index is
 * row-major over (h_s,w_s)).
 *   src : [B,C,Ns]  trg : [B,C,Nt]   out : [B,Ns,Nt]
 * mode bit0: apply mutual matching; bit1: apply relu+l2norm over Ns.
 * ---------------------------------------------------------------------- */
void orc_global_corr(const float *src, const float *trg, float *out,
                     int B, int C, long Ns, long Nt, int mode) {
  for (int b = 0; b < B; ++b) {
    const float *S = src + (long)b * C * Ns;
    const float *T = trg + (long)b * C * Nt;
    float *O = out + (long)b * Ns * Nt;
    long s;
#pragma omp parallel for
    for (s = 0; s < Ns; ++s) {
      float *row = O + s * Nt;
      for (long t = 0; t < Nt; ++t) row[t] = 0.0f;
      for (int c = 0; c < C; ++c) {
        const float sv = S[(long)c * Ns + s];
        const float *tr = T + (long)c * Nt;
        for (long t = 0; t < Nt; ++t) row[t] += sv * tr[t];
      }
    }
    if (mode & 1) {
      float *maxA = (float *)malloc(sizeof(float) * Ns); /* over target  */
      float *maxB = (float *)malloc(sizeof(float) * Nt); /* over source  */
      for (long t = 0; t < Nt; ++t) maxB[t] = -INFINITY;
      for (s = 0; s < Ns; ++s) {
        float m = -INFINITY;
        for (long t = 0; t < Nt; ++t) {
          const float v = O[s * Nt + t];
          if (v > m) m = v;
          if (v > maxB[t]) maxB[t] = v;
        }
        maxA[s] = m;
      }
#pragma omp parallel for
      for (s = 0; s < Ns; ++s) {
        for (long t = 0; t < Nt; ++t) {
          const float v = O[s * Nt + t];
          const float ra = v / (maxA[s] + 1e-5f);
          const float rb = v / (maxB[t] + 1e-5f);
          O[s * Nt + t] = v * (ra * rb);
        }
      }
      free(maxA);
      free(maxB);
    }
  }
  if (mode & 2) orc_relu_l2norm(out, B, Ns, Nt);
}

/* ------------------------------------------------------------------------
 * Bilinear warp with validity mask.
 * Restates helpers/matching_utils.py:11-49: vgrid = meshgrid + flow,
 * normalised with 2*v/max(W-1,1)-1, sampled with
 * grid_sample(bilinear, zeros, align_corners=True) -- whose coordinate
 * un-normalisation is ((g+1)/2)*(size-1) -- and the strict-inside mask
 * (:45-47) evaluated on the normalised float32 grid.
 * The all-zero-flow early exit (:19-22) is handled by the Python wrapper.
 *   x : [B,C,H,W]  flow : [B,2,H,W]  out : [B,C,H,W]  mask : [B,H,W] (u8)
 * ---------------------------------------------------------------------- */
void orc_warp_bilinear(const float *x, const float *flow, float *out,
                       uint8_t *mask, int B, int C, int H, int W) {
  const long plane = (long)H * W;
  const float dw = (float)(W - 1 > 1 ? W - 1 : 1);
  const float dh = (float)(H - 1 > 1 ? H - 1 : 1);
  long job;
#pragma omp parallel for
  for (job = 0; job < (long)B * H; ++job) {
    const int b = (int)(job / H), yy = (int)(job % H);
    const float *fx = flow + ((long)b * 2 + 0) * plane + (long)yy * W;
    const float *fy = flow + ((long)b * 2 + 1) * plane + (long)yy * W;
    for (int xx = 0; xx < W; ++xx) {
      float gx = (float)xx + fx[xx];
      float gy = (float)yy + fy[xx];
      gx = 2.0f * gx;
      gx = gx / dw;
      gx = gx - 1.0f;
      gy = 2.0f * gy;
      gy = gy / dh;
      gy = gy - 1.0f;
      if (mask)
        mask[(long)b * plane + (long)yy * W + xx] =
            (gx > -1.0f) && (gy > -1.0f) && (gx < 1.0f) && (gy < 1.0f);
      float ix = ((gx + 1.0f) / 2.0f) * (float)(W - 1);
      float iy = ((gy + 1.0f) / 2.0f) * (float)(H - 1);
      const float x0f = floorf(ix), y0f = floorf(iy);
      const float wx1 = ix - x0f, wy1 = iy - y0f;
      const float wx0 = (x0f + 1.0f) - ix, wy0 = (y0f + 1.0f) - iy;
      /* NaN / huge coordinates sample nothing */
      int ok = (ix > -2.0f) && (ix < (float)W + 1.0f) && (iy > -2.0f) &&
               (iy < (float)H + 1.0f);
      const int x0 = ok ? (int)x0f : -5, y0 = ok ? (int)y0f : -5;
      const int x1 = x0 + 1, y1 = y0 + 1;
      const int vx0 = x0 >= 0 && x0 < W, vx1 = x1 >= 0 && x1 < W;
      const int vy0 = y0 >= 0 && y0 < H, vy1 = y1 >= 0 && y1 < H;
      const float w00 = wx0 * wy0, w01 = wx1 * wy0, w10 = wx0 * wy1,
                  w11 = wx1 * wy1;
      for (int c = 0; c < C; ++c) {
        const float *p = x + ((long)b * C + c) * plane;
        float acc = 0.0f;
        if (vy0 && vx0) acc += p[(long)y0 * W + x0] * w00;
        if (vy0 && vx1) acc += p[(long)y0 * W + x1] * w01;
        if (vy1 && vx0) acc += p[(long)y1 * W + x0] * w10;
        if (vy1 && vx1) acc += p[(long)y1 * W + x1] * w11;
        out[((long)b * C + c) * plane + (long)yy * W + xx] = acc;
      }
    }
  }
}

/* confidence map alone (helpers/matching_utils.py:52-57), R = 1 */
static inline float orc_expf(float x);
void orc_cert(const float *logvar, float *cert, long n);

/* ======================================================================
 * "Exact" section: label refinement.
 *
 * The pseudo-label map is an integer output that has to be bit-exact
 * between the CUDA kernel and this checker, so both sides evaluate the SAME
 * sequence of correctly-rounded IEEE-754 binary32 operations (mul, add, div,
 * fma, rint) -- no libm exp/log whose last bit differs between platforms.
 * The kernel mirrors these helpers with __fmul_rn/__fadd_rn/__fdiv_rn/
 * fmaf; this file is built with -ffp-contract=off.
 * The per-image entropy mean is accumulated in 2^-40 fixed point (int64),
 * which makes it independent of summation order.
 * ==================================================================== */
static inline float bits_to_float(uint32_t u) {
  float f;
  memcpy(&f, &u, 4);
  return f;
}
static inline uint32_t float_to_bits(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  return u;
}

/* exp(x), any finite x; Cephes-style range reduction and degree-5
 * polynomial, every step a single rounded operation. */
static inline float orc_expf(float x) {
  if (!(x > -103.0f)) return 0.0f;
  if (x > 88.72f) return INFINITY;
  const float t = x * 1.44269504088896341f;
  const float n = rintf(t);
  float r = __builtin_fmaf(n, -0.693359375f, x);
  r = __builtin_fmaf(n, 2.12194440e-4f, r);
  float p = 1.9875691500e-4f;
  p = __builtin_fmaf(p, r, 1.3981999507e-3f);
  p = __builtin_fmaf(p, r, 8.3334519073e-3f);
  p = __builtin_fmaf(p, r, 4.1665795894e-2f);
  p = __builtin_fmaf(p, r, 1.6666665459e-1f);
  p = __builtin_fmaf(p, r, 5.0000001201e-1f);
  const float r2 = r * r;
  float y = __builtin_fmaf(p, r2, r);
  y = y + 1.0f;
  int e = (int)n;
  if (e < -126) { /* two-step scaling keeps the denormal rounding IEEE */
    y = y * bits_to_float((uint32_t)(127 - 100) << 23);
    e += 100;
  } else if (e > 127) {
    y = y * 2.0f;
    e -= 1;
  }
  return y * bits_to_float((uint32_t)(e + 127) << 23);
}

/* log(x) for normal x > 0 (used on the softmax denominator, 1 <= x <= K). */
static inline float orc_log_pos(float x) {
  uint32_t u = float_to_bits(x);
  int e = (int)(u >> 23) - 126;
  float m = bits_to_float((u & 0x007fffffu) | 0x3f000000u); /* [0.5,1) */
  if (m < 0.707106781186547524f) {
    e -= 1;
    m = m + m;
  }
  m = m - 1.0f;
  const float z = m * m;
  float p = 7.0376836292e-2f;
  p = __builtin_fmaf(p, m, -1.1514610310e-1f);
  p = __builtin_fmaf(p, m, 1.1676998740e-1f);
  p = __builtin_fmaf(p, m, -1.2420140846e-1f);
  p = __builtin_fmaf(p, m, 1.4249322787e-1f);
  p = __builtin_fmaf(p, m, -1.6668057665e-1f);
  p = __builtin_fmaf(p, m, 2.0000714765e-1f);
  p = __builtin_fmaf(p, m, -2.4999993993e-1f);
  p = __builtin_fmaf(p, m, 3.3333331174e-1f);
  float y = (m * z) * p;
  const float fe = (float)e;
  y = __builtin_fmaf(fe, -2.12194440e-4f, y);
  y = __builtin_fmaf(z, -0.5f, y);
  float r = m + y;
  r = __builtin_fmaf(fe, 0.693359375f, r);
  return r;
}


/* exp(max(x, -86)) for x <= 0 (softmax arguments); the clamp keeps the result normal so
 * 2^n is applied on the exponent field; rint via the 1.5*2^23 magic constant.  Mirrors
 * exact_exp_nonpos / exact_exp_nonpos2 in refign_b200/csrc/refine.cu step for step. */
static inline float orc_exp_nonpos(float x) {
  x = fmaxf(x, -86.0f);
  const float t = x * 1.44269504088896341f;
  const float m = t + 12582912.0f;
  const float n = m - 12582912.0f;
  float r = __builtin_fmaf(n, -0.693359375f, x);
  r = __builtin_fmaf(n, 2.12194440e-4f, r);
  float p = 1.9875691500e-4f;
  p = __builtin_fmaf(p, r, 1.3981999507e-3f);
  p = __builtin_fmaf(p, r, 8.3334519073e-3f);
  p = __builtin_fmaf(p, r, 4.1665795894e-2f);
  p = __builtin_fmaf(p, r, 1.6666665459e-1f);
  p = __builtin_fmaf(p, r, 5.0000001201e-1f);
  const float r2 = r * r;
  float y = __builtin_fmaf(p, r2, r);
  y = y + 1.0f;
  return bits_to_float(float_to_bits(y) + (float_to_bits(m) << 23));
}

/* a / b given r = RN(1/b): q0 = RN(a r), rem = fma(-q0, b, a), q = fma(rem, r, q0)
 * (mirrors exact_div_by in refine.cu). */
static inline float orc_div_by(float a, float b, float r) {
  const float q0 = a * r;
  const float rem = __builtin_fmaf(-q0, b, a);
  return __builtin_fmaf(rem, r, q0);
}

#define ORC_MAXK 64
#define ORC_ENT_SCALE 1099511627776.0 /* 2^40 */

/* Pass 1 of refine: per-image sum of normalised entropies in fixed point.
 * Restates models/segmentation_model.py:484-491 (eta) and the spatial mean
 * of :449.  ent_fix[b] = sum_pix round(2^40 * H(p)/log K). */
void orc_refine_entropy(const float *logits_trg, int64_t *ent_fix, int B,
                        int K, long HW) {
  const float inv_logk = 1.0f / orc_log_pos((float)K);
  for (int b = 0; b < B; ++b) {
    int64_t total = 0;
    long i;
#pragma omp parallel for reduction(+ : total)
    for (i = 0; i < HW; ++i) {
      const float *p = logits_trg + (long)b * K * HW + i;
      float mx = p[0];
      for (int k = 1; k < K; ++k) mx = p[k * HW] > mx ? p[k * HW] : mx;
      /* H = lse - (sum_k e_k v_k) / sum with v_k = x_k - max, e_k = exp(v_k) */
      float sum = 0.0f, dot = 0.0f;
      for (int k = 0; k < K; ++k) {
        const float vk = fmaxf(p[k * HW] - mx, -86.0f);
        const float ek = orc_exp_nonpos(vk);
        sum = sum + ek;
        dot = __builtin_fmaf(ek, vk, dot);
      }
      const float lse = orc_log_pos(sum);
      float ent = lse - dot / sum;
      ent = ent * inv_logk;
      total += (int64_t)llrint((double)ent * ORC_ENT_SCALE);
    }
    ent_fix[b] = total;
  }
}

/* trust score s_b = (mean entropy)^gamma  (segmentation_model.py:449) */
float orc_refine_trust(int64_t ent_fix, long HW, float gamma) {
  const double mean = ((double)ent_fix / ORC_ENT_SCALE) / (double)HW;
  return (float)pow(mean, (double)gamma);
}

/* Pass 2 of refine + pseudo-label.
 * Restates models/segmentation_model.py:438-482 (refine): softmaxes,
 * argmaxes (first maximal index), static-class mask M over
 * S={0,1,2,3,4,8,9,10} (:452-461), P broadcast or 0.5 (:466-473),
 * eps = s*max(P,M) zeroed outside the warp mask (:475-479), convex mix
 * (:481); helpers/matching_utils.py:52-57 for the confidence
 * P_R = 1-exp(-1/(2 exp(u))) when `logvar` is given instead of `certs`;
 * and segmentation_model.py:551 (torch.max over classes -> prob, int64 label).
 *   flags bit0: disable_M, bit1: disable_P
 *   certs / logvar / warp_mask may be NULL.  static_mask: bit k set if k in S.
 */
void orc_refine_mix(const float *logits_trg, const float *logits_ref,
                    const float *certs, const float *logvar,
                    const uint8_t *warp_mask, const float *trust,
                    float *probs_out, int64_t *label_out, float *maxprob_out,
                    int B, int K, long HW, uint64_t static_mask, int flags) {
  long job;
#pragma omp parallel for
  for (job = 0; job < (long)B * HW; ++job) {
    const int b = (int)(job / HW);
    const long i = job % HW;
    const float *pt = logits_trg + (long)b * K * HW + i;
    const float *pr = logits_ref + (long)b * K * HW + i;
    float et[ORC_MAXK], er[ORC_MAXK];
    float mt = pt[0], mr = pr[0];
    for (int k = 1; k < K; ++k) {
      mt = pt[k * HW] > mt ? pt[k * HW] : mt;
      mr = pr[k * HW] > mr ? pr[k * HW] : mr;
    }
    float st = 0.0f, sr = 0.0f;
    for (int k = 0; k < K; ++k) {
      et[k] = orc_exp_nonpos(pt[k * HW] - mt);
      st = st + et[k];
      er[k] = orc_exp_nonpos(pr[k * HW] - mr);
      sr = sr + er[k];
    }
    const float rt = 1.0f / st, rr = 1.0f / sr;
    int at = 0, ar = 0;
    float bt = -1.0f, br = -1.0f;
    for (int k = 0; k < K; ++k) {
      et[k] = orc_div_by(et[k], st, rt);
      er[k] = orc_div_by(er[k], sr, rr);
      if (et[k] > bt) { bt = et[k]; at = k; }
      if (er[k] > br) { br = er[k]; ar = k; }
    }
    float P = 0.5f;
    if (!(flags & 2)) {
      if (certs) {
        P = certs[(long)b * HW + i];
      } else if (logvar) {
        /* 1 - exp(-R^2 / (2 exp(u))), R = 1, same operation order as
         * matching_utils.py:55-56, with the exact exp above. */
        const float u = logvar[(long)b * HW + i];
        const float var = orc_expf(u);
        P = 1.0f - orc_expf(-1.0f / (2.0f * var));
      }
    }
    const int pairS = !(flags & 1) && ((static_mask >> at) & 1) &&
                      ((static_mask >> ar) & 1);
    const int inside = warp_mask ? warp_mask[(long)b * HW + i] : 1;
    const float s = trust[b];
    int best = 0;
    float bestv = -INFINITY;
    for (int k = 0; k < K; ++k) {
      const float Mk = (pairS && ((static_mask >> k) & 1)) ? 1.0f : 0.0f;
      float eps = s * (P > Mk ? P : Mk);
      if (!inside) eps = 0.0f;
      const float a = (1.0f - eps) * et[k];
      const float c = eps * er[k];
      const float v = a + c;
      probs_out[(long)b * K * HW + k * HW + i] = v;
      if (v > bestv) { bestv = v; best = k; }
    }
    if (label_out) label_out[(long)b * HW + i] = best;
    if (maxprob_out) maxprob_out[(long)b * HW + i] = bestv;
  }
}

void orc_cert(const float *logvar, float *cert, long n) {
  long i;
#pragma omp parallel for
  for (i = 0; i < n; ++i) {
    const float var = orc_expf(logvar[i]);
    cert[i] = 1.0f - orc_expf(-1.0f / (2.0f * var));
  }
}

/* exported for unit tests of the exact transcendental helpers */
void orc_expf_array(const float *x, float *y, long n) {
  for (long i = 0; i < n; ++i) y[i] = orc_expf(x[i]);
}
void orc_logf_array(const float *x, float *y, long n) {
  for (long i = 0; i < n; ++i) y[i] = orc_log_pos(x[i]);
}

/* ---------------------------------------------------------------------------
 * Loss tail of the student forwards (models/segmentation_model.py:160-170,
 * 228-240 + models/losses.py:10-22): bilinear up-sampling of the logits
 * [B,K,h,w] to the label size (align_corners=False: ATen's
 * area_pixel_compute_source_index, scale = in/out in float, negative source
 * positions clamped to 0) followed by the pixel-weighted cross-entropy with
 * ignore_index, averaged over ALL B*H*W pixels.  Index / lambda arithmetic in
 * float as ATen does, values accumulated in double.  Also returns the
 * gradient with respect to the low-resolution logits (grad may be NULL;
 * weight may be NULL = 1).
 * ------------------------------------------------------------------------- */
static inline void orc_up_coord(int dst, int in, float scale, int *i0, int *i1, float *l1) {
  float src = ((float)dst + 0.5f) * scale - 0.5f;
  if (src < 0.f) src = 0.f;
  int a = (int)src;
  if (a > in - 1) a = in - 1;
  *i0 = a;
  *i1 = a + (a < in - 1 ? 1 : 0);
  *l1 = src - (float)a;
}

double orc_upsample_ce(const float *logits, const int64_t *target, const float *weight, int B, int K, int h, int w,
                       int H, int W, int ignore_index, double *grad) {
  const float sy = (float)h / (float)H, sx = (float)w / (float)W;
  const long plane = (long)h * w;
  const double inv_n = 1.0 / ((double)B * H * W);
  double total = 0.0;
  if (grad) memset(grad, 0, sizeof(double) * (size_t)B * K * plane);
  for (int b = 0; b < B; ++b) {   /* serial: the gradient scatter is not parallelised (test sizes only) */
    const float *lb = logits + (long)b * K * plane;
    double *gb = grad ? grad + (long)b * K * plane : NULL;
    double *z = (double *)malloc(sizeof(double) * K);
    for (int y = 0; y < H; ++y) {
      int y0, y1;
      float ly;
      orc_up_coord(y, h, sy, &y0, &y1, &ly);
      for (int x = 0; x < W; ++x) {
        const long pix = ((long)b * H + y) * W + x;
        const int64_t t = target[pix];
        if (t == ignore_index || t < 0 || t >= K) continue;
        int x0, x1;
        float lx;
        orc_up_coord(x, w, sx, &x0, &x1, &lx);
        const double w00 = (1.0 - ly) * (1.0 - lx), w01 = (1.0 - ly) * lx, w10 = (double)ly * (1.0 - lx), w11 = (double)ly * lx;
        double mx = -INFINITY;
        for (int k = 0; k < K; ++k) {
          const float *p = lb + (long)k * plane;
          z[k] = w00 * p[(long)y0 * w + x0] + w01 * p[(long)y0 * w + x1] + w10 * p[(long)y1 * w + x0] + w11 * p[(long)y1 * w + x1];
          if (z[k] > mx) mx = z[k];
        }
        double s = 0.0;
        for (int k = 0; k < K; ++k) s += exp(z[k] - mx);
        const double lse = mx + log(s);
        const double wi = weight ? (double)weight[pix] : 1.0;
        total += wi * (lse - z[t]);
        if (gb) {
          for (int k = 0; k < K; ++k) {
            const double g = wi * (exp(z[k] - lse) - (k == t ? 1.0 : 0.0)) * inv_n;
            double *q = gb + (long)k * plane;
            q[(long)y0 * w + x0] += w00 * g;
            q[(long)y0 * w + x1] += w01 * g;
            q[(long)y1 * w + x0] += w10 * g;
            q[(long)y1 * w + x1] += w11 * g;
          }
        }
      }
    }
    free(z);
  }
  return total * inv_n;
}
