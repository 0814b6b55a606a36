// Exact fp32 SR-attention core (forward + backward) for the fp32 PARITY mode of the MiT encoder
// (precision='fp32': every model-level parity test against the reference / the oracle runs in this mode).
//   O = softmax(scale * Q K^T) V, head_dim 64 (mit_b1..b5) or 32 (mit_b0) -- /root/reference/models/backbones/mix_transformer.py:150-160
// The timed bf16 path is the tcgen05 kernel pair (sr_attention.cu, sr_attention_bwd_ws.cu); this file exists so that
// the fp32 mode runs the same fused formulation (no [B, heads, N, N_kv] matrix, q / kv read in place, log-sum-exp
// saved for the backward) through the C-ABI instead of falling back to library batched GEMMs + softmax.  Plain
// fp32 FFMA with correctly-rounded-to-2-ulp expf / logf: errors are summation-order noise (~1e-6 relative).
//
// All three kernels work on 64 x 64 tiles held in shared memory with a 65-float row pitch (conflict-free for both
// the row-major and the transposed access); 256 threads, thread (ty, tx) owns rows {ty + 16 i} x columns {tx + 16 j},
// i, j < 4, so every row of a tile is shared by 16 consecutive lanes (row reductions = 4 shuffles).
//   fwd : CTA = 64 queries, online softmax over 64-key chunks
//   bwd : dQ kernel (CTA = 64 queries, streams keys) and dK/dV kernel (CTA = 64 keys x a split of the queries,
//         fp32 atomics into the zeroed [B, M, 2C] result when there is more than one split); P is recomputed from
//         the saved log-sum-exp, D = rowsum(dO * O) comes from a small pre-pass.
#include <math.h>

#include "rf_common.cuh"

namespace rf {

constexpr int F32_T = 64;            // tile edge
constexpr int F32_P = 65;            // shared-memory row pitch (floats)
constexpr int F32_TILE = F32_T * F32_P;

// rows [row0, row0 + 64) x D channels at channel offset `ch` of a [rows_total, ld] fp32 matrix -> smem tile (zero rows beyond)
template <int D>
__device__ __forceinline__ void f32_load_tile(float* s, const float* __restrict__ g, long ld, int row0, int rows_total, int ch) {
  for (int idx = threadIdx.x; idx < F32_T * (D / 4); idx += 256) {
    const int r = idx / (D / 4), c4 = (idx % (D / 4)) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row0 + r < rows_total) v = __ldg(reinterpret_cast<const float4*>(g + (long)(row0 + r) * ld + ch + c4));
    float* d = s + r * F32_P + c4;
    d[0] = v.x;
    d[1] = v.y;
    d[2] = v.z;
    d[3] = v.w;
  }
}
// acc[i][j] += sum_{k < K} A[ty + 16 i][k] * B[tx + 16 j][k]
template <int K>
__device__ __forceinline__ void f32_tile_nt(float (&acc)[4][4], const float* A, const float* B, int ty, int tx) {
#pragma unroll 8
  for (int k = 0; k < K; ++k) {
    float a[4], b[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = A[(ty + 16 * i) * F32_P + k];
#pragma unroll
    for (int j = 0; j < 4; ++j) b[j] = B[(tx + 16 * j) * F32_P + k];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
  }
}
// acc[i][j] += sum_{k < 64} A[ty + 16 i][k] * B[k][tx + 16 j],  j < NJ (NJ = head_dim / 16 output columns per thread)
template <int NJ>
__device__ __forceinline__ void f32_tile_nn(float (&acc)[4][NJ], const float* A, const float* B, int ty, int tx) {
#pragma unroll 8
  for (int k = 0; k < F32_T; ++k) {
    float a[4], b[NJ];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = A[(ty + 16 * i) * F32_P + k];
#pragma unroll
    for (int j = 0; j < NJ; ++j) b[j] = B[k * F32_P + tx + 16 * j];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < NJ; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
  }
}
__device__ __forceinline__ float f32_row_max(float v) {   // over the 16 lanes that share a row
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float f32_row_sum(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------------------ forward
template <int D>
__global__ void __launch_bounds__(256)
sr_attention_f32_fwd_kernel(const float* __restrict__ q, const float* __restrict__ kv, float* __restrict__ out,
                            float* __restrict__ lse, int N, int M, int heads, float scale) {
  extern __shared__ float sm[];
  float *Qs = sm, *Ks = sm + F32_TILE, *Vs = sm + 2 * F32_TILE, *Ps = sm + 3 * F32_TILE;
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  const int q0 = blockIdx.x * F32_T, head = blockIdx.y, b = blockIdx.z;
  const int C = heads * D;
  const float* qb = q + (long)b * N * C;
  const float* kvb = kv + (long)b * M * 2 * C;
  f32_load_tile<D>(Qs, qb, C, q0, N, head * D);
  constexpr int NJ = D / 16;
  float m[4], l[4], o[4][NJ];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m[i] = -INFINITY;
    l[i] = 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j) o[i][j] = 0.f;
  }
  for (int k0 = 0; k0 < M; k0 += F32_T) {
    __syncthreads();   // previous chunk's Ps / Vs reads are done (and Qs is visible on the first pass)
    f32_load_tile<D>(Ks, kvb, 2 * C, k0, M, head * D);
    f32_load_tile<D>(Vs, kvb, 2 * C, k0, M, C + head * D);
    __syncthreads();
    float s[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
    f32_tile_nt<D>(s, Qs, Ks, ty, tx);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        s[i][j] *= scale;
        if (k0 + tx + 16 * j < M) mx = fmaxf(mx, s[i][j]);
      }
      mx = f32_row_max(mx);
      const float m_new = fmaxf(m[i], mx);          // finite: every chunk holds at least one valid key
      const float alpha = expf(m[i] - m_new);       // 0 on the first chunk
      float rs = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float p = (k0 + tx + 16 * j < M) ? expf(s[i][j] - m_new) : 0.f;
        rs += p;
        Ps[(ty + 16 * i) * F32_P + tx + 16 * j] = p;
      }
      rs = f32_row_sum(rs);
      l[i] = l[i] * alpha + rs;
      m[i] = m_new;
#pragma unroll
      for (int j = 0; j < NJ; ++j) o[i][j] *= alpha;
    }
    __syncthreads();
    f32_tile_nn<NJ>(o, Ps, Vs, ty, tx);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = q0 + ty + 16 * i;
    if (row >= N) continue;
    const float inv = 1.f / l[i];
#pragma unroll
    for (int j = 0; j < NJ; ++j) out[((long)b * N + row) * C + head * D + tx + 16 * j] = o[i][j] * inv;
    if (lse != nullptr && tx == 0) lse[((long)b * heads + head) * N + row] = m[i] + logf(l[i]);
  }
}

// ------------------------------------------------------------------------------------------------ backward: D
__global__ void __launch_bounds__(256)
sr_attention_f32_dvec_kernel(const float* __restrict__ o, const float* __restrict__ dout, float* __restrict__ dvec, int N,
                             int heads, int D, long total) {
  const long t = blockIdx.x * (long)blockDim.x + threadIdx.x;
  const long u = t >> 4;                 // ((b * N + row) * heads + head); 16 lanes x D / 16 channels each
  float d = 0.f;
  if (u < total) {
    const int per = D / 16;              // 4 (head_dim 64) or 2 (head_dim 32) consecutive channels per lane
    const float* po = o + u * D + (t & 15) * per;
    const float* pg = dout + u * D + (t & 15) * per;
    for (int e = 0; e < per; ++e) d = fmaf(__ldg(po + e), __ldg(pg + e), d);
  }
  d = f32_row_sum(d);
  if (u < total && (t & 15) == 0) {
    const int head = (int)(u % heads);
    const long brow = u / heads;
    dvec[((brow / N) * heads + head) * N + brow % N] = d;
  }
}

// ------------------------------------------------------------------------------------------------ backward: dQ
template <int D>
__global__ void __launch_bounds__(256)
sr_attention_f32_bwd_dq_kernel(const float* __restrict__ q, const float* __restrict__ kv, const float* __restrict__ dout,
                               const float* __restrict__ lse, const float* __restrict__ dvec, float* __restrict__ dq, int N,
                               int M, int heads, float scale) {
  extern __shared__ float sm[];
  float *Qs = sm, *dOs = sm + F32_TILE, *Ks = sm + 2 * F32_TILE, *Vs = sm + 3 * F32_TILE, *dSs = sm + 4 * F32_TILE;
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  const int q0 = blockIdx.x * F32_T, head = blockIdx.y, b = blockIdx.z;
  const int C = heads * D;
  const float* kvb = kv + (long)b * M * 2 * C;
  f32_load_tile<D>(Qs, q + (long)b * N * C, C, q0, N, head * D);
  f32_load_tile<D>(dOs, dout + (long)b * N * C, C, q0, N, head * D);
  constexpr int NJ = D / 16;
  float lr[4], dr[4], acc[4][NJ];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = q0 + ty + 16 * i;
    lr[i] = row < N ? __ldg(lse + ((long)b * heads + head) * N + row) : INFINITY;   // P = 0 for rows outside
    dr[i] = row < N ? __ldg(dvec + ((long)b * heads + head) * N + row) : 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j) acc[i][j] = 0.f;
  }
  for (int k0 = 0; k0 < M; k0 += F32_T) {
    __syncthreads();
    f32_load_tile<D>(Ks, kvb, 2 * C, k0, M, head * D);
    f32_load_tile<D>(Vs, kvb, 2 * C, k0, M, C + head * D);
    __syncthreads();
    float s[4][4], dp[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = dp[i][j] = 0.f;
    f32_tile_nt<D>(s, Qs, Ks, ty, tx);
    f32_tile_nt<D>(dp, dOs, Vs, ty, tx);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float p = (k0 + tx + 16 * j < M) ? expf(s[i][j] * scale - lr[i]) : 0.f;
        dSs[(ty + 16 * i) * F32_P + tx + 16 * j] = p * (dp[i][j] - dr[i]) * scale;
      }
    __syncthreads();
    f32_tile_nn<NJ>(acc, dSs, Ks, ty, tx);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = q0 + ty + 16 * i;
    if (row >= N) continue;
#pragma unroll
    for (int j = 0; j < NJ; ++j) dq[((long)b * N + row) * C + head * D + tx + 16 * j] = acc[i][j];
  }
}

// ------------------------------------------------------------------------------------------------ backward: dK, dV
template <int D>
__global__ void __launch_bounds__(256)
sr_attention_f32_bwd_dkv_kernel(const float* __restrict__ q, const float* __restrict__ kv, const float* __restrict__ dout,
                                const float* __restrict__ lse, const float* __restrict__ dvec, float* __restrict__ dkv, int N,
                                int M, int heads, int splits, int chunks_per_split, float scale) {
  extern __shared__ float sm[];
  float *Ks = sm, *Vs = sm + F32_TILE, *Qs = sm + 2 * F32_TILE, *dOs = sm + 3 * F32_TILE, *Pt = sm + 4 * F32_TILE,
        *dSt = sm + 5 * F32_TILE, *lv = sm + 6 * F32_TILE, *dv = lv + F32_T;
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  const int k0 = blockIdx.x * F32_T, head = blockIdx.y;
  const int b = blockIdx.z / splits, split = blockIdx.z % splits;
  const int C = heads * D;
  const float* qb = q + (long)b * N * C;
  const float* dob = dout + (long)b * N * C;
  const float* kvb = kv + (long)b * M * 2 * C;
  f32_load_tile<D>(Ks, kvb, 2 * C, k0, M, head * D);
  f32_load_tile<D>(Vs, kvb, 2 * C, k0, M, C + head * D);
  constexpr int NJ = D / 16;
  float dk[4][NJ], dvv[4][NJ];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < NJ; ++j) dk[i][j] = dvv[i][j] = 0.f;
  const int nchunks = (N + F32_T - 1) / F32_T;
  const int c_begin = split * chunks_per_split, c_end = min(nchunks, c_begin + chunks_per_split);
  for (int ch = c_begin; ch < c_end; ++ch) {
    const int r0 = ch * F32_T;
    __syncthreads();
    f32_load_tile<D>(Qs, qb, C, r0, N, head * D);
    f32_load_tile<D>(dOs, dob, C, r0, N, head * D);
    if (threadIdx.x < F32_T) {
      const int row = r0 + threadIdx.x;
      lv[threadIdx.x] = row < N ? __ldg(lse + ((long)b * heads + head) * N + row) : INFINITY;
      dv[threadIdx.x] = row < N ? __ldg(dvec + ((long)b * heads + head) * N + row) : 0.f;
    }
    __syncthreads();
    float st[4][4], dpt[4][4];   // [key ty + 16 i][query tx + 16 j]
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) st[i][j] = dpt[i][j] = 0.f;
    f32_tile_nt<D>(st, Ks, Qs, ty, tx);
    f32_tile_nt<D>(dpt, Vs, dOs, ty, tx);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int qi = tx + 16 * j;
        const float p = expf(st[i][j] * scale - lv[qi]);
        Pt[(ty + 16 * i) * F32_P + qi] = p;
        dSt[(ty + 16 * i) * F32_P + qi] = p * (dpt[i][j] - dv[qi]) * scale;
      }
    __syncthreads();
    f32_tile_nn<NJ>(dvv, Pt, dOs, ty, tx);
    f32_tile_nn<NJ>(dk, dSt, Qs, ty, tx);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int key = k0 + ty + 16 * i;
    if (key >= M) continue;
    float* dst = dkv + ((long)b * M + key) * 2 * C + head * D;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      if (splits == 1) {
        dst[tx + 16 * j] = dk[i][j];
        dst[C + tx + 16 * j] = dvv[i][j];
      } else {
        atomicAdd(dst + tx + 16 * j, dk[i][j]);
        atomicAdd(dst + C + tx + 16 * j, dvv[i][j]);
      }
    }
  }
}

}  // namespace rf

using namespace rf;

static int f32_check(const char* who, int B, int N, int M, int heads, int head_dim) {
  RF_REQUIRE(B > 0 && N > 0 && M > 0 && heads > 0 && heads <= 65535 && B <= 65535, "%s: bad shape", who);
  RF_REQUIRE(head_dim == 64 || head_dim == 32, "%s: head_dim must be 64 (mit_b1..b5) or 32 (mit_b0), got %d", who, head_dim);
  return RF_OK;
}

extern "C" int rf_sr_attention_f32_fwd(const float* q, const float* kv, float* out, float* lse, int B, int N, int M, int heads,
                                       int head_dim, float scale, void* stream) {
  RF_REQUIRE(q && kv && out, "rf_sr_attention_f32_fwd: null pointer");
  RF_REQUIRE((((uintptr_t)q | (uintptr_t)kv) & 15) == 0, "rf_sr_attention_f32_fwd: q / kv must be 16-byte aligned");
  int rc = f32_check("rf_sr_attention_f32_fwd", B, N, M, heads, head_dim);
  if (rc != RF_OK) return rc;
  const int smem = 4 * F32_TILE * (int)sizeof(float);
  static bool attr = false;
  if (!attr) {
    RF_CUDA(cudaFuncSetAttribute(sr_attention_f32_fwd_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    RF_CUDA(cudaFuncSetAttribute(sr_attention_f32_fwd_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr = true;
  }
  dim3 grid((unsigned)((N + F32_T - 1) / F32_T), (unsigned)heads, (unsigned)B);
  if (head_dim == 64)
    sr_attention_f32_fwd_kernel<64><<<grid, 256, smem, (cudaStream_t)stream>>>(q, kv, out, lse, N, M, heads, scale);
  else
    sr_attention_f32_fwd_kernel<32><<<grid, 256, smem, (cudaStream_t)stream>>>(q, kv, out, lse, N, M, heads, scale);
  RF_CHECK_LAUNCH("sr_attention_f32_fwd_kernel");
  return RF_OK;
}

extern "C" int64_t rf_sr_attention_f32_bwd_workspace_bytes(int B, int N, int heads) {
  return (int64_t)sizeof(float) * B * heads * N;
}

extern "C" int rf_sr_attention_f32_bwd(const float* q, const float* kv, const float* out, const float* grad_out,
                                       const float* lse, float* grad_q, float* grad_kv, void* workspace, int B, int N, int M,
                                       int heads, int head_dim, float scale, void* stream) {
  RF_REQUIRE(q && kv && out && grad_out && lse && grad_q && grad_kv && workspace, "rf_sr_attention_f32_bwd: null pointer");
  RF_REQUIRE((((uintptr_t)q | (uintptr_t)kv | (uintptr_t)out | (uintptr_t)grad_out) & 15) == 0,
             "rf_sr_attention_f32_bwd: operands must be 16-byte aligned");
  int rc = f32_check("rf_sr_attention_f32_bwd", B, N, M, heads, head_dim);
  if (rc != RF_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int C = heads * head_dim;
  float* dvec = (float*)workspace;
  const int smem_q = 5 * F32_TILE * (int)sizeof(float), smem_kv = (6 * F32_TILE + 2 * F32_T) * (int)sizeof(float);
  static bool attr = false;
  if (!attr) {
    RF_CUDA(cudaFuncSetAttribute(sr_attention_f32_bwd_dq_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_q));
    RF_CUDA(cudaFuncSetAttribute(sr_attention_f32_bwd_dkv_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_kv));
    RF_CUDA(cudaFuncSetAttribute(sr_attention_f32_bwd_dq_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_q));
    RF_CUDA(cudaFuncSetAttribute(sr_attention_f32_bwd_dkv_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_kv));
    attr = true;
  }
  {
    const long total = (long)B * N * heads;
    sr_attention_f32_dvec_kernel<<<(unsigned)((total * 16 + 255) / 256), 256, 0, st>>>(out, grad_out, dvec, N, heads, head_dim, total);
    RF_CHECK_LAUNCH("sr_attention_f32_dvec_kernel");
  }
  {
    dim3 grid((unsigned)((N + F32_T - 1) / F32_T), (unsigned)heads, (unsigned)B);
    if (head_dim == 64)
      sr_attention_f32_bwd_dq_kernel<64><<<grid, 256, smem_q, st>>>(q, kv, grad_out, lse, dvec, grad_q, N, M, heads, scale);
    else
      sr_attention_f32_bwd_dq_kernel<32><<<grid, 256, smem_q, st>>>(q, kv, grad_out, lse, dvec, grad_q, N, M, heads, scale);
    RF_CHECK_LAUNCH("sr_attention_f32_bwd_dq_kernel");
  }
  {
    const int kblocks = (M + F32_T - 1) / F32_T, nchunks = (N + F32_T - 1) / F32_T;
    int splits = (int)((2l * kNumSMs + (long)kblocks * heads * B - 1) / ((long)kblocks * heads * B));
    if (splits > nchunks / 4) splits = nchunks / 4;
    if (splits < 1) splits = 1;
    if ((long)B * splits > 65535) splits = 65535 / B;
    const int cps = (nchunks + splits - 1) / splits;
    splits = (nchunks + cps - 1) / cps;
    if (splits > 1) RF_CUDA(cudaMemsetAsync(grad_kv, 0, sizeof(float) * (size_t)B * M * 2 * C, st));
    dim3 grid((unsigned)kblocks, (unsigned)heads, (unsigned)(B * splits));
    if (head_dim == 64)
      sr_attention_f32_bwd_dkv_kernel<64><<<grid, 256, smem_kv, st>>>(q, kv, grad_out, lse, dvec, grad_kv, N, M, heads, splits,
                                                                     cps, scale);
    else
      sr_attention_f32_bwd_dkv_kernel<32><<<grid, 256, smem_kv, st>>>(q, kv, grad_out, lse, dvec, grad_kv, N, M, heads, splits,
                                                                     cps, scale);
    RF_CHECK_LAUNCH("sr_attention_f32_bwd_dkv_kernel");
  }
  return RF_OK;
}
