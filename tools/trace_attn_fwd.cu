// Per-phase timeline of the ping-pong attention forward (sr_attention.cu built with -DWS_TRACE).  Not product code.
// usage: trace_attn_fwd B N M heads [block_a block_b]
#define WS_TRACE 1
#include "../refign_b200/csrc/sr_attention.cu"

#include <stdarg.h>
#include <vector>
namespace rf {
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vfprintf(stderr, fmt, ap);
  va_end(ap);
  fputc('\n', stderr);
}
}  // namespace rf
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)
__global__ void fill_bf16(__nv_bfloat16* p, long n, unsigned seed, float amp) {
  long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned h = (unsigned)i * 2654435761u ^ seed;
  h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
  p[i] = __float2bfloat16(amp * ((h & 0xffff) / 32768.f - 1.f));
}
int main(int argc, char** argv) {
  const int B = argc > 1 ? atoi(argv[1]) : 2, N = argc > 2 ? atoi(argv[2]) : 4096, M = argc > 3 ? atoi(argv[3]) : 1024,
            heads = argc > 4 ? atoi(argv[4]) : 5;
  const int C = heads * 64;
  CK(cudaSetDevice(0));
  __nv_bfloat16 *q, *kv, *o;
  float* lse;
  long long* trace;
  const long nq = (long)B * N * C, nkv = (long)B * M * 2 * C;
  CK(cudaMalloc(&q, nq * 2)); CK(cudaMalloc(&o, nq * 2)); CK(cudaMalloc(&kv, nkv * 2));
  CK(cudaMalloc(&lse, (long)B * heads * N * 4));
  CK(cudaMalloc(&trace, 2 * 11 * 256 * 8));
  fill_bf16<<<(nq + 255) / 256, 256>>>(q, nq, 1, 1.f);
  fill_bf16<<<(nkv + 255) / 256, 256>>>(kv, nkv, 4, 1.f);
  CK(cudaDeviceSynchronize());
  const int nblocks = ((N + 255) / 256) * heads * B;
  int blocks[2] = {argc > 5 ? atoi(argv[5]) : 3, argc > 6 ? atoi(argv[6]) : nblocks / 2};
  CK(cudaMemcpyToSymbol(rf::g_ws_trace, &trace, sizeof(trace)));
  CK(cudaMemcpyToSymbol(rf::g_ws_trace_blocks, blocks, sizeof(blocks)));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float ms = 0;
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaMemset(trace, 0, 2 * 11 * 256 * 8));
    CK(cudaEventRecord(e0));
    if (rf_sr_attention_fwd(q, kv, o, lse, B, N, M, heads, 0.125f, 0) != 0) return 1;
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    CK(cudaEventElapsedTime(&ms, e0, e1));
  }
  printf("B %d N %d M %d heads %d: %d blocks; kernel %.1f us (traced build)\n", B, N, M, heads, nblocks, ms * 1e3);
  std::vector<long long> h(2 * 11 * 256);
  CK(cudaMemcpy(h.data(), trace, 2 * 11 * 256 * 8, cudaMemcpyDeviceToHost));
  for (int slot = 0; slot < 2; ++slot) {
    long long t0 = 0;
    for (int w = 0; w < 11; ++w) {
      long long* base = h.data() + (slot * 11 + w) * 256;
      if (base[0] > 0) { long long c = base[1] & 0xffffffffffffll; if (t0 == 0 || c < t0) t0 = c; }
    }
    for (int w = 0; w < 11; ++w) {
      long long* base = h.data() + (slot * 11 + w) * 256;
      if (base[0] == 0) continue;
      printf("block %d warp %d:", blocks[slot], w);
      for (int i = 0; i < (int)base[0]; ++i) printf(" %d@%lld", (int)(base[1 + i] >> 48), (base[1 + i] & 0xffffffffffffll) - t0);
      printf("\n");
    }
  }
  return 0;
}
