// Per-phase timeline of the tcgen05 GEMM (gemm_bf16.cu built with -DWS_TRACE).  Not product code.
// usage: trace_gemm M N K [a_mn b_mn out_f32 accumulate] [block_a block_b]
#define WS_TRACE 1
#include "../refign_b200/csrc/gemm_bf16.cu"
#include <stdarg.h>
#include <vector>
namespace rf {
void set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); fputc('\n', stderr); }
}
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)
__global__ void fill_bf16(__nv_bfloat16* p, long n, unsigned seed) {
  long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned h = (unsigned)i * 2654435761u ^ seed; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
  p[i] = __float2bfloat16(((h & 0xffff) / 32768.f - 1.f) * 0.1f);
}
int main(int argc, char** argv) {
  const int M = argc > 1 ? atoi(argv[1]) : 8192, N = argc > 2 ? atoi(argv[2]) : 1280, K = argc > 3 ? atoi(argv[3]) : 320;
  const int amn = argc > 4 ? atoi(argv[4]) : 0, bmn = argc > 5 ? atoi(argv[5]) : 0, of32 = argc > 6 ? atoi(argv[6]) : 0, acc = argc > 7 ? atoi(argv[7]) : 0;
  CK(cudaSetDevice(0));
  __nv_bfloat16 *a, *b; void* out; float* bias; long long* trace;
  CK(cudaMalloc(&a, (long)M * K * 2)); CK(cudaMalloc(&b, (long)N * K * 2)); CK(cudaMalloc(&out, (long)M * N * 4));
  CK(cudaMalloc(&bias, N * 4)); CK(cudaMemset(bias, 0, N * 4)); CK(cudaMemset(out, 0, (long)M * N * 4));
  CK(cudaMalloc(&trace, 2 * 11 * 256 * 8));
  fill_bf16<<<((long)M * K + 255) / 256, 256>>>(a, (long)M * K, 1);
  fill_bf16<<<((long)N * K + 255) / 256, 256>>>(b, (long)N * K, 2);
  CK(cudaDeviceSynchronize());
  int blocks[2] = {argc > 8 ? atoi(argv[8]) : 0, argc > 9 ? atoi(argv[9]) : 100};
  CK(cudaMemcpyToSymbol(rf::g_ws_trace, &trace, sizeof(trace)));
  CK(cudaMemcpyToSymbol(rf::g_ws_trace_blocks, blocks, sizeof(blocks)));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float ms = 0;
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaMemset(trace, 0, 2 * 11 * 256 * 8));
    CK(cudaEventRecord(e0));
    if (rf_gemm_bf16(a, b, (acc || of32) ? nullptr : bias, out, M, N, K, amn, bmn, of32, acc, nullptr, 0) != 0) return 1;
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    CK(cudaEventElapsedTime(&ms, e0, e1));
  }
  printf("M %d N %d K %d amn %d bmn %d f32 %d acc %d: %.1f us (traced build)\n", M, N, K, amn, bmn, of32, acc, ms * 1e3);
  std::vector<long long> h(2 * 11 * 256);
  CK(cudaMemcpy(h.data(), trace, 2 * 11 * 256 * 8, cudaMemcpyDeviceToHost));
  for (int slot = 0; slot < 2; ++slot) {
    long long t0 = 0;
    for (int w = 0; w < 11; ++w) { long long* base = h.data() + (slot * 11 + w) * 256; if (base[0] > 0) { long long c = base[1] & 0xffffffffffffll; if (t0 == 0 || c < t0) t0 = c; } }
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
