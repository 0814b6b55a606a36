// Per-phase timeline of the tensor-core local correlation (local_corr_tc.cu built with -DWS_TRACE).  Not product code.
// usage: trace_local_corr [B C H W fuse] [block_a block_b]
#define WS_TRACE 1
#include "../refign_b200/csrc/local_corr_tc.cu"
#include <stdarg.h>
#include <vector>
namespace rf {
void set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); fputc('\n', stderr); }
}
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)
__global__ void fill_f32(float* p, long n, unsigned seed) {
  long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned h = (unsigned)i * 2654435761u ^ seed; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
  p[i] = ((h & 0xffff) / 32768.f - 1.f) * 0.1f;
}
int main(int argc, char** argv) {
  const int B = argc > 1 ? atoi(argv[1]) : 2, C = argc > 2 ? atoi(argv[2]) : 128, H = argc > 3 ? atoi(argv[3]) : 256, W = argc > 4 ? atoi(argv[4]) : 256;
  const int fuse = argc > 5 ? atoi(argv[5]) : 0;
  CK(cudaSetDevice(0));
  float *a, *b, *out, *nrm; long long* trace;
  const long n = (long)B * C * H * W;
  CK(cudaMalloc(&a, n * 4)); CK(cudaMalloc(&b, n * 4)); CK(cudaMalloc(&out, (long)B * 81 * H * W * 4)); CK(cudaMalloc(&nrm, (long)B * H * W * 4));
  CK(cudaMalloc(&trace, 2 * 11 * 256 * 8));
  fill_f32<<<(n + 255) / 256, 256>>>(a, n, 1);
  fill_f32<<<(n + 255) / 256, 256>>>(b, n, 2);
  CK(cudaDeviceSynchronize());
  int blocks[2] = {argc > 6 ? atoi(argv[6]) : 0, argc > 7 ? atoi(argv[7]) : 100};
  CK(cudaMemcpyToSymbol(rf::g_ws_trace, &trace, sizeof(trace)));
  CK(cudaMemcpyToSymbol(rf::g_ws_trace_blocks, blocks, sizeof(blocks)));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float ms = 0;
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaMemset(trace, 0, 2 * 11 * 256 * 8));
    CK(cudaEventRecord(e0));
    if (rf::local_corr_tc_launch(a, b, out, nrm, B, C, H, W, fuse != 0, 0) != 0) return 1;
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    CK(cudaEventElapsedTime(&ms, e0, e1));
  }
  printf("B %d C %d H %d W %d fuse %d: %.1f us (traced build)\n", B, C, H, W, fuse, ms * 1e3);
  std::vector<long long> h(2 * 11 * 256);
  CK(cudaMemcpy(h.data(), trace, 2 * 11 * 256 * 8, cudaMemcpyDeviceToHost));
  for (int slot = 0; slot < 1; ++slot) {
    long long t0 = 0;
    for (int w = 0; w < 11; ++w) { long long* base = h.data() + (slot * 11 + w) * 256; if (base[0] > 0) { long long c = base[1] & 0xffffffffffffll; if (t0 == 0 || c < t0) t0 = c; } }
    for (int w = 0; w < 11; ++w) {
      long long* base = h.data() + (slot * 11 + w) * 256;
      if (base[0] == 0) continue;
      printf("block %d warp %d:", blocks[slot], w);
      for (int i = 0; i < (int)base[0] && i < 60; ++i) printf(" %d@%lld", (int)(base[1 + i] >> 48), (base[1 + i] & 0xffffffffffffll) - t0);
      printf("\n");
    }
  }
  return 0;
}
