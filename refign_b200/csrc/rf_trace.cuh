// Optional per-phase timeline for the warp-specialised kernels (tools/trace_attn_*.cu build the kernel files with
// -DWS_TRACE): lane 0 of every warp of two chosen blocks appends (tag << 48 | clock) records to a shared-memory log
// (inline and in shared memory: a call would spill the live score registers around every probe, a global counter costs
// a ~400-cycle round trip per probe); the log is flushed to global memory at the end of the kernel.
#pragma once
#ifdef WS_TRACE
namespace rf {
__device__ long long* g_ws_trace;
__device__ int g_ws_trace_blocks[2];
__shared__ long long ws_trace_buf[11][256];     // per warp: [0] = count, then the records
#ifndef WS_TRACE_SLOT
#define WS_TRACE_SLOT(w) (w)          // which of the 11 log rows a warp writes (-1: none); kernels with more warps remap
#endif
__device__ __forceinline__ void ws_trace(int tag) {
  const int warp = WS_TRACE_SLOT((int)(threadIdx.x >> 5));
  if ((threadIdx.x & 31) != 0 || warp < 0 || warp > 10) return;
  long long* base = ws_trace_buf[warp];
  const int n = (int)base[0];
  if (n < 254) {
    base[1 + n] = ((long long)tag << 48) | (clock64() & 0xffffffffffffll);
    base[0] = n + 1;
  }
}
__device__ __forceinline__ void ws_trace_init() {
  if (threadIdx.x < 11) ws_trace_buf[threadIdx.x][0] = 0;
  __syncthreads();
}
__device__ __forceinline__ void ws_trace_flush() {   // after the final __syncthreads
  __syncthreads();
  const int slot = (int)blockIdx.x == g_ws_trace_blocks[0] ? 0 : ((int)blockIdx.x == g_ws_trace_blocks[1] ? 1 : -1);
  if (slot < 0) return;
  for (int i = threadIdx.x; i < 11 * 256; i += blockDim.x) g_ws_trace[slot * 11 * 256 + i] = ws_trace_buf[i >> 8][i & 255];
}
}  // namespace rf
#define WS_T(tag) rf::ws_trace(tag)
#define WS_T_INIT() rf::ws_trace_init()
#define WS_T_FLUSH() rf::ws_trace_flush()
#else
#define WS_T(tag)
#define WS_T_INIT()
#define WS_T_FLUSH()
#endif
