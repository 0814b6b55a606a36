// Library-level entry points: error string, ABI version, device check.
#include <stdarg.h>

#include "rf_common.cuh"

namespace rf {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace rf

extern "C" const char* rf_last_error(void) { return rf::g_err; }
extern "C" int rf_version(void) { return 1; }
extern "C" int rf_device_check(void) {
  int dev = 0;
  cudaDeviceProp p;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&p, dev) != cudaSuccess) {
    rf::set_error("rf_device_check: no CUDA device");
    return RF_ENODEV;
  }
  if (p.major != 10) {
    rf::set_error("rf_device_check: device %d is sm_%d%d, this library is built for sm_100a only", dev, p.major, p.minor);
    return RF_ENODEV;
  }
  return RF_OK;
}

// Binds `device` (and its primary context) to the calling host thread for this library's CUDA runtime
// instance.  The library links cudart statically, so its per-thread current device is separate from
// the caller's runtime (e.g. torch's): a fresh thread -- autograd's backward worker -- would otherwise
// default to device 0 and have no context current for driver-API calls (cuTensorMapEncodeTiled).
extern "C" int rf_set_device(int device) {
  cudaError_t e = cudaSetDevice(device);
  if (e == cudaSuccess) e = cudaFree(nullptr);
  if (e != cudaSuccess) {
    rf::set_error("rf_set_device(%d): %s", device, cudaGetErrorString(e));
    return RF_ECUDA;
  }
  return RF_OK;
}
