// Library-level entry points: version, error string, device check.
#include <stdarg.h>

#include "common.cuh"

namespace mvit {
static thread_local char g_err[512] = "";
void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace mvit

extern "C" int mvit_abi_version(void) { return 1; }
extern "C" const char *mvit_last_error(void) { return mvit::g_err; }
extern "C" int mvit_device_supported(void) {
  int dev = 0;
  cudaDeviceProp prop;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
    mvit::set_error("no CUDA device: %s", cudaGetErrorString(cudaGetLastError()));
    return -2;
  }
  return prop.major == 10 ? 1 : 0;
}
