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
int attention_tc_fault_take();
int attention_bwd_tc_fault_take();
int gemm_tc_fault_take();
int gemm_wgrad_tc_fault_take();
int pool_tma_fault_take();
int mlp_fused_fault_take();
}  // namespace mvit

extern "C" int mvit_abi_version(void) { return 1; }
extern "C" int mvit_device_fault(void) {
  int n = 0;
  const char *names[6] = {"attention", "attention_bwd", "linear", "linear_wgrad", "attention_pool", "mlp_fused"};
  int (*take[6])() = {mvit::attention_tc_fault_take, mvit::attention_bwd_tc_fault_take, mvit::gemm_tc_fault_take,
                      mvit::gemm_wgrad_tc_fault_take, mvit::pool_tma_fault_take, mvit::mlp_fused_fault_take};
  for (int i = 0; i < 6; ++i) {
    const int v = take[i]();
    if (v < 0) {
      mvit::set_error("mvit_device_fault: reading the fault flag failed: %s", cudaGetErrorString(cudaGetLastError()));
      return -2;
    }
    if (v > 0) {
      mvit::set_error("a %s kernel abandoned an mbarrier wait (pipeline protocol fault): its results are invalid", names[i]);
      ++n;
    }
  }
  return n;
}
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
