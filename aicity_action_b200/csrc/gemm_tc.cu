#include "linear.cuh"
namespace mvit {
bool linear_tc_supported(const LinearArgs &, const char **why) { *why = "not built yet"; return false; }
int linear_tc(const LinearArgs &, cudaStream_t) { set_error("linear_tc: not built"); return -1; }
}  // namespace mvit
