#include "attention.cuh"
namespace mvit {
bool attention_tc_supported(const AttnArgs &, const char **why) { *why = "not built yet"; return false; }
int attention_tc(const AttnArgs &, cudaStream_t) { set_error("attention_tc: not built"); return -1; }
}  // namespace mvit
