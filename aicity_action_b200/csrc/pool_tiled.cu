// Tuned attention_pool path (3x3x3 depthwise conv, stride (1,s,s), d = 96) — see DESIGN.md §kernels.
#include "common.cuh"

namespace mvit {
struct PoolParams;
// returns 1 when the tuned path does not apply (caller falls through to the generic CUDA kernel)
int pool_tiled_try(const void *, const float *, const float *, const float *, void *,
                   const PoolParams &, int, int, cudaStream_t) {
  return 1;
}
}  // namespace mvit
