#pragma once
#include "common.cuh"

namespace mvit {

struct AttnArgs {
  const void *q, *k, *v;  // [B, heads, L, 96]
  void *out;              // [B, Lq, heads*96]
  float *lse;             // [B, heads, Lq] or NULL
  int B, heads, Lq, Lk;
  float scale;
  int add_q;
};

int attention_simt(const AttnArgs &a, int dtype, cudaStream_t st);
// tcgen05 / TMEM / TMA path (attention_tc.cu), bf16 only
int attention_tc(const AttnArgs &a, cudaStream_t st);
bool attention_tc_supported(const AttnArgs &a, const char **why);

}  // namespace mvit
