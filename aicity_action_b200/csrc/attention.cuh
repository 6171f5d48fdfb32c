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
  // default-off relative-position operand (relpos.cu): [B, heads, L, 64] in the activation dtype;
  // scores become scale * (q.k + q_ext.k_ext).  Both NULL = the reference's attention.
  const void *q_ext = nullptr, *k_ext = nullptr;
};

int attention_simt(const AttnArgs &a, int dtype, cudaStream_t st);
// tcgen05 / TMEM / TMA path (attention_tc.cu), bf16 only
int attention_tc(const AttnArgs &a, cudaStream_t st);
bool attention_tc_supported(const AttnArgs &a, const char **why);

struct AttnBwdArgs {
  const void *q, *k, *v;   // [B, heads, L, 96]
  const void *out, *dout;  // [B, Lq, heads*96]
  const float *lse;        // [B, heads, Lq]
  void *dq;                // [B, heads, Lq, 96]
  float *dk, *dv;          // [B, heads, Lk, 96] fp32, accumulated
  float *workspace;        // attention_bwd_workspace_floats(B, heads, Lq) floats or NULL
  int B, heads, Lq, Lk;
  float scale;
  int add_q;
};
// tcgen05 backward (attention_bwd_tc.cu), bf16 only
int attention_bwd_tc(const AttnBwdArgs &a, cudaStream_t st);
bool attention_bwd_tc_supported(const AttnBwdArgs &a, const char **why);
size_t attention_bwd_workspace_floats(int B, int heads, int Lq);

}  // namespace mvit
