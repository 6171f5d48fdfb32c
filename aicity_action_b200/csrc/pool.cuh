#pragma once
#include "common.cuh"

namespace mvit {

struct PoolParams {
  int64_t in_bs, in_ls, in_hs, out_bs, out_ls, out_hs;
  int B, heads, d, T, H, W, kt, kh, kw, st, sh, sw, pt, ph, pw, To, Ho, Wo;
  int has_cls, has_ln;
  float eps;
  void *pre_out = nullptr;   // optional second output of the tuned kernel: the pooled values BEFORE the LayerNorm,
                             // contiguous [B, heads, L', d] (saved for backward instead of recomputing the conv)
};

// tuned path (pool_tiled.cu); returns 1 if it does not apply, 0 on launch, <0 on error
int pool_tiled_try(const void *in, const float *w, const float *g, const float *b, void *out,
                   const PoolParams &p, int mode, int dtype, cudaStream_t st);

// tuned conv weight gradient (pool_bwd_tiled.cu): dw[96, 27] += ; dy is [B, heads, L', 96] contiguous; same return codes
int pool_wgrad_tiled_try(const void *in, const void *dy, float *dw, const PoolParams &p, int dtype, cudaStream_t st);

}  // namespace mvit
