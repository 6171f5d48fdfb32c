#pragma once
#include "common.cuh"

namespace mvit {

struct LinearArgs {
  const void *x, *w, *residual;
  const float *bias, *row_scale;
  void *y;
  int64_t M, rows_per_sample, ldy, ldr;
  int64_t res_period;   // residual row = m % res_period (0: residual has M rows)
  int N, K, epilogue;
};

int linear_simt(const LinearArgs &a, int dtype, cudaStream_t st);
// tcgen05 path (gemm_tc.cu), bf16 only
int linear_tc(const LinearArgs &a, cudaStream_t st);
bool linear_tc_supported(const LinearArgs &a, const char **why);
int patch_conv_tc(const void *folded, const void *wf, const float *bias, const void *pos, void *out, int B, int Tf,
                  int Hf, int Wf, int Cf, int nt, int nh, int nw, int lo_t, int lo_h, int lo_w, int N, cudaStream_t st);

// weight / bias gradient on tcgen05 (gemm_wgrad_tc.cu), bf16 only: dw[N,K] += dy^T x, db[N] += colsum(dy)
int linear_wgrad_tc(const void *dy, const void *x, float *dw, float *db, int64_t M, int N, int K, cudaStream_t st);
bool linear_wgrad_tc_supported(const void *dy, const void *x, const float *dw, int64_t M, int N, int K, const char **why);

}  // namespace mvit
