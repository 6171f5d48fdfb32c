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
  // LayerNorm folding (tcgen05 path only; see gemm_tc.cu): at most one of {ln_stats, stats_out} is set
  const float *colsum = nullptr;      // [N] sum_k w'[n,k]            (with ln_stats)
  const float *ln_stats = nullptr;    // [ln_parts][M][2] input-row (sum, sum of squares)
  float *stats_out = nullptr;         // [linear_tc_stat_parts()][M][2] output-row (sum, sum of squares)
  int ln_parts = 0;
  float ln_eps = 0.f;
};

int linear_simt(const LinearArgs &a, int dtype, cudaStream_t st);
// tcgen05 path (gemm_tc.cu), bf16 only
int linear_tc(const LinearArgs &a, cudaStream_t st);
bool linear_tc_supported(const LinearArgs &a, const char **why);
int patch_conv_tc(const void *folded, const void *wf, const float *bias, const void *pos, void *out, float *stats_out, int B,
                  int Tf, int Hf, int Wf, int Cf, int nt, int nh, int nw, int lo_t, int lo_h, int lo_w, int N, cudaStream_t st);
// number of N tiles (= row-statistics parts written through LinearArgs::stats_out) the tcgen05 path uses for this shape
int linear_tc_stat_parts(int64_t M, int N, int K);

// weight / bias gradient on tcgen05 (gemm_wgrad_tc.cu), bf16 only: dw[N,K] += dy^T x, db[N] += colsum(dy)
int linear_wgrad_tc(const void *dy, const void *x, float *dw, float *db, int64_t M, int N, int K, cudaStream_t st);
bool linear_wgrad_tc_supported(const void *dy, const void *x, const float *dw, int64_t M, int N, int K, const char **why);

}  // namespace mvit
