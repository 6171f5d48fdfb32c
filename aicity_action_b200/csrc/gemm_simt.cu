// CUDA-core Linear: y = epi(x·wᵀ + bias)·row_scale + residual, fp32 accumulation, any M/N/K.
// This is the fp32 parity path (north_star: fp32 logits within 1e-4) and the cross-check for the
// tcgen05 GEMM (gemm_tc.cu); bf16 inference never takes it unless forced with MVIT_IMPL_SIMT.
#include "linear.cuh"

namespace mvit {

constexpr int BM = 64, BN = 64, BK = 16;

template <typename T>
__global__ void __launch_bounds__(256) linear_simt_kernel(LinearArgs a) {
  __shared__ float sA[BK][BM + 4];
  __shared__ float sB[BK][BN + 4];
  const T *x = static_cast<const T *>(a.x);
  const T *w = static_cast<const T *>(a.w);
  const int tid = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int tm = (tid / 16) * 4, tn = (tid % 16) * 4;
  float acc[4][4] = {};
  // loader mapping: 256 threads load 64 rows x 16 k (4 per thread along k)
  const int lr = tid / 4, lk = (tid % 4) * 4;
  for (int k0 = 0; k0 < a.K; k0 += BK) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int k = k0 + lk + e;
      const int64_t m = m0 + lr;
      const int n = n0 + lr;
      sA[lk + e][lr] = (m < a.M && k < a.K) ? to_f32(x[m * a.K + k]) : 0.f;
      sB[lk + e][lr] = (n < a.N && k < a.K) ? to_f32(w[(int64_t)n * a.K + k]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 av = *reinterpret_cast<const float4 *>(&sA[k][tm]);
      const float4 bv = *reinterpret_cast<const float4 *>(&sB[k][tn]);
      const float ar[4] = {av.x, av.y, av.z, av.w};
      const float br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
    __syncthreads();
  }
  T *y = static_cast<T *>(a.y);
  const T *res = static_cast<const T *>(a.residual);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + tm + i;
    if (m >= a.M) continue;
    const float rs = a.row_scale ? a.row_scale[m / a.rows_per_sample] : 1.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tn + j;
      if (n >= a.N) continue;
      float v = acc[i][j] + (a.bias ? a.bias[n] : 0.f);
      if (a.epilogue == MVIT_EPI_GELU) v = gelu_erf(v);
      if (a.row_scale) v *= rs;
      if (a.epilogue == MVIT_EPI_GELU_GRAD) v = gelu_grad(v) * to_f32(res[m * a.ldr + n]);
      else if (res) v += to_f32(res[(a.res_period ? m % a.res_period : m) * a.ldr + n]);
      y[m * a.ldy + n] = from_f32<T>(v);
    }
  }
}

int linear_simt(const LinearArgs &a, int dtype, cudaStream_t st) {
  dim3 grid((unsigned)((a.M + BM - 1) / BM), (unsigned)((a.N + BN - 1) / BN));
  MVIT_REQUIRE(grid.y < 65536, "linear: N too large for the CUDA-core path");
  if (dtype == MVIT_F32) linear_simt_kernel<float><<<grid, 256, 0, st>>>(a);
  else linear_simt_kernel<bf16><<<grid, 256, 0, st>>>(a);
  MVIT_LAUNCH_OK("linear(simt)");
  return 0;
}

}  // namespace mvit

extern "C" int mvit_linear_fwd(const void *x, const void *w, const float *bias, const void *residual,
                               const float *row_scale, int64_t rows_per_sample, void *y, int64_t M,
                               int N, int K, int64_t ldy, int64_t ldr, int64_t residual_row_period,
                               int epilogue, int dtype, int impl, void *stream) {
  using namespace mvit;
  MVIT_REQUIRE(x && w && y, "linear: null pointer");
  MVIT_REQUIRE(M >= 0 && N > 0 && K > 0, "linear: bad shape M=%lld N=%d K=%d", (long long)M, N, K);
  MVIT_REQUIRE(epilogue == MVIT_EPI_NONE || epilogue == MVIT_EPI_GELU || epilogue == MVIT_EPI_GELU_GRAD,
               "linear: unknown epilogue %d", epilogue);
  MVIT_REQUIRE(epilogue != MVIT_EPI_GELU_GRAD || (residual && residual_row_period == 0 && !row_scale),
               "linear: EPI_GELU_GRAD needs a full residual (the upstream gradient) and no row_scale");
  MVIT_REQUIRE(dtype == MVIT_F32 || dtype == MVIT_BF16, "linear: unknown dtype %d", dtype);
  MVIT_REQUIRE(ldy >= N && (!residual || ldr >= N), "linear: leading dimension smaller than N");
  MVIT_REQUIRE(!row_scale || rows_per_sample > 0, "linear: row_scale needs rows_per_sample");
  if (M == 0) return 0;
  MVIT_REQUIRE(residual_row_period >= 0, "linear: negative residual_row_period");
  LinearArgs a{x, w, residual, bias, row_scale, y, M, rows_per_sample, ldy, ldr, residual ? residual_row_period : 0, N, K, epilogue};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  bool use_tc = false;
  if (impl == MVIT_IMPL_TCGEN05 || (impl == MVIT_IMPL_AUTO && dtype == MVIT_BF16)) {
    const char *why = "";
    MVIT_REQUIRE(dtype == MVIT_BF16, "linear: the tcgen05 path is bf16 only");
    if (linear_tc_supported(a, &why)) use_tc = true;
    else MVIT_REQUIRE(impl == MVIT_IMPL_AUTO, "linear: tcgen05 path rejected: %s", why);
  }
  if (use_tc) return linear_tc(a, st);
  return linear_simt(a, dtype, st);
}
