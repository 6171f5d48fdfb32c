// LayerNorm over the channel axis of channels-last tokens (memory-bound; one pass over x).
// Replaces nn.LayerNorm calls at attention.py:421,436 and video_model_builder.py:1249.
#include "common.cuh"

namespace mvit {

// C == 24*G: G lanes cooperate on a row, each keeps 24 values in registers (16-byte loads,
// every warp-level request covers whole 32-byte sectors).  32/G rows per warp and iteration.
// Warps walk the rows with a grid stride: gamma / beta are staged once per CTA in shared memory (16-byte reads),
// and the next row is already in flight while the current one is reduced, normalised and stored - the one-row-per-warp
// version spent a fixed ~20 us per launch on 48 scalar parameter loads per lane and on ramp-up / tail.
template <typename T, int G>
__global__ void __launch_bounds__(256) layernorm_rows_kernel(const T *__restrict__ x,
                                                             const float *__restrict__ gamma,
                                                             const float *__restrict__ beta,
                                                             T *__restrict__ y, int64_t rows,
                                                             float eps) {
  constexpr int VE = DType<T>::vec;   // elements per 16B vector
  constexpr int NV = 24 / VE;         // vectors per lane
  constexpr int C = 24 * G;
  constexpr int RPW = 32 / G;         // rows per warp and iteration
  const int lane = threadIdx.x & 31;
  const int sub = lane % G;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t stride = (int64_t)gridDim.x * (blockDim.x >> 5) * RPW;
  __shared__ __align__(16) float s_gb[2 * C];         // gamma | beta, read once per CTA
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    s_gb[c] = gamma[c];
    s_gb[C + c] = beta[c];
  }
  __syncthreads();
  auto load_row = [&](int64_t row, float (&v)[NV][VE]) {
    if (row < rows) {
      const T *px = x + row * C;
#pragma unroll
      for (int j = 0; j < NV; ++j) Vec16<T>::load(px + (j * G + sub) * VE, v[j]);
    } else {
#pragma unroll
      for (int j = 0; j < NV; ++j)
#pragma unroll
        for (int e = 0; e < VE; ++e) v[j][e] = 0.f;
    }
  };
  float v[NV][VE], vn[NV][VE];
  int64_t row = warp * RPW + lane / G;
  load_row(row, v);
  // every lane of a warp runs the same number of iterations (the shuffles are warp-wide)
  for (int64_t r0 = warp * RPW; r0 < rows; r0 += stride, row += stride) {
    load_row(row + stride, vn);                      // prefetch: in flight during the reductions below
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j)
#pragma unroll
      for (int e = 0; e < VE; ++e) s += v[j][e];
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.0f / C);
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j)
#pragma unroll
      for (int e = 0; e < VE; ++e) {
        const float d = v[j][e] - mean;
        ss += d * d;
      }
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float rstd = rsqrtf(ss * (1.0f / C) + eps);
    if (row < rows) {
      T *py = y + row * C;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        float o[VE];
        const int c0 = (j * G + sub) * VE;
#pragma unroll
        for (int e = 0; e < VE; e += 4) {
          const float4 g4 = *reinterpret_cast<const float4 *>(s_gb + c0 + e);
          const float4 b4 = *reinterpret_cast<const float4 *>(s_gb + C + c0 + e);
          o[e] = fmaf((v[j][e] - mean) * rstd, g4.x, b4.x);
          o[e + 1] = fmaf((v[j][e + 1] - mean) * rstd, g4.y, b4.y);
          o[e + 2] = fmaf((v[j][e + 2] - mean) * rstd, g4.z, b4.z);
          o[e + 3] = fmaf((v[j][e + 3] - mean) * rstd, g4.w, b4.w);
        }
        Vec16<T>::store(py + c0, o);
      }
    }
#pragma unroll
    for (int j = 0; j < NV; ++j)
#pragma unroll
      for (int e = 0; e < VE; ++e) v[j][e] = vn[j][e];
  }
}

// One-shot form of the kernel above (each warp handles 32/G rows once, no staging): at C <= 192 the tensors are large
// (>= 200k rows), occupancy matters more than the per-launch fixed cost, and this form streams at 5.6 TB/s.
template <typename T, int G>
__global__ void __launch_bounds__(256) layernorm_rows_once_kernel(const T *__restrict__ x,
                                                             const float *__restrict__ gamma,
                                                             const float *__restrict__ beta,
                                                             T *__restrict__ y, int64_t rows,
                                                             float eps) {
  constexpr int VE = DType<T>::vec;   // elements per 16B vector
  constexpr int NV = 24 / VE;         // vectors per lane
  constexpr int C = 24 * G;
  const int lane = threadIdx.x & 31;
  const int sub = lane % G;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t row = warp * (32 / G) + lane / G;
  const bool live = row < rows;
  float v[NV][VE];
  float s = 0.f;
  if (live) {
    const T *px = x + row * C;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      Vec16<T>::load(px + (j * G + sub) * VE, v[j]);
#pragma unroll
      for (int e = 0; e < VE; ++e) s += v[j][e];
    }
  } else {
#pragma unroll
    for (int j = 0; j < NV; ++j)
#pragma unroll
      for (int e = 0; e < VE; ++e) v[j][e] = 0.f;
  }
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s * (1.0f / C);
  float ss = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j)
#pragma unroll
    for (int e = 0; e < VE; ++e) {
      const float d = v[j][e] - mean;
      ss += d * d;
    }
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float rstd = rsqrtf(ss * (1.0f / C) + eps);
  if (live) {
    T *py = y + row * C;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int c0 = (j * G + sub) * VE;
      float o[VE];
#pragma unroll
      for (int e = 0; e < VE; ++e)
        o[e] = (v[j][e] - mean) * rstd * __ldg(gamma + c0 + e) + __ldg(beta + c0 + e);
      Vec16<T>::store(py + c0, o);
    }
  }
}

// any C: one warp per row, three passes (x stays in L1/L2 between them)
template <typename T>
__global__ void layernorm_generic_kernel(const T *__restrict__ x, const float *__restrict__ gamma,
                                         const float *__restrict__ beta, T *__restrict__ y,
                                         int64_t rows, int C, float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const T *px = x + row * C;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += to_f32(px[c]);
  const float mean = warp_sum(s) / C;
  float ss = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float d = to_f32(px[c]) - mean;
    ss += d * d;
  }
  const float rstd = rsqrtf(warp_sum(ss) / C + eps);
  T *py = y + row * C;
  for (int c = lane; c < C; c += 32)
    py[c] = from_f32<T>((to_f32(px[c]) - mean) * rstd * gamma[c] + beta[c]);
}

template <typename T>
static int launch_ln(const void *x, const float *g, const float *b, void *y, int64_t rows, int C,
                     float eps, cudaStream_t st) {
  const T *px = static_cast<const T *>(x);
  T *py = static_cast<T *>(y);
  const int threads = 256;
  auto grid_for = [&](int G) {
    const int64_t rows_per_block = (threads / 32) * (32 / G);
    const int64_t need = (rows + rows_per_block - 1) / rows_per_block;
    return (unsigned)std::min<int64_t>(need, (int64_t)num_sms() * 6);     // grid-stride: ~6 resident CTAs per SM
  };
  auto grid_once = [&](int G) {
    const int64_t rows_per_block = (threads / 32) * (32 / G);
    return (unsigned)((rows + rows_per_block - 1) / rows_per_block);
  };
  switch (C) {
    case 96: layernorm_rows_once_kernel<T, 4><<<grid_once(4), threads, 0, st>>>(px, g, b, py, rows, eps); break;
    case 192: layernorm_rows_once_kernel<T, 8><<<grid_once(8), threads, 0, st>>>(px, g, b, py, rows, eps); break;
    case 384: layernorm_rows_kernel<T, 16><<<grid_for(16), threads, 0, st>>>(px, g, b, py, rows, eps); break;
    case 768: layernorm_rows_kernel<T, 32><<<grid_for(32), threads, 0, st>>>(px, g, b, py, rows, eps); break;
    default: {
      const int64_t rpb = threads / 32;
      layernorm_generic_kernel<T><<<(unsigned)((rows + rpb - 1) / rpb), threads, 0, st>>>(px, g, b, py, rows, C, eps);
    }
  }
  MVIT_LAUNCH_OK("layernorm");
  return 0;
}

}  // namespace mvit

extern "C" int mvit_layernorm_fwd(const void *x, const float *gamma, const float *beta, void *y,
                                  int64_t rows, int channels, float eps, int dtype, void *stream) {
  using namespace mvit;
  MVIT_REQUIRE(x && y && gamma && beta, "layernorm: null pointer");
  MVIT_REQUIRE(rows >= 0 && channels > 0, "layernorm: bad shape rows=%lld C=%d", (long long)rows, channels);
  if (rows == 0) return 0;
  MVIT_REQUIRE(rows < (int64_t)1 << 31, "layernorm: too many rows");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == MVIT_F32) return launch_ln<float>(x, gamma, beta, y, rows, channels, eps, st);
  if (dtype == MVIT_BF16) return launch_ln<bf16>(x, gamma, beta, y, rows, channels, eps, st);
  MVIT_REQUIRE(false, "layernorm: unknown dtype %d", dtype);
}
