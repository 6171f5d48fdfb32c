// Shared helpers for libmvit_b200 (sm_100a).  No torch / ATen types anywhere in this library.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/mvit_b200.h"

namespace mvit {

// thread-local error string behind mvit_last_error()
void set_error(const char *fmt, ...);

#define MVIT_REQUIRE(cond, ...)        \
  do {                                 \
    if (!(cond)) {                     \
      ::mvit::set_error(__VA_ARGS__);  \
      return -1;                       \
    }                                  \
  } while (0)

#define MVIT_CUDA_OK(expr)                                                          \
  do {                                                                              \
    cudaError_t e__ = (expr);                                                       \
    if (e__ != cudaSuccess) {                                                       \
      ::mvit::set_error("%s failed: %s", #expr, cudaGetErrorString(e__));           \
      return -2;                                                                    \
    }                                                                               \
  } while (0)

#define MVIT_LAUNCH_OK(name)                                                        \
  do {                                                                              \
    cudaError_t e__ = cudaGetLastError();                                           \
    if (e__ != cudaSuccess) {                                                       \
      ::mvit::set_error("launch of %s failed: %s", name, cudaGetErrorString(e__));  \
      return -3;                                                                    \
    }                                                                               \
  } while (0)

using bf16 = __nv_bfloat16;

template <typename T> struct DType;
template <> struct DType<float> {
  static constexpr int id = MVIT_F32;
  static constexpr int vec = 4;  // elements per 16-byte vector
};
template <> struct DType<bf16> {
  static constexpr int id = MVIT_BF16;
  static constexpr int vec = 8;
};

__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f32<bf16>(float v) { return __float2bfloat16_rn(v); }

// 16-byte vector <-> fp32 registers
template <typename T> struct Vec16;
template <> struct Vec16<float> {
  static constexpr int N = 4;
  __device__ __forceinline__ static void load(const float *p, float (&f)[4]) {
    float4 v = *reinterpret_cast<const float4 *>(p);
    f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
  }
  __device__ __forceinline__ static void store(float *p, const float (&f)[4]) {
    *reinterpret_cast<float4 *>(p) = make_float4(f[0], f[1], f[2], f[3]);
  }
};
template <> struct Vec16<bf16> {
  static constexpr int N = 8;
  __device__ __forceinline__ static void load(const bf16 *p, float (&f)[8]) {
    uint4 v = *reinterpret_cast<const uint4 *>(p);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      f[2 * i] = __uint_as_float(w[i] << 16);
      f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
  __device__ __forceinline__ static void store(bf16 *p, const float (&f)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t *>(&h);
    }
    *reinterpret_cast<uint4 *>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}
// d/dx of the exact GELU: Phi(x) + x * phi(x)
__device__ __forceinline__ float gelu_grad(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

// SM count of the CURRENT device (cached per device: a process may drive several GPUs; racing first calls write
// the same value).
inline int num_sms() {
  static std::atomic<int> cache[64];
  int dev = 0;
  cudaGetDevice(&dev);
  std::atomic<int> &slot = cache[dev & 63];
  int n = slot.load(std::memory_order_relaxed);
  if (n == 0) {
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
    slot.store(n, std::memory_order_relaxed);
  }
  return n;
}

// ---------------------------------------------------------------- programmatic dependent launch (PDL)
// The forward is a chain of ~170 short dependent kernels.  A kernel launched through launch_pdl() may become resident while
// its predecessor in the stream is still draining: everything it does before pdl_wait() (shared-memory carve-up, mbarrier
// init, TMEM allocation, tensor-map prefetch) overlaps the predecessor's tail; pdl_wait() returns once the predecessor
// grid has COMPLETED and its writes are visible, so no global memory may be touched before it.  Every kernel launched
// this way must execute pdl_wait() in all threads that read or write global memory (completion is transitive: a kernel
// cannot finish before its own wait returned).  pdl_launch_dependents() only allows the next grid to be scheduled early.
// MVIT_B200_PDL=0 launches the same kernels fully serialised (the wait is then a no-op).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

inline bool pdl_enabled() {
  static const bool on = [] {
    const char *e = getenv("MVIT_B200_PDL");
    return e == nullptr || atoi(e) != 0;
  }();
  return on;
}
// Fills `attr` (room for one more entry at index n) with the PDL attribute when enabled; returns the new count.
inline int pdl_attr(cudaLaunchAttribute *attr, int n) {
  if (!pdl_enabled()) return n;
  attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[n].val.programmaticStreamSerializationAllowed = 1;
  return n + 1;
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args &&...args) {
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute attr[1];
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cfg.attrs = attr;
  cfg.numAttrs = (unsigned)pdl_attr(attr, 0);
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) applies to the device that is current when it is called: do it once
// per (kernel, device), thread-safe (a bit per device in an atomic mask; a lost race only repeats an idempotent call).
#define MVIT_SMEM_OPT_IN(kernel, bytes)                                                                          \
  do {                                                                                                           \
    static std::atomic<uint64_t> done__{0};                                                                      \
    int dev__ = 0;                                                                                               \
    cudaGetDevice(&dev__);                                                                                       \
    const uint64_t bit__ = 1ull << (dev__ & 63);                                                                 \
    if (!(done__.load(std::memory_order_acquire) & bit__)) {                                                     \
      MVIT_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)));     \
      done__.fetch_or(bit__, std::memory_order_release);                                                         \
    }                                                                                                            \
  } while (0)

}  // namespace mvit
