// Pieces shared by the tuned attention_pool kernels (forward: pool_tiled.cu, weight gradient: pool_bwd_tiled.cu):
// per-dtype 3-channel lane I/O, halo-tile geometry for stride (1,S,S), the cp.async helper.
#pragma once
#include "pool.cuh"

namespace mvit {
namespace ptile {

constexpr int TW = 8;
constexpr int kThreads = 128;
constexpr int kWarps = kThreads / 32;
constexpr int kStgPitch = 100;   // floats per staged column (96 + 4 pad: conflict-free 16-byte reads)

template <typename T> struct IO;
template <> struct IO<bf16> {
  static constexpr int kPitch = 192;  // bytes per position
  __device__ __forceinline__ static void load3(const uint8_t *pos, int lane, float2 &xy, float &z) {
    const uint32_t u = *reinterpret_cast<const uint32_t *>(pos + 4 * lane);
    xy.x = __uint_as_float(u << 16);
    xy.y = __uint_as_float(u & 0xffff0000u);
    z = __uint_as_float((uint32_t)(*reinterpret_cast<const uint16_t *>(pos + 128 + 2 * lane)) << 16);
  }
  __device__ __forceinline__ static void store3(bf16 *row, int lane, float2 xy, float z) {
    *reinterpret_cast<__nv_bfloat162 *>(row + 2 * lane) = __floats2bfloat162_rn(xy.x, xy.y);
    row[64 + lane] = __float2bfloat16_rn(z);
  }
};
template <> struct IO<float> {
  static constexpr int kPitch = 384;
  __device__ __forceinline__ static void load3(const uint8_t *pos, int lane, float2 &xy, float &z) {
    xy = *reinterpret_cast<const float2 *>(pos + 8 * lane);
    z = *reinterpret_cast<const float *>(pos + 256 + 4 * lane);
  }
  __device__ __forceinline__ static void store3(float *row, int lane, float2 xy, float z) {
    *reinterpret_cast<float2 *>(row + 2 * lane) = xy;
    row[64 + lane] = z;
  }
};

template <int S, int CPW> struct Geo {
  static constexpr int MS = S < 3 ? S : 3;              // compact stride between neighbouring outputs
  static constexpr int TH = CPW == 8 ? kWarps : kWarps / 2;   // one warp per row, or two warps per row
  static constexpr int NR = (TH - 1) * MS + 3;          // compact input rows / cols of the halo tile
  static constexpr int NC = (TW - 1) * MS + 3;
  static constexpr int NPOS = NR * NC;
  static constexpr int NPOS_PAD = (NPOS + 3) & ~3;     // keeps what follows the offset table 16-byte aligned
  static constexpr int WC = (CPW - 1) * MS + 3;         // compact cols one warp touches
};

__device__ __forceinline__ void cp_async16(void *dst, const void *src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)),
               "l"(src), "r"(src_bytes)
               : "memory");
}

template <typename T> struct Ring { static constexpr int kStages = sizeof(T) == 2 ? 4 : 3; };

}  // namespace ptile
}  // namespace mvit
