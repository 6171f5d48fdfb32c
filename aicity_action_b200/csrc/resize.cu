// Frame gather + uint8 bilinear resize on the device, bit-exact with OpenCV's `cv2.resize(u8, INTER_LINEAR)`.
//
// Replaces, for the sliding-window input pipeline, the per-window host work of
// scripts/module_wrapper.py:304-331: `video.get_batch(frame_idxs)` followed by `cv2.resize(frame, (S, S),
// interpolation=cv2.INTER_LINEAR)` on every uint8 frame BEFORE the cast to float (scripts/utils.py:207-211,
// keep_scale=False: the aspect ratio is ignored).  The host uploads each needed raw frame once; this kernel gathers
// the frames of every window by index and resizes them straight into the [B, T, S, S, 3] uint8 clip buffer that the
// patch-embed fold normalises.
//
// Arithmetic = OpenCV's fixed-point path (modules/imgproc/src/resize.cpp: hal::resize tables, HResizeLinear<uchar,int,
// short,2048>, VResizeLinear<uchar,int,short,FixedPtCast<int,uchar,22>>), restated in oracle/resize_oracle.py:
//   fx = float((dx + 0.5) * scale_x - 0.5) (double product, one rounding to float); sx = floor(fx); fx -= sx
//   columns: sx < 0 -> (0, 0);  sx >= W-1 -> (W-1, 0);  rows: indices clamp to [0, H-1], weights are kept
//   alpha/beta = round-half-even(w * 2048) as int16
//   D[dx] = S[sx] * a0 + S[sx+1] * a1;   out = (((b0 * (D0 >> 4)) >> 16) + ((b1 * (D1 >> 4)) >> 16) + 2) >> 2
//
// HBM-bound (≈ 1.6 MB read + 0.6 MB written per 540p -> 448 frame).  A CTA owns one output row of one output frame:
// the two source rows it blends are staged in shared memory with 16-byte loads, each thread produces whole pixels,
// and the finished row leaves with 16-byte stores.
#include <algorithm>

#include "common.cuh"

namespace mvit {
namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ void linear_coef(int d, double scale, int src, bool clamp_weights, int &s, int &c0, int &c1) {
  // product and difference rounded separately, as the host code of the reference does (no FMA contraction)
  float f = (float)__dadd_rn(__dmul_rn((double)d + 0.5, scale), -0.5);
  s = (int)floorf(f);
  f -= (float)s;
  if (clamp_weights) {
    if (s < 0) { s = 0; f = 0.f; }
    if (s >= src - 1) { s = src - 1; f = 0.f; }
  }
  c0 = max(-32768, min(32767, __float2int_rn((1.0f - f) * 2048.0f)));
  c1 = max(-32768, min(32767, __float2int_rn(f * 2048.0f)));
}

template <int C>
__global__ void __launch_bounds__(kThreads)
resize_gather_u8_kernel(const uint8_t *__restrict__ src, const int32_t *__restrict__ frame_idx, uint8_t *__restrict__ dst,
                        int n_src, int H, int W, int out_h, int out_w, double scale_x, double scale_y, int vec_ok) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int row_bytes = W * C, out_bytes = out_w * C;
  const int row_pad = (row_bytes + 15) & ~15, out_pad = (out_bytes + 15) & ~15;
  uint8_t *r0 = smem, *r1 = smem + row_pad, *orow = smem + 2 * row_pad;
  const int oy = blockIdx.x, fo = blockIdx.y;
  int f = frame_idx ? frame_idx[fo] : fo;
  f = max(0, min(n_src - 1, f));
  int sy, b0, b1;
  linear_coef(oy, scale_y, H, false, sy, b0, b1);
  const int y0 = max(0, min(H - 1, sy)), y1 = max(0, min(H - 1, sy + 1));
  const uint8_t *g0 = src + ((int64_t)f * H + y0) * row_bytes, *g1 = src + ((int64_t)f * H + y1) * row_bytes;
  if (vec_ok) {
    const int nv = row_bytes >> 4;
    for (int i = threadIdx.x; i < nv; i += kThreads) {
      reinterpret_cast<uint4 *>(r0)[i] = __ldg(reinterpret_cast<const uint4 *>(g0) + i);
      reinterpret_cast<uint4 *>(r1)[i] = __ldg(reinterpret_cast<const uint4 *>(g1) + i);
    }
  } else {
    for (int i = threadIdx.x; i < row_bytes; i += kThreads) {
      r0[i] = __ldg(g0 + i);
      r1[i] = __ldg(g1 + i);
    }
  }
  __syncthreads();
  for (int ox = threadIdx.x; ox < out_w; ox += kThreads) {
    int sx, a0, a1;
    linear_coef(ox, scale_x, W, true, sx, a0, a1);
    const int x1 = min(sx + 1, W - 1);
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const int d0 = (int)r0[sx * C + c] * a0 + (int)r0[x1 * C + c] * a1;
      const int d1 = (int)r1[sx * C + c] * a0 + (int)r1[x1 * C + c] * a1;
      const int v = (((b0 * (d0 >> 4)) >> 16) + ((b1 * (d1 >> 4)) >> 16) + 2) >> 2;
      orow[ox * C + c] = (uint8_t)max(0, min(255, v));
    }
  }
  __syncthreads();
  uint8_t *o = dst + ((int64_t)fo * out_h + oy) * out_bytes;
  if (vec_ok) {
    for (int i = threadIdx.x; i < (out_bytes >> 4); i += kThreads)
      reinterpret_cast<uint4 *>(o)[i] = reinterpret_cast<const uint4 *>(orow)[i];
  } else {
    for (int i = threadIdx.x; i < out_bytes; i += kThreads) o[i] = orow[i];
  }
  (void)out_pad;
}

}  // namespace
}  // namespace mvit

extern "C" int mvit_resize_gather_u8(const uint8_t *src, int n_src, int H, int W, const int32_t *frame_idx, int n_out,
                                     uint8_t *dst, int out_h, int out_w, int channels, void *stream) {
  using namespace mvit;
  MVIT_REQUIRE(src && dst, "resize: null pointer");
  MVIT_REQUIRE(n_src > 0 && H > 0 && W > 0 && out_h > 0 && out_w > 0 && n_out >= 0, "resize: bad shape");
  MVIT_REQUIRE(channels == 3 || channels == 1, "resize: 1 or 3 interleaved channels supported, got %d", channels);
  MVIT_REQUIRE(frame_idx != nullptr || n_out == n_src, "resize: without an index list n_out must equal n_src");
  MVIT_REQUIRE(n_out < 65536 && out_h < (1 << 30), "resize: too many output frames per call (max 65535)");
  if (n_out == 0) return 0;
  const int row_bytes = W * channels, out_bytes = out_w * channels;
  const size_t smem = 2 * (size_t)((row_bytes + 15) & ~15) + (size_t)((out_bytes + 15) & ~15);
  MVIT_REQUIRE(smem <= 200 * 1024, "resize: source rows of %d bytes do not fit in shared memory", row_bytes);
  // the reference derives the scale from the inverse ratio in double (cv::resize -> hal::resize)
  const double scale_x = 1.0 / ((double)out_w / (double)W), scale_y = 1.0 / ((double)out_h / (double)H);
  const int vec_ok = (row_bytes % 16 == 0) && (out_bytes % 16 == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(dst) & 15) == 0);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  dim3 grid((unsigned)out_h, (unsigned)n_out);
  if (channels == 3) {
    if (smem > 48 * 1024) MVIT_SMEM_OPT_IN(resize_gather_u8_kernel<3>, 200 * 1024);
    resize_gather_u8_kernel<3><<<grid, kThreads, smem, st>>>(src, frame_idx, dst, n_src, H, W, out_h, out_w, scale_x,
                                                              scale_y, vec_ok);
  } else {
    if (smem > 48 * 1024) MVIT_SMEM_OPT_IN(resize_gather_u8_kernel<1>, 200 * 1024);
    resize_gather_u8_kernel<1><<<grid, kThreads, smem, st>>>(src, frame_idx, dst, n_src, H, W, out_h, out_w, scale_x,
                                                              scale_y, vec_ok);
  }
  MVIT_LAUNCH_OK("resize_gather_u8");
  return 0;
}
