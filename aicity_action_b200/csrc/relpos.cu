// Decomposed relative-position operands for the fused attention (default-off feature; north_star item 2).
//
// NOT in the reference: this restates upstream PySlowFast's cal_rel_pos_spatial / cal_rel_pos_temporal as
// SURVEY.md Appendix F records them, and is validated only against the in-repo restatement oracle/mvit_oracle.py
// (`rel_pos_bias`) — parity unpinned.
//
//   bias[i, j] = q_i . Rh[h_i, h'_j] + q_i . Rw[w_i, w'_j] + q_i . Rt[t_i, t'_j],   R*[a, b] = rel_pos_*[dist(a, b)]
//
// is a product of per-query tables A (k_h + k_w + k_t columns) and one-hot key indicators E.  This file builds both as
// 64-column operands that the attention kernels append to the Q / K contraction (attention_tc.cu REL, attention_simt.cu):
//   q_ext[bh, i, :] = [A_h(i, 0..k_h) | A_w(i, 0..k_w) | A_t(i, 0..k_t) | 0] / scale      (scores are scale * (q.k + q_ext.k_ext))
//   k_ext[bh, j, :] = one-hot(h'_j) | one-hot(w'_j) | one-hot(t'_j) | 0
// Tokens are ordered (t, h, w), w fastest, no cls token.
#include "common.cuh"

namespace mvit {
namespace relpos {

constexpr int D = 96, E = 64, ROWS = 64;

struct Geo {
  int qt, qh, qw, kt, kh, kw;
  float inv_scale;
};

// upstream index: dist = a * max(k/q, 1) - b * max(q/k, 1) + (k - 1) * max(q/k, 1), truncated
__device__ __forceinline__ int rel_index(int a, int b, int qn, int kn) {
  const float qr = fmaxf((float)kn / (float)qn, 1.0f), kr = fmaxf((float)qn / (float)kn, 1.0f);
  return (int)((float)a * qr - (float)b * kr + (float)(kn - 1) * kr);
}

template <typename T>
__global__ void __launch_bounds__(256) q_tables_kernel(const T *__restrict__ q, const float *__restrict__ rel_h,
                                                        const float *__restrict__ rel_w, const float *__restrict__ rel_t,
                                                        T *__restrict__ q_ext, int Lq, Geo g) {
  __shared__ float sQ[ROWS][D + 1];
  const int bh = blockIdx.y, r0 = blockIdx.x * ROWS;
  const T *qb = q + ((int64_t)bh * Lq + r0) * D;
  for (int i = threadIdx.x; i < ROWS * D; i += 256) {
    const int r = i / D, c = i % D;
    sQ[r][c] = (r0 + r < Lq) ? to_f32(qb[(int64_t)r * D + c]) : 0.f;
  }
  __syncthreads();
  const int j = threadIdx.x & (E - 1);
  // which table this output column belongs to, and the key coordinate it stands for
  const float *tab = nullptr;
  int kpos = 0, qn = 1, kn = 1, which = 3;
  if (j < g.kh) { tab = rel_h; kpos = j; qn = g.qh; kn = g.kh; which = 0; }
  else if (j < g.kh + g.kw) { tab = rel_w; kpos = j - g.kh; qn = g.qw; kn = g.kw; which = 1; }
  else if (j < g.kh + g.kw + g.kt) { tab = rel_t; kpos = j - g.kh - g.kw; qn = g.qt; kn = g.kt; which = 2; }
  for (int r = threadIdx.x / E; r < ROWS; r += 256 / E) {
    const int row = r0 + r;
    if (row >= Lq) break;
    float acc = 0.f;
    if (tab != nullptr) {
      const int w = row % g.qw, h = (row / g.qw) % g.qh, t = row / (g.qw * g.qh);
      const int qpos = which == 0 ? h : (which == 1 ? w : t);
      const float4 *tr = reinterpret_cast<const float4 *>(tab + (int64_t)rel_index(qpos, kpos, qn, kn) * D);
#pragma unroll 4
      for (int c = 0; c < D / 4; ++c) {
        const float4 v = __ldg(tr + c);
        acc = fmaf(sQ[r][4 * c + 0], v.x, acc);
        acc = fmaf(sQ[r][4 * c + 1], v.y, acc);
        acc = fmaf(sQ[r][4 * c + 2], v.z, acc);
        acc = fmaf(sQ[r][4 * c + 3], v.w, acc);
      }
    }
    q_ext[((int64_t)bh * Lq + row) * E + j] = from_f32<T>(acc * g.inv_scale);
  }
}

template <typename T>
__global__ void k_onehot_kernel(T *__restrict__ k_ext, int64_t total, int Lk, Geo g) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % E);
  const int key = (int)((i / E) % Lk);
  const int w = key % g.kw, h = (key / g.kw) % g.kh, t = key / (g.kw * g.kh);
  const bool one = c == h || c == g.kh + w || c == g.kh + g.kw + t;
  k_ext[i] = from_f32<T>(one ? 1.f : 0.f);
}

}  // namespace relpos
}  // namespace mvit

/* See include/mvit_b200.h. */
extern "C" int mvit_relpos_operands_fwd(const void *q, const float *rel_h, const float *rel_w, const float *rel_t,
                                        void *q_ext, void *k_ext, int BH, int qt, int qh, int qw, int kt, int kh, int kw,
                                        float scale, int dtype, void *stream) {
  using namespace mvit;
  MVIT_REQUIRE(q && rel_h && rel_w && rel_t && q_ext && k_ext, "relpos: null pointer");
  MVIT_REQUIRE(BH > 0 && qt > 0 && qh > 0 && qw > 0 && kt > 0 && kh > 0 && kw > 0, "relpos: bad grid");
  MVIT_REQUIRE(kt + kh + kw <= relpos::E, "relpos: k_t + k_h + k_w = %d exceeds the %d bias columns", kt + kh + kw, relpos::E);
  MVIT_REQUIRE(BH < 65536, "relpos: B*heads too large");
  MVIT_REQUIRE(scale > 0.f, "relpos: scale must be positive");
  MVIT_REQUIRE(dtype == MVIT_F32 || dtype == MVIT_BF16, "relpos: unknown dtype %d", dtype);
  MVIT_REQUIRE((reinterpret_cast<uintptr_t>(rel_h) & 15) == 0 && (reinterpret_cast<uintptr_t>(rel_w) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(rel_t) & 15) == 0, "relpos: tables must be 16-byte aligned");
  const int Lq = qt * qh * qw, Lk = kt * kh * kw;
  relpos::Geo g{qt, qh, qw, kt, kh, kw, 1.0f / scale};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  dim3 grid((unsigned)((Lq + relpos::ROWS - 1) / relpos::ROWS), (unsigned)BH);
  const int64_t total = (int64_t)BH * Lk * relpos::E;
  const unsigned kblocks = (unsigned)((total + 255) / 256);
  if (dtype == MVIT_F32) {
    relpos::q_tables_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float *>(q), rel_h, rel_w, rel_t,
                                                         static_cast<float *>(q_ext), Lq, g);
    relpos::k_onehot_kernel<float><<<kblocks, 256, 0, st>>>(static_cast<float *>(k_ext), total, Lk, g);
  } else {
    relpos::q_tables_kernel<bf16><<<grid, 256, 0, st>>>(static_cast<const bf16 *>(q), rel_h, rel_w, rel_t,
                                                        static_cast<bf16 *>(q_ext), Lq, g);
    relpos::k_onehot_kernel<bf16><<<kblocks, 256, 0, st>>>(static_cast<bf16 *>(k_ext), total, Lk, g);
  }
  MVIT_LAUNCH_OK("relpos_operands");
  return 0;
}
