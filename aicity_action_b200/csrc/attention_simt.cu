// CUDA-core pooling attention: softmax(q·kᵀ·scale)·v (+ q), flash-style (scores never leave the SM).
// fp32 parity path and on-device cross-check for the tcgen05 kernel (attention_tc.cu).
// Replaces attention.py:267-279 (QKᵀ, *scale, softmax, @V, head-merge transpose, +q).
#include "attention.cuh"

namespace mvit {

constexpr int D = 96;
constexpr int BQ = 32, BKV = 32;

constexpr int E = 64;          // relative-position contraction columns (REL)

// REL: default-off relative-position operand: the dot product runs over 96 + 64 columns,
// scores = scale * (q.k + q_ext.k_ext); shared memory is carved from the dynamic allocation in that case.
template <typename T, bool REL>
__global__ void __launch_bounds__(128) attention_simt_kernel(AttnArgs a) {
  __shared__ float sQ[BQ][D + 1];
  __shared__ float sK[BKV][D + 1];
  __shared__ float sV[BKV][D];
  __shared__ float sP[BQ][BKV + 1];
  extern __shared__ float s_ext[];                      // REL: sQe[BQ][E + 1] | sKe[BKV][E + 1]
  float (*sQe)[E + 1] = reinterpret_cast<float (*)[E + 1]>(s_ext);
  float (*sKe)[E + 1] = reinterpret_cast<float (*)[E + 1]>(s_ext + BQ * (E + 1));
  const int tid = threadIdx.x;
  const int bh = blockIdx.y;
  const int b = bh / a.heads, head = bh % a.heads;
  const int q0 = blockIdx.x * BQ;
  const T *q = static_cast<const T *>(a.q) + (int64_t)bh * a.Lq * D;
  const T *k = static_cast<const T *>(a.k) + (int64_t)bh * a.Lk * D;
  const T *v = static_cast<const T *>(a.v) + (int64_t)bh * a.Lk * D;
  for (int i = tid; i < BQ * D; i += 128) {
    const int r = i / D, c = i % D;
    sQ[r][c] = (q0 + r < a.Lq) ? to_f32(q[(int64_t)(q0 + r) * D + c]) : 0.f;
  }
  if constexpr (REL) {
    const T *qe = static_cast<const T *>(a.q_ext) + (int64_t)bh * a.Lq * E;
    for (int i = tid; i < BQ * E; i += 128) {
      const int r = i / E, c = i % E;
      sQe[r][c] = (q0 + r < a.Lq) ? to_f32(qe[(int64_t)(q0 + r) * E + c]) : 0.f;
    }
  }
  const int r = tid / 4, sub = tid % 4;
  float o[24];
#pragma unroll
  for (int i = 0; i < 24; ++i) o[i] = 0.f;
  float m_run = -INFINITY, l_run = 0.f;
  for (int k0 = 0; k0 < a.Lk; k0 += BKV) {
    __syncthreads();
    for (int i = tid; i < BKV * D; i += 128) {
      const int rr = i / D, c = i % D;
      const bool ok = k0 + rr < a.Lk;
      sK[rr][c] = ok ? to_f32(k[(int64_t)(k0 + rr) * D + c]) : 0.f;
      sV[rr][c] = ok ? to_f32(v[(int64_t)(k0 + rr) * D + c]) : 0.f;
    }
    if constexpr (REL) {
      const T *ke = static_cast<const T *>(a.k_ext) + (int64_t)bh * a.Lk * E;
      for (int i = tid; i < BKV * E; i += 128) {
        const int rr = i / E, c = i % E;
        sKe[rr][c] = (k0 + rr < a.Lk) ? to_f32(ke[(int64_t)(k0 + rr) * E + c]) : 0.f;
      }
    }
    __syncthreads();
    float s[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = 0.f;
    for (int c = 0; c < D; ++c) {
      const float qv = sQ[r][c];
#pragma unroll
      for (int j = 0; j < 8; ++j) s[j] = fmaf(qv, sK[sub + 4 * j][c], s[j]);
    }
    if constexpr (REL) {
      for (int c = 0; c < E; ++c) {
        const float qv = sQe[r][c];
#pragma unroll
        for (int j = 0; j < 8; ++j) s[j] = fmaf(qv, sKe[sub + 4 * j][c], s[j]);
      }
    }
    float tmax = -INFINITY;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s[j] = (k0 + sub + 4 * j < a.Lk) ? s[j] * a.scale : -INFINITY;
      tmax = fmaxf(tmax, s[j]);
    }
    tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, 1));
    tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, 2));
    const float m_new = fmaxf(m_run, tmax);      // finite: every tile has >= 1 valid key
    const float alpha = expf(m_run - m_new);     // exp(-inf) = 0 on the first tile
    float psum = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float pj = expf(s[j] - m_new);
      psum += pj;
      sP[r][sub + 4 * j] = pj;
    }
    psum += __shfl_xor_sync(0xffffffffu, psum, 1);
    psum += __shfl_xor_sync(0xffffffffu, psum, 2);
    l_run = l_run * alpha + psum;
    m_run = m_new;
    __syncwarp();  // the 4 threads of a row live in the same warp
#pragma unroll
    for (int i = 0; i < 24; ++i) o[i] *= alpha;
    for (int c = 0; c < BKV; ++c) {
      const float pv = sP[r][c];
#pragma unroll
      for (int i = 0; i < 24; ++i) o[i] = fmaf(pv, sV[c][sub + 4 * i], o[i]);
    }
  }
  if (q0 + r < a.Lq) {
    const float inv = 1.0f / l_run;
    T *out = static_cast<T *>(a.out) + ((int64_t)b * a.Lq + q0 + r) * (a.heads * D) + head * D;
#pragma unroll
    for (int i = 0; i < 24; ++i) {
      float val = o[i] * inv;
      if (a.add_q) val += sQ[r][sub + 4 * i];
      out[sub + 4 * i] = from_f32<T>(val);
    }
    if (a.lse && sub == 0) a.lse[(int64_t)bh * a.Lq + q0 + r] = m_run + logf(l_run);
  }
}

int attention_simt(const AttnArgs &a, int dtype, cudaStream_t st) {
  dim3 grid((unsigned)((a.Lq + BQ - 1) / BQ), (unsigned)(a.B * a.heads));
  MVIT_REQUIRE(grid.y < 65536, "attention: B*heads too large");
  if (a.q_ext) {
    constexpr size_t ext = (size_t)(BQ + BKV) * (E + 1) * sizeof(float);   // static 41.3 KB + 16.6 KB dynamic: opt in
    if (dtype == MVIT_F32) {
      MVIT_SMEM_OPT_IN((attention_simt_kernel<float, true>), ext);
      attention_simt_kernel<float, true><<<grid, 128, ext, st>>>(a);
    } else {
      MVIT_SMEM_OPT_IN((attention_simt_kernel<bf16, true>), ext);
      attention_simt_kernel<bf16, true><<<grid, 128, ext, st>>>(a);
    }
  } else if (dtype == MVIT_F32) attention_simt_kernel<float, false><<<grid, 128, 0, st>>>(a);
  else attention_simt_kernel<bf16, false><<<grid, 128, 0, st>>>(a);
  MVIT_LAUNCH_OK("attention(simt)");
  return 0;
}

}  // namespace mvit

static int attention_entry(const void *q, const void *k, const void *v, const void *q_ext, const void *k_ext, void *out,
                           float *lse, int B, int heads, int Lq, int Lk, int d, float scale, int add_q_residual, int dtype,
                           int impl, void *stream);

extern "C" int mvit_attention_fwd(const void *q, const void *k, const void *v, void *out, float *lse,
                                  int B, int heads, int Lq, int Lk, int d, float scale,
                                  int add_q_residual, int dtype, int impl, void *stream) {
  return attention_entry(q, k, v, nullptr, nullptr, out, lse, B, heads, Lq, Lk, d, scale, add_q_residual, dtype, impl, stream);
}

extern "C" int mvit_attention_rel_fwd(const void *q, const void *k, const void *v, const void *q_ext, const void *k_ext,
                                      void *out, float *lse, int B, int heads, int Lq, int Lk, int d, float scale,
                                      int add_q_residual, int dtype, int impl, void *stream) {
  MVIT_REQUIRE(q_ext && k_ext, "attention_rel: q_ext / k_ext are NULL (use mvit_attention_fwd)");
  return attention_entry(q, k, v, q_ext, k_ext, out, lse, B, heads, Lq, Lk, d, scale, add_q_residual, dtype, impl, stream);
}

static int attention_entry(const void *q, const void *k, const void *v, const void *q_ext, const void *k_ext, void *out,
                           float *lse, int B, int heads, int Lq, int Lk, int d, float scale, int add_q_residual, int dtype,
                           int impl, void *stream) {
  using namespace mvit;
  MVIT_REQUIRE(q && k && v && out, "attention: null pointer");
  MVIT_REQUIRE(B >= 0 && heads > 0 && Lq > 0 && Lk > 0, "attention: bad shape");
  MVIT_REQUIRE(d == 96, "attention: head_dim %d unsupported (every Aicity MViT config uses 96)", d);
  MVIT_REQUIRE(dtype == MVIT_F32 || dtype == MVIT_BF16, "attention: unknown dtype %d", dtype);
  if (B == 0) return 0;
  AttnArgs a{q, k, v, out, lse, B, heads, Lq, Lk, scale, add_q_residual ? 1 : 0};
  a.q_ext = q_ext;
  a.k_ext = k_ext;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  bool use_tc = false;
  if (impl == MVIT_IMPL_TCGEN05 || (impl == MVIT_IMPL_AUTO && dtype == MVIT_BF16)) {
    MVIT_REQUIRE(dtype == MVIT_BF16, "attention: the tcgen05 path is bf16 only");
    const char *why = "";
    if (attention_tc_supported(a, &why)) use_tc = true;
    else MVIT_REQUIRE(impl == MVIT_IMPL_AUTO, "attention: tcgen05 path rejected: %s", why);
  }
  if (use_tc) return attention_tc(a, st);
  return attention_simt(a, dtype, st);
}
