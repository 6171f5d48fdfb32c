// attention_pool (attention.py:12-83): depthwise Conv3d / MaxPool3d / AvgPool3d over the (T,H,W) token
// grid, fused with the LayerNorm that follows it (attention.py:66-67), operating on channels-last
// tokens in place of the reference's [B*h, d, T, H, W] round trip (attention.py:34-36, 58-60).
//
// Generic kernel (any kernel size / stride / d % 32 == 0, cls token, three modes): one warp per output
// token*head, lane owns channels {lane, lane+32, ...} so every global access of a warp is one contiguous
// d*sizeof(T) segment.  The tuned stride-(1,s,s) 3x3x3 path lives in pool_tiled.cu.
#include "pool.cuh"

namespace mvit {

template <typename T, int NC, int MODE>
__global__ void __launch_bounds__(256) pool_generic_kernel(const T *__restrict__ in,
                                                           const float *__restrict__ weight,
                                                           const float *__restrict__ gamma,
                                                           const float *__restrict__ beta,
                                                           T *__restrict__ out, PoolParams p) {
  extern __shared__ float w_s[];  // [taps][d] (CONV only)
  const int taps = p.kt * p.kh * p.kw;
  if (MODE == MVIT_POOL_CONV) {
    for (int i = threadIdx.x; i < taps * p.d; i += blockDim.x) {
      const int tap = i / p.d, c = i - tap * p.d;
      w_s[i] = weight[c * taps + tap];
    }
    __syncthreads();
  }
  const int lane = threadIdx.x & 31;
  const int Lo = p.To * p.Ho * p.Wo + p.has_cls;
  const int64_t total = (int64_t)p.B * Lo * p.heads;
  const int64_t o = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (o >= total) return;
  const int head = (int)(o % p.heads);
  const int64_t bl = o / p.heads;
  const int lo = (int)(bl % Lo);
  const int b = (int)(bl / Lo);
  const T *src = in + b * p.in_bs + head * p.in_hs;
  float acc[NC];
  if (p.has_cls && lo == 0) {
#pragma unroll
    for (int j = 0; j < NC; ++j) acc[j] = to_f32(src[lane + 32 * j]);
  } else {
    const int l = lo - p.has_cls;
    const int wo = l % p.Wo, ho = (l / p.Wo) % p.Ho, to = l / (p.Wo * p.Ho);
    const int t0 = to * p.st - p.pt, h0 = ho * p.sh - p.ph, w0 = wo * p.sw - p.pw;
#pragma unroll
    for (int j = 0; j < NC; ++j) acc[j] = (MODE == MVIT_POOL_MAX) ? -INFINITY : 0.f;
    for (int a = 0; a < p.kt; ++a) {
      const int t = t0 + a;
      if (t < 0 || t >= p.T) continue;
      for (int bq = 0; bq < p.kh; ++bq) {
        const int h = h0 + bq;
        if (h < 0 || h >= p.H) continue;
        for (int c = 0; c < p.kw; ++c) {
          const int w = w0 + c;
          if (w < 0 || w >= p.W) continue;
          const T *px = src + ((int64_t)((t * p.H + h) * p.W + w) + p.has_cls) * p.in_ls;
          const float *pw = w_s + ((a * p.kh + bq) * p.kw + c) * p.d;
#pragma unroll
          for (int j = 0; j < NC; ++j) {
            const float v = to_f32(px[lane + 32 * j]);
            if (MODE == MVIT_POOL_CONV) acc[j] = fmaf(v, pw[lane + 32 * j], acc[j]);
            else if (MODE == MVIT_POOL_MAX) acc[j] = fmaxf(acc[j], v);
            else acc[j] += v;
          }
        }
      }
    }
    if (MODE == MVIT_POOL_AVG) {
      const float inv = 1.0f / (float)taps;  // count_include_pad=True (torch default)
#pragma unroll
      for (int j = 0; j < NC; ++j) acc[j] *= inv;
    }
  }
  if (p.has_ln) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < NC; ++j) s += acc[j];
    const float mean = warp_sum(s) / (float)p.d;
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < NC; ++j) {
      const float dlt = acc[j] - mean;
      ss += dlt * dlt;
    }
    const float rstd = rsqrtf(warp_sum(ss) / (float)p.d + p.eps);
#pragma unroll
    for (int j = 0; j < NC; ++j)
      acc[j] = (acc[j] - mean) * rstd * gamma[lane + 32 * j] + beta[lane + 32 * j];
  }
  T *dst = out + b * p.out_bs + (int64_t)lo * p.out_ls + head * p.out_hs;
#pragma unroll
  for (int j = 0; j < NC; ++j) dst[lane + 32 * j] = from_f32<T>(acc[j]);
}

template <typename T, int NC>
static int launch_generic(const void *in, const float *w, const float *g, const float *b, void *out,
                          const PoolParams &p, int mode, cudaStream_t st) {
  const int Lo = p.To * p.Ho * p.Wo + p.has_cls;
  const int64_t total = (int64_t)p.B * Lo * p.heads;
  const int threads = 256;
  const int64_t blocks = (total + 7) / 8;
  MVIT_REQUIRE(blocks < ((int64_t)1 << 31), "attention_pool: grid too large");
  const size_t smem = mode == MVIT_POOL_CONV ? (size_t)p.kt * p.kh * p.kw * p.d * sizeof(float) : 0;
  MVIT_REQUIRE(smem <= 48 * 1024, "attention_pool: conv kernel too large for the generic path");
  const T *pi = static_cast<const T *>(in);
  T *po = static_cast<T *>(out);
  if (mode == MVIT_POOL_CONV)
    pool_generic_kernel<T, NC, MVIT_POOL_CONV><<<(unsigned)blocks, threads, smem, st>>>(pi, w, g, b, po, p);
  else if (mode == MVIT_POOL_MAX)
    pool_generic_kernel<T, NC, MVIT_POOL_MAX><<<(unsigned)blocks, threads, 0, st>>>(pi, w, g, b, po, p);
  else
    pool_generic_kernel<T, NC, MVIT_POOL_AVG><<<(unsigned)blocks, threads, 0, st>>>(pi, w, g, b, po, p);
  MVIT_LAUNCH_OK("attention_pool(generic)");
  return 0;
}

template <typename T>
static int dispatch_generic(const void *in, const float *w, const float *g, const float *b, void *out,
                            const PoolParams &p, int mode, cudaStream_t st) {
  switch (p.d / 32) {
    case 1: return launch_generic<T, 1>(in, w, g, b, out, p, mode, st);
    case 2: return launch_generic<T, 2>(in, w, g, b, out, p, mode, st);
    case 3: return launch_generic<T, 3>(in, w, g, b, out, p, mode, st);
    case 4: return launch_generic<T, 4>(in, w, g, b, out, p, mode, st);
  }
  MVIT_REQUIRE(false, "attention_pool: head_dim %d unsupported (need 32..128, multiple of 32)", p.d);
}


// Channels-last MaxPool3d for the skip path (attention.py:427-432: MaxPool3d([1,3,3],[1,2,2],[0,1,1]) on [B, L, C]):
// one thread = 8 consecutive channels (one 16-byte vector) of one output token; the window's vectors are read
// with coalesced 16-byte loads and reduced with packed bf16 max.  Memory-bound: every input byte is read ~2.25x
// through L1/L2, once from HBM.
__global__ void __launch_bounds__(256) maxpool_tokens_bf16_kernel(const bf16 *__restrict__ in, bf16 *__restrict__ out,
                                                                 PoolParams p, int C) {
  const int vecs = C / 8;
  const int64_t total = (int64_t)p.B * p.To * p.Ho * p.Wo * vecs;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int v = (int)(i % vecs);
    int64_t r = i / vecs;
    const int wo = (int)(r % p.Wo); r /= p.Wo;
    const int ho = (int)(r % p.Ho); r /= p.Ho;
    const int to = (int)(r % p.To);
    const int b = (int)(r / p.To);
    const int t0 = to * p.st - p.pt, h0 = ho * p.sh - p.ph, w0 = wo * p.sw - p.pw;
    const __nv_bfloat162 ninf = __floats2bfloat162_rn(-INFINITY, -INFINITY);
    __nv_bfloat162 m[4] = {ninf, ninf, ninf, ninf};
    const bf16 *src = in + (int64_t)b * p.in_bs + v * 8;
    for (int a = 0; a < p.kt; ++a) {
      const int t = t0 + a;
      if (t < 0 || t >= p.T) continue;
      for (int bq = 0; bq < p.kh; ++bq) {
        const int h = h0 + bq;
        if (h < 0 || h >= p.H) continue;
        for (int c = 0; c < p.kw; ++c) {
          const int w = w0 + c;
          if (w < 0 || w >= p.W) continue;
          const uint4 x = *reinterpret_cast<const uint4 *>(src + (int64_t)((t * p.H + h) * p.W + w) * p.in_ls);
          const __nv_bfloat162 *xv = reinterpret_cast<const __nv_bfloat162 *>(&x);
#pragma unroll
          for (int k = 0; k < 4; ++k) m[k] = __hmax2(m[k], xv[k]);
        }
      }
    }
    uint4 o;
    __nv_bfloat162 *ov = reinterpret_cast<__nv_bfloat162 *>(&o);
#pragma unroll
    for (int k = 0; k < 4; ++k) ov[k] = m[k];
    *reinterpret_cast<uint4 *>(out + (int64_t)b * p.out_bs + (int64_t)((to * p.Ho + ho) * p.Wo + wo) * p.out_ls + v * 8) = o;
  }
}

// The shipped geometry (kernel [1,3,3], stride [1,2,2], padding [0,1,1]) without per-element 64-bit index arithmetic:
// a block owns one output row (b, t, ho), a thread walks its (wo, 16-byte channel vector) slots with 32-bit indices and
// issues the nine window loads back to back.
__global__ void __launch_bounds__(256) maxpool_133_s122_bf16_kernel(const bf16 *__restrict__ in, bf16 *__restrict__ out,
                                                                   PoolParams p, int C) {
  const int vecs = C >> 3;
  int r = blockIdx.x;
  const int ho = r % p.Ho; r /= p.Ho;
  const int t = r % p.T;
  const int b = r / p.T;
  const int h0 = 2 * ho - 1;
  const uint4 *src = reinterpret_cast<const uint4 *>(in + (int64_t)b * p.in_bs) + (int64_t)t * p.H * p.W * vecs;
  uint4 *dst = reinterpret_cast<uint4 *>(out + (int64_t)b * p.out_bs) + ((int64_t)t * p.Ho + ho) * p.Wo * vecs;
  const int n = p.Wo * vecs;
  for (int i = threadIdx.x; i < n; i += 256) {
    const int wo = i / vecs, v = i - wo * vecs;
    const int w0 = 2 * wo - 1;
    // a tap outside the image is clamped onto the nearest in-image tap OF THE SAME WINDOW, which leaves the maximum
    // unchanged (the reference pads with -inf): nine unconditional, independent loads
    uint4 x[9];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const int h = min(max(h0 + a, 0), p.H - 1);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int w = min(max(w0 + c, 0), p.W - 1);
        x[a * 3 + c] = __ldg(src + (h * p.W + w) * vecs + v);
      }
    }
    __nv_bfloat162 m[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) m[k] = reinterpret_cast<const __nv_bfloat162 *>(&x[0])[k];
#pragma unroll
    for (int q = 1; q < 9; ++q)
#pragma unroll
      for (int k = 0; k < 4; ++k) m[k] = __hmax2(m[k], reinterpret_cast<const __nv_bfloat162 *>(&x[q])[k]);
    uint4 o;
#pragma unroll
    for (int k = 0; k < 4; ++k) reinterpret_cast<__nv_bfloat162 *>(&o)[k] = m[k];
    dst[i] = o;
  }
}

// returns 1 when it does not apply
static int maxpool_tokens_try(const void *in, void *out, const PoolParams &p, int mode, int dtype, cudaStream_t st) {
  if (mode != MVIT_POOL_MAX || dtype != MVIT_BF16 || p.has_cls || p.has_ln) return 1;
  const int C = p.heads * p.d;
  // only the plain [B, L, C] -> [B, L', C] layout (heads are consecutive d-channel groups of a token)
  if (p.in_hs != p.d || p.out_hs != p.d || p.in_ls != C || p.out_ls != C || C % 8 != 0) return 1;
  if ((reinterpret_cast<uintptr_t>(in) & 15) || (reinterpret_cast<uintptr_t>(out) & 15) || (p.in_bs % 8) || (p.out_bs % 8)) return 1;
  if (p.kt == 1 && p.kh == 3 && p.kw == 3 && p.st == 1 && p.sh == 2 && p.sw == 2 && p.pt == 0 && p.ph == 1 && p.pw == 1 &&
      (int64_t)p.B * p.T * p.Ho < ((int64_t)1 << 31) && (int64_t)p.H * p.W * (C / 8) < ((int64_t)1 << 31)) {
    maxpool_133_s122_bf16_kernel<<<(unsigned)(p.B * p.T * p.Ho), 256, 0, st>>>(static_cast<const bf16 *>(in),
                                                                              static_cast<bf16 *>(out), p, C);
    MVIT_LAUNCH_OK("attention_pool(maxpool 1x3x3 / 1x2x2)");
    return 0;
  }
  const int64_t total = (int64_t)p.B * p.To * p.Ho * p.Wo * (C / 8);
  const unsigned blocks = (unsigned)std::min<int64_t>((total + 255) / 256, (int64_t)num_sms() * 32);
  maxpool_tokens_bf16_kernel<<<blocks, 256, 0, st>>>(static_cast<const bf16 *>(in), static_cast<bf16 *>(out), p, C);
  MVIT_LAUNCH_OK("attention_pool(maxpool tokens)");
  return 0;
}

}  // namespace mvit

static int pool_fwd_impl(const void *in, int64_t in_bs, int64_t in_ls, int64_t in_hs, const float *weight,
                         const float *gamma, const float *beta, void *out, int64_t out_bs, int64_t out_ls, int64_t out_hs,
                         void *pre_ln_out, int B, int heads, int d, int T, int H, int W, int kt, int kh, int kw, int st,
                         int sh, int sw, int mode, int has_cls, float eps, int dtype, void *stream) {
  using namespace mvit;
  MVIT_REQUIRE(in && out, "attention_pool: null pointer");
  MVIT_REQUIRE(B >= 0 && heads > 0 && d > 0 && T > 0 && H > 0 && W > 0, "attention_pool: bad shape");
  MVIT_REQUIRE(kt > 0 && kh > 0 && kw > 0 && st > 0 && sh > 0 && sw > 0, "attention_pool: bad kernel/stride");
  MVIT_REQUIRE(mode == MVIT_POOL_CONV || mode == MVIT_POOL_MAX || mode == MVIT_POOL_AVG,
               "attention_pool: unknown mode %d", mode);
  MVIT_REQUIRE(mode != MVIT_POOL_CONV || weight, "attention_pool: conv mode needs a weight");
  MVIT_REQUIRE((gamma == nullptr) == (beta == nullptr), "attention_pool: gamma/beta must both be set or NULL");
  MVIT_REQUIRE(d % 32 == 0 && d <= 128, "attention_pool: head_dim %d unsupported (multiple of 32, <= 128)", d);
  MVIT_REQUIRE(dtype == MVIT_F32 || dtype == MVIT_BF16, "attention_pool: unknown dtype %d", dtype);
  if (B == 0) return 0;
  PoolParams p;
  p.in_bs = in_bs; p.in_ls = in_ls; p.in_hs = in_hs;
  p.out_bs = out_bs; p.out_ls = out_ls; p.out_hs = out_hs;
  p.B = B; p.heads = heads; p.d = d; p.T = T; p.H = H; p.W = W;
  p.kt = kt; p.kh = kh; p.kw = kw; p.st = st; p.sh = sh; p.sw = sw;
  p.pt = kt / 2; p.ph = kh / 2; p.pw = kw / 2;
  p.To = (T + 2 * p.pt - kt) / st + 1;
  p.Ho = (H + 2 * p.ph - kh) / sh + 1;
  p.Wo = (W + 2 * p.pw - kw) / sw + 1;
  MVIT_REQUIRE(p.To > 0 && p.Ho > 0 && p.Wo > 0, "attention_pool: empty output");
  p.has_cls = has_cls ? 1 : 0;
  p.has_ln = gamma ? 1 : 0;
  p.eps = eps;
  p.pre_out = pre_ln_out;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int r = pool_tiled_try(in, weight, gamma, beta, out, p, mode, dtype, s);
  if (r <= 0) return r;
  MVIT_REQUIRE(pre_ln_out == nullptr, "attention_pool: the pre-LayerNorm output is only produced by the tuned kernel "
                                      "(3x3x3 depthwise conv, head_dim 96, stride (1,s,s), s in {1,2,4,8}, 16-byte aligned)");
  r = maxpool_tokens_try(in, out, p, mode, dtype, s);
  if (r <= 0) return r;
  if (dtype == MVIT_F32) return dispatch_generic<float>(in, weight, gamma, beta, out, p, mode, s);
  return dispatch_generic<bf16>(in, weight, gamma, beta, out, p, mode, s);
}

extern "C" int mvit_attention_pool_fwd(const void *in, int64_t in_bs, int64_t in_ls, int64_t in_hs,
                                       const float *weight, const float *gamma, const float *beta,
                                       void *out, int64_t out_bs, int64_t out_ls, int64_t out_hs, int B,
                                       int heads, int d, int T, int H, int W, int kt, int kh, int kw,
                                       int st, int sh, int sw, int mode, int has_cls, float eps,
                                       int dtype, void *stream) {
  return pool_fwd_impl(in, in_bs, in_ls, in_hs, weight, gamma, beta, out, out_bs, out_ls, out_hs, nullptr, B, heads, d, T, H,
                       W, kt, kh, kw, st, sh, sw, mode, has_cls, eps, dtype, stream);
}

extern "C" int mvit_attention_pool_fwd_save(const void *in, int64_t in_bs, int64_t in_ls, int64_t in_hs,
                                            const float *weight, const float *gamma, const float *beta, void *out,
                                            int64_t out_bs, int64_t out_ls, int64_t out_hs, void *pre_ln_out, int B,
                                            int heads, int d, int T, int H, int W, int kt, int kh, int kw, int st,
                                            int sh, int sw, float eps, int dtype, void *stream) {
  MVIT_REQUIRE(pre_ln_out && gamma && beta, "attention_pool_fwd_save: needs the LayerNorm and the second output");
  return pool_fwd_impl(in, in_bs, in_ls, in_hs, weight, gamma, beta, out, out_bs, out_ls, out_hs, pre_ln_out, B, heads, d, T,
                       H, W, kt, kh, kw, st, sh, sw, MVIT_POOL_CONV, 0, eps, dtype, stream);
}
