// Small memory-bound kernels around the block stack: positional-embedding add and mean-pool + head.
#include "linear.cuh"

namespace mvit {

// x[b, t*HW + s, c] = src[...] + pos_spatial[s, c] + pos_temporal[t, c]
// (video_model_builder.py:1206-1223: spatial.repeat(1,T,1) + repeat_interleave(temporal, HW))
template <typename TS, typename TD>
__global__ void pos_embed_add_kernel(const TS *__restrict__ src, const float *__restrict__ ps,
                                     const float *__restrict__ pt, TD *__restrict__ dst, int64_t total,
                                     int T, int HW, int C) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int c = (int)(i % C);
    const int64_t tok = (i / C) % ((int64_t)T * HW);
    const int s = (int)(tok % HW), t = (int)(tok / HW);
    dst[i] = from_f32<TD>(to_f32(src[i]) + (ps[(int64_t)s * C + c] + pt[(int64_t)t * C + c]));
  }
}

// stage 1 (optional): partial[b][chunk][c] = sum of x over a chunk of kHeadChunk tokens (coalesced over c)
constexpr int kHeadChunk = 32;
template <typename T>
__global__ void __launch_bounds__(256) token_partial_sum_kernel(const T *__restrict__ x, float *__restrict__ partial,
                                                                int L, int C, int chunks) {
  const int b = blockIdx.y, ck = blockIdx.x;
  const int l0 = ck * kHeadChunk, l1 = min(l0 + kHeadChunk, L);
  const T *px = x + (int64_t)b * L * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int l = l0; l < l1; ++l) s += to_f32(px[(int64_t)l * C + c]);
    partial[((int64_t)b * chunks + ck) * C + c] = s;
  }
}

// one CTA per sample: feat = mean over tokens (or over the stage-1 partial sums), logits = feat·Wᵀ + b,
// optional softmax.  Summation order is fixed (no atomics) so results are run-to-run reproducible.
template <typename T>
__global__ void __launch_bounds__(256) mean_head_kernel(const T *__restrict__ x, const float *__restrict__ partial,
                                                        int chunks, const float *__restrict__ w,
                                                        const float *__restrict__ bias,
                                                        float *__restrict__ feat_out, float *__restrict__ out,
                                                        int L, int C, int NCLS, int apply_softmax) {
  extern __shared__ float sm[];  // feat[C] + logits[NCLS]
  float *feat = sm, *logits = sm + C;
  const int b = blockIdx.x;
  const T *px = x + (int64_t)b * L * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    if (partial) {
      for (int k = 0; k < chunks; ++k) s += partial[((int64_t)b * chunks + k) * C + c];
    } else {
      for (int l = 0; l < L; ++l) s += to_f32(px[(int64_t)l * C + c]);
    }
    feat[c] = s / (float)L;
    if (feat_out) feat_out[(int64_t)b * C + c] = feat[c];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int n = warp; n < NCLS; n += blockDim.x >> 5) {
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s = fmaf(feat[c], w[(int64_t)n * C + c], s);
    s = warp_sum(s);
    if (lane == 0) logits[n] = s + (bias ? bias[n] : 0.f);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (apply_softmax) {
      float m = -INFINITY;
      for (int n = 0; n < NCLS; ++n) m = fmaxf(m, logits[n]);
      float z = 0.f;
      for (int n = 0; n < NCLS; ++n) z += expf(logits[n] - m);
      for (int n = 0; n < NCLS; ++n) out[(int64_t)b * NCLS + n] = expf(logits[n] - m) / z;
    } else {
      for (int n = 0; n < NCLS; ++n) out[(int64_t)b * NCLS + n] = logits[n];
    }
  }
}

// uint8 THWC -> normalised CTHW (module_wrapper.py:332-346: /255, transpose, (x-mean)/std)
template <typename T>
__global__ void preprocess_u8_kernel(const uint8_t *__restrict__ in, T *__restrict__ out, int64_t pixels_per_clip,
                                     int64_t total_pixels, float mean, float stdv) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_pixels; i += stride) {
    const int64_t b = i / pixels_per_clip, pix = i - b * pixels_per_clip;
    const uint8_t *px = in + i * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float v = __fdiv_rn(__fsub_rn(__fdiv_rn((float)px[c], 255.0f), mean), stdv);
      out[(b * 3 + c) * pixels_per_clip + pix] = from_f32<T>(v);
    }
  }
}


// im2col for Conv3d.  One thread produces 8 consecutive columns of one patch row (a single 16-byte store for
// bf16, two for fp32).  Column k -> (c, a, b, d) is decoded once per CTA into a shared-memory table holding the
// element offset inside a [C, T, H, W] clip and the (a, b, d) tap coordinates for the border test.
template <typename T>
__global__ void __launch_bounds__(256) im2col3d_kernel(const T *__restrict__ x, T *__restrict__ out, int B, int C, int Ti,
                                                       int H, int W, int kt, int kh, int kw, int st, int sh, int sw,
                                                       int pt, int ph, int pw, int To, int Ho, int Wo, int Kp) {
  extern __shared__ int2 lut[];                 // [Kp]: .x = element offset (or -1 for padding columns), .y = a | b<<8 | d<<16
  const int kreal = C * kt * kh * kw;
  for (int k = threadIdx.x; k < Kp; k += blockDim.x) {
    if (k < kreal) {
      const int d = k % kw, bq = (k / kw) % kh, a = (k / (kw * kh)) % kt, c = k / (kw * kh * kt);
      lut[k] = make_int2(((c * Ti + a) * H + bq) * W + d, a | (bq << 8) | (d << 16));
    } else {
      lut[k] = make_int2(-1, 0);
    }
  }
  __syncthreads();
  // one thread = 8 columns of one patch row; a CTA covers blockDim.x / vpt consecutive tokens (vpt = vectors per
  // token rounded up to a power of two) so the token decode is a handful of 32-bit operations per thread
  const int vecs = Kp / 8;
  int vpt = 1;
  while (vpt < vecs) vpt <<= 1;
  const int tok_per_cta = blockDim.x / vpt;
  const int v = threadIdx.x & (vpt - 1);
  const int64_t n_tok = (int64_t)B * To * Ho * Wo;
  const int64_t clip_elems = (int64_t)C * Ti * H * W;
  for (int64_t m = (int64_t)blockIdx.x * tok_per_cta + threadIdx.x / vpt; m < n_tok; m += (int64_t)gridDim.x * tok_per_cta) {
    if (v >= vecs) continue;
    const int per_clip = To * Ho * Wo;
    const int b = (int)(m / per_clip);
    int r = (int)(m - (int64_t)b * per_clip);
    const int wo = r % Wo; r /= Wo;
    const int ho = r % Ho;
    const int to = r / Ho;
    const int t0 = to * st - pt, h0 = ho * sh - ph, w0 = wo * sw - pw;
    const T *base = x + b * clip_elems + ((int64_t)t0 * H + h0) * W + w0;     // may point before the clip: only
    const bool inside = t0 >= 0 && t0 + kt <= Ti && h0 >= 0 && h0 + kh <= H && w0 >= 0 && w0 + kw <= W;   // offset it
    T vals[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int2 l = lut[v * 8 + e];
      bool ok = l.x >= 0;
      if (!inside && ok) {
        const int t = t0 + (l.y & 0xff), h = h0 + ((l.y >> 8) & 0xff), w = w0 + (l.y >> 16);
        ok = t >= 0 && t < Ti && h >= 0 && h < H && w >= 0 && w < W;
      }
      vals[e] = ok ? base[l.x] : from_f32<T>(0.f);
    }
    T *dst = out + m * Kp + v * 8;
    if constexpr (sizeof(T) == 2) {
      *reinterpret_cast<uint4 *>(dst) = *reinterpret_cast<const uint4 *>(vals);
    } else {
      *reinterpret_cast<uint4 *>(dst) = *reinterpret_cast<const uint4 *>(vals);
      *reinterpret_cast<uint4 *>(dst + 4) = *reinterpret_cast<const uint4 *>(vals + 4);
    }
  }
}


// Space-to-depth "fold" of a clip for the implicit-GEMM patch embedding (gemm_tc.cu, ConvGeom): the conv stride
// (st, sh, sw) is folded into channels,  folded[b, t/st, h/sh, w/sw, ((t%st*sh + h%sh)*sw + w%sw)*C + c] = x[b,c,t,h,w]
// (channels [st*sh*sw*C, Cf) are zero padding).  A CTA handles 16 consecutive output tokens of one (b, tf, hf) line:
// coalesced row loads into shared memory, coalesced 2*Cf-byte token stores.  SRC = float / bf16 channels-first clip, or
// uint8 channels-last frames with the reference normalisation ((x/255 - mean)/std, module_wrapper.py:332-346) fused in.
constexpr int kFoldTok = 16;
template <typename SRC, bool FRAMES_U8>
__global__ void __launch_bounds__(256) fold_clip_kernel(const SRC *__restrict__ x, bf16 *__restrict__ out, int B, int C, int T,
                                                        int H, int W, int st, int sh, int sw, int Cf, float mean, float stdv,
                                                        int strips) {
  extern __shared__ __align__(16) uint8_t fold_smem[];
  SRC *stage = reinterpret_cast<SRC *>(fold_smem);
  const int Tf = T / st, Hf = H / sh, Wf = W / sw;
  int bid = blockIdx.x;
  const int strip = bid % strips; bid /= strips;
  const int hf = bid % Hf; bid /= Hf;
  const int tf = bid % Tf;
  const int b = bid / Tf;
  const int wf0 = strip * kFoldTok;
  const int ntok = min(kFoldTok, Wf - wf0);
  const int RL = kFoldTok * sw * (FRAMES_U8 ? C : 1);          // staged elements per row
  const int rows = FRAMES_U8 ? st * sh : C * st * sh;
  for (int i = threadIdx.x; i < rows * RL; i += blockDim.x) {
    const int r = i / RL, col = i - r * RL;
    SRC v = SRC(0);
    if (FRAMES_U8) {                                           // row r = (ot, oh); col = (w_local, c)
      const int oh = r % sh, ot = r / sh;
      const int wl = col / C;
      if (wl < ntok * sw)
        v = x[((((int64_t)b * T + tf * st + ot) * H + hf * sh + oh) * W + wf0 * sw) * C + col];
    } else {                                                   // row r = (c, ot, oh); col = w_local
      const int oh = r % sh, ot = (r / sh) % st, c = r / (sh * st);
      if (col < ntok * sw)
        v = x[((((int64_t)b * C + c) * T + tf * st + ot) * H + hf * sh + oh) * W + wf0 * sw + col];
    }
    stage[i] = v;
  }
  __syncthreads();
  const int creal = st * sh * sw * C;
  bf16 *dst = out + ((((int64_t)b * Tf + tf) * Hf + hf) * Wf + wf0) * Cf;
  for (int i = threadIdx.x; i < ntok * Cf; i += blockDim.x) {
    const int tk = i / Cf, ch = i - tk * Cf;
    float v = 0.f;
    if (ch < creal) {
      const int c = ch % C, ow = (ch / C) % sw, oh = (ch / (C * sw)) % sh, ot = ch / (C * sw * sh);
      if (FRAMES_U8) {
        const float raw = (float)stage[(ot * sh + oh) * RL + (tk * sw + ow) * C + c];
        v = __fdiv_rn(__fsub_rn(__fdiv_rn(raw, 255.0f), mean), stdv);
      } else {
        v = to_f32(stage[((c * st + ot) * sh + oh) * RL + tk * sw + ow]);
      }
    }
    dst[i] = __float2bfloat16_rn(v);
  }
}


// Specialisation of the fold for the MViT patch embedding (C = 3, stride (2,4,4), Cf = 128, bf16 / uint8 sources):
// compile-time geometry, 16-byte staging loads, one 16-byte store (8 folded channels) per thread.
// Folded channel ch = ((ot*4 + oh)*4 + ow)*3 + c.
// TOK = folded tokens (of one image row) per CTA: a whole row where the width allows (112 @448, 56 @224) so that a
// thread has several independent 16-byte loads in flight before the barrier (with 16-token CTAs the kernel was bound by
// the latency of one load round per 7 KB of traffic: 0.31 of HBM).
template <typename SRC, bool FRAMES_U8, int TOK>
__global__ void __launch_bounds__(256) fold_clip_244x3_kernel(const SRC *__restrict__ x, bf16 *__restrict__ out, int B, int T,
                                                              int H, int W, float mean, float stdv, int strips) {
  constexpr int C = 3, ST = 2, SH = 4, SW = 4, CF = 128;
  constexpr int ROWS = FRAMES_U8 ? ST * SH : C * ST * SH;          // 8 / 24 staged rows
  constexpr int RL = TOK * SW * (FRAMES_U8 ? C : 1);               // 192 / 64 elements per row
  constexpr int VEC = 16 / (int)sizeof(SRC);                       // elements per 16-byte staging vector
  __shared__ __align__(16) SRC stage[ROWS * RL];
  const int Tf = T / ST, Hf = H / SH, Wf = W / SW;
  int bid = blockIdx.x;
  const int strip = bid % strips; bid /= strips;
  const int hf = bid % Hf; bid /= Hf;
  const int tf = bid % Tf;
  const int b = bid / Tf;
  const int wf0 = strip * TOK;
  const int ntok = min(TOK, Wf - wf0);
#pragma unroll 4
  for (int i = threadIdx.x; i < ROWS * RL / VEC; i += 256) {
    const int r = i / (RL / VEC), cv = (i % (RL / VEC)) * VEC;
    const SRC *src;
    int live_elems;
    if (FRAMES_U8) {
      const int oh = r % SH, ot = r / SH;
      src = x + ((((int64_t)b * T + tf * ST + ot) * H + hf * SH + oh) * W + wf0 * SW) * C + cv;
      live_elems = ntok * SW * C;
    } else {
      const int oh = r % SH, ot = (r / SH) % ST, c = r / (SH * ST);
      src = x + ((((int64_t)b * C + c) * T + tf * ST + ot) * H + hf * SH + oh) * W + wf0 * SW + cv;
      live_elems = ntok * SW;
    }
    uint4 v = make_uint4(0, 0, 0, 0);
    if (cv + VEC <= live_elems) v = *reinterpret_cast<const uint4 *>(src);
    *reinterpret_cast<uint4 *>(&stage[r * RL + cv]) = v;
  }
  __syncthreads();
  const int v8 = threadIdx.x & 15;                                 // group of 8 folded channels
#pragma unroll 1
  for (int tk = threadIdx.x >> 4; tk < ntok; tk += 16) {           // token
  alignas(16) bf16 o[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int ch = v8 * 8 + e;
    float val = 0.f;
    if (ch < ST * SH * SW * C) {
      const int c = ch % C, ow = (ch / C) % SW, oh = (ch / (C * SW)) % SH, ot = ch / (C * SW * SH);
      if (FRAMES_U8) {
        const float raw = (float)stage[(ot * SH + oh) * RL + (tk * SW + ow) * C + c];
        val = __fdiv_rn(__fsub_rn(__fdiv_rn(raw, 255.0f), mean), stdv);
      } else {
        val = to_f32(stage[((c * ST + ot) * SH + oh) * RL + tk * SW + ow]);
      }
    }
    o[e] = __float2bfloat16_rn(val);
  }
  bf16 *dst = out + ((((int64_t)b * Tf + tf) * Hf + hf) * Wf + wf0 + tk) * CF + v8 * 8;
  *reinterpret_cast<uint4 *>(dst) = *reinterpret_cast<const uint4 *>(o);
  }
}

}  // namespace mvit

extern "C" int mvit_pos_embed_add(const void *src, int src_dtype, const float *pos_spatial,
                                  const float *pos_temporal, void *dst, int B, int T, int HW, int C,
                                  int dtype, void *stream) {
  using namespace mvit;
  MVIT_REQUIRE(src && dst && pos_spatial && pos_temporal, "pos_embed_add: null pointer");
  MVIT_REQUIRE(B >= 0 && T > 0 && HW > 0 && C > 0, "pos_embed_add: bad shape");
  if (B == 0) return 0;
  const int64_t total = (int64_t)B * T * HW * C;
  const int threads = 256;
  const unsigned blocks = (unsigned)std::min<int64_t>((total + threads - 1) / threads, (int64_t)num_sms() * 16);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define PE(TS, TD)                                                                                     \
  pos_embed_add_kernel<TS, TD><<<blocks, threads, 0, st>>>(static_cast<const TS *>(src), pos_spatial,  \
                                                           pos_temporal, static_cast<TD *>(dst), total, T, HW, C)
  if (src_dtype == MVIT_F32 && dtype == MVIT_F32) PE(float, float);
  else if (src_dtype == MVIT_F32 && dtype == MVIT_BF16) PE(float, bf16);
  else if (src_dtype == MVIT_BF16 && dtype == MVIT_BF16) PE(bf16, bf16);
  else if (src_dtype == MVIT_BF16 && dtype == MVIT_F32) PE(bf16, float);
  else MVIT_REQUIRE(false, "pos_embed_add: unknown dtype");
#undef PE
  MVIT_LAUNCH_OK("pos_embed_add");
  return 0;
}

extern "C" size_t mvit_mean_head_workspace_floats(int B, int L, int C) {
  return (size_t)B * ((L + mvit::kHeadChunk - 1) / mvit::kHeadChunk) * C;
}

extern "C" int mvit_mean_head_fwd(const void *x, const float *w, const float *bias, float *feat_out,
                                  float *out, float *workspace, int B, int L, int C, int num_classes,
                                  int apply_softmax, int dtype, void *stream) {
  using namespace mvit;
  MVIT_REQUIRE(x && w && out, "mean_head: null pointer");
  MVIT_REQUIRE(B >= 0 && L > 0 && C > 0 && num_classes > 0, "mean_head: bad shape");
  MVIT_REQUIRE(dtype == MVIT_F32 || dtype == MVIT_BF16, "mean_head: unknown dtype");
  if (B == 0) return 0;
  const size_t smem = (size_t)(C + num_classes) * sizeof(float);
  MVIT_REQUIRE(smem <= 48 * 1024, "mean_head: C + classes too large");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int chunks = (L + kHeadChunk - 1) / kHeadChunk;
  const bool two_stage = workspace != nullptr && chunks > 1;
  if (dtype == MVIT_F32) {
    if (two_stage) token_partial_sum_kernel<float><<<dim3(chunks, B), 256, 0, st>>>(static_cast<const float *>(x), workspace, L, C, chunks);
    mean_head_kernel<float><<<B, 256, smem, st>>>(static_cast<const float *>(x), two_stage ? workspace : nullptr, chunks, w, bias, feat_out, out, L, C, num_classes, apply_softmax);
  } else {
    if (two_stage) token_partial_sum_kernel<bf16><<<dim3(chunks, B), 256, 0, st>>>(static_cast<const bf16 *>(x), workspace, L, C, chunks);
    mean_head_kernel<bf16><<<B, 256, smem, st>>>(static_cast<const bf16 *>(x), two_stage ? workspace : nullptr, chunks, w, bias, feat_out, out, L, C, num_classes, apply_softmax);
  }
  MVIT_LAUNCH_OK("mean_head");
  return 0;
}

extern "C" int mvit_preprocess_u8_fwd(const uint8_t *frames, void *clip, int B, int T, int H, int W, float mean,
                                      float stdv, int dtype, void *stream) {
  using namespace mvit;
  MVIT_REQUIRE(frames && clip, "preprocess: null pointer");
  MVIT_REQUIRE(B >= 0 && T > 0 && H > 0 && W > 0 && stdv != 0.f, "preprocess: bad arguments");
  if (B == 0) return 0;
  const int64_t ppc = (int64_t)T * H * W, total = ppc * B;
  const int threads = 256;
  const unsigned blocks = (unsigned)std::min<int64_t>((total + threads - 1) / threads, (int64_t)num_sms() * 32);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == MVIT_F32)
    preprocess_u8_kernel<float><<<blocks, threads, 0, st>>>(frames, static_cast<float *>(clip), ppc, total, mean, stdv);
  else if (dtype == MVIT_BF16)
    preprocess_u8_kernel<bf16><<<blocks, threads, 0, st>>>(frames, static_cast<bf16 *>(clip), ppc, total, mean, stdv);
  else MVIT_REQUIRE(false, "preprocess: unknown dtype");
  MVIT_LAUNCH_OK("preprocess_u8");
  return 0;
}

extern "C" int mvit_im2col3d_fwd(const void *clip, void *patches, int B, int C, int T, int H, int W, int kt, int kh,
                                 int kw, int st, int sh, int sw, int pt, int ph, int pw, int Kp, int dtype,
                                 void *stream) {
  using namespace mvit;
  MVIT_REQUIRE(clip && patches, "im2col3d: null pointer");
  MVIT_REQUIRE(B >= 0 && C > 0 && T > 0 && H > 0 && W > 0 && kt > 0 && kh > 0 && kw > 0 && st > 0 && sh > 0 && sw > 0,
               "im2col3d: bad shape");
  MVIT_REQUIRE(Kp >= C * kt * kh * kw && Kp % 8 == 0, "im2col3d: Kp must be a multiple of 8 and >= C*kt*kh*kw");
  MVIT_REQUIRE((reinterpret_cast<uintptr_t>(patches) & 15) == 0, "im2col3d: output must be 16-byte aligned");
  MVIT_REQUIRE(kt < 256 && kh < 256 && kw < 256 && Kp <= 4096, "im2col3d: kernel too large");
  const int To = (T + 2 * pt - kt) / st + 1, Ho = (H + 2 * ph - kh) / sh + 1, Wo = (W + 2 * pw - kw) / sw + 1;
  MVIT_REQUIRE(To > 0 && Ho > 0 && Wo > 0, "im2col3d: empty output");
  if (B == 0) return 0;
  int vpt = 1;
  while (vpt < Kp / 8) vpt <<= 1;
  MVIT_REQUIRE(vpt <= 256, "im2col3d: Kp too large");
  const int64_t n_tok = (int64_t)B * To * Ho * Wo;
  const int tok_per_cta = 256 / vpt;
  const unsigned blocks = (unsigned)std::min<int64_t>((n_tok + tok_per_cta - 1) / tok_per_cta, (int64_t)num_sms() * 64);
  const size_t smem = (size_t)Kp * sizeof(int2);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dtype == MVIT_F32)
    im2col3d_kernel<float><<<blocks, 256, smem, s>>>(static_cast<const float *>(clip), static_cast<float *>(patches), B, C, T, H, W, kt, kh, kw, st, sh, sw, pt, ph, pw, To, Ho, Wo, Kp);
  else if (dtype == MVIT_BF16)
    im2col3d_kernel<bf16><<<blocks, 256, smem, s>>>(static_cast<const bf16 *>(clip), static_cast<bf16 *>(patches), B, C, T, H, W, kt, kh, kw, st, sh, sw, pt, ph, pw, To, Ho, Wo, Kp);
  else MVIT_REQUIRE(false, "im2col3d: unknown dtype");
  MVIT_LAUNCH_OK("im2col3d");
  return 0;
}

extern "C" int mvit_fold_clip_fwd(const void *clip, int src_kind, void *folded, int B, int C, int T, int H, int W,
                                  int st, int sh, int sw, int Cf, float mean, float stdv, void *stream) {
  using namespace mvit;
  MVIT_REQUIRE(clip && folded, "fold_clip: null pointer");
  MVIT_REQUIRE(B >= 0 && C > 0 && st > 0 && sh > 0 && sw > 0 && T % st == 0 && H % sh == 0 && W % sw == 0,
               "fold_clip: clip size must be a multiple of the stride");
  MVIT_REQUIRE(Cf >= st * sh * sw * C, "fold_clip: Cf too small");
  MVIT_REQUIRE(src_kind >= 0 && src_kind <= 2, "fold_clip: src_kind 0 = f32 [B,C,T,H,W], 1 = bf16 [B,C,T,H,W], 2 = u8 [B,T,H,W,C]");
  if (B == 0) return 0;
  const int Wf = W / sw;
  int strips = (Wf + kFoldTok - 1) / kFoldTok;
  int64_t blocks = (int64_t)B * (T / st) * (H / sh) * strips;
  MVIT_REQUIRE(blocks < ((int64_t)1 << 31), "fold_clip: grid too large");
  const size_t esz = src_kind == 0 ? 4 : (src_kind == 1 ? 2 : 1);
  const size_t smem = (size_t)C * st * sh * kFoldTok * sw * esz;
  MVIT_REQUIRE(smem <= 48 * 1024, "fold_clip: staging tile too large");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  bf16 *o = static_cast<bf16 *>(folded);
  // tuned path: the MViT patch-embed geometry with 16-byte aligned rows
  const bool fast = C == 3 && st == 2 && sh == 4 && sw == 4 && Cf == 128 && src_kind != 0 && Wf % 8 == 0 &&
                    (reinterpret_cast<uintptr_t>(clip) & 15) == 0 && (reinterpret_cast<uintptr_t>(folded) & 15) == 0 &&
                    (src_kind == 2 ? (W * 3) % 16 == 0 : W % 8 == 0);
  if (fast) {
    const int tok = Wf % 112 == 0 ? 112 : (Wf % 56 == 0 ? 56 : kFoldTok);      // a whole image row per CTA when it divides
    strips = (Wf + tok - 1) / tok;
    blocks = (int64_t)B * (T / st) * (H / sh) * strips;
    const bf16 *cb = static_cast<const bf16 *>(clip);
    const unsigned char *cu = static_cast<const unsigned char *>(clip);
    const unsigned nb = (unsigned)blocks;
    if (src_kind == 1) {
      if (tok == 112) fold_clip_244x3_kernel<bf16, false, 112><<<nb, 256, 0, s>>>(cb, o, B, T, H, W, mean, stdv, strips);
      else if (tok == 56) fold_clip_244x3_kernel<bf16, false, 56><<<nb, 256, 0, s>>>(cb, o, B, T, H, W, mean, stdv, strips);
      else fold_clip_244x3_kernel<bf16, false, kFoldTok><<<nb, 256, 0, s>>>(cb, o, B, T, H, W, mean, stdv, strips);
    } else {
      if (tok == 112) fold_clip_244x3_kernel<unsigned char, true, 112><<<nb, 256, 0, s>>>(cu, o, B, T, H, W, mean, stdv, strips);
      else if (tok == 56) fold_clip_244x3_kernel<unsigned char, true, 56><<<nb, 256, 0, s>>>(cu, o, B, T, H, W, mean, stdv, strips);
      else fold_clip_244x3_kernel<unsigned char, true, kFoldTok><<<nb, 256, 0, s>>>(cu, o, B, T, H, W, mean, stdv, strips);
    }
    MVIT_LAUNCH_OK("fold_clip(2,4,4)");
    return 0;
  }
  if (src_kind == 0)
    fold_clip_kernel<float, false><<<(unsigned)blocks, 256, smem, s>>>(static_cast<const float *>(clip), o, B, C, T, H, W, st, sh, sw, Cf, mean, stdv, strips);
  else if (src_kind == 1)
    fold_clip_kernel<bf16, false><<<(unsigned)blocks, 256, smem, s>>>(static_cast<const bf16 *>(clip), o, B, C, T, H, W, st, sh, sw, Cf, mean, stdv, strips);
  else
    fold_clip_kernel<unsigned char, true><<<(unsigned)blocks, 256, smem, s>>>(static_cast<const unsigned char *>(clip), o, B, C, T, H, W, st, sh, sw, Cf, mean, stdv, strips);
  MVIT_LAUNCH_OK("fold_clip");
  return 0;
}

static int patch_conv_entry(const void *folded, const void *wf, const float *bias, const void *pos, void *out, float *stats_out,
                            int B, int Tf, int Hf, int Wf, int Cf, int nt, int nh, int nw, int lo_t, int lo_h, int lo_w, int N,
                            void *stream) {
  using namespace mvit;
  MVIT_REQUIRE(folded && wf && out, "patch_conv: null pointer");
  MVIT_REQUIRE(B >= 0 && Tf > 0 && Hf > 0 && Wf > 0 && nt > 0 && nh > 0 && nw > 0 && N > 0, "patch_conv: bad shape");
  MVIT_REQUIRE((reinterpret_cast<uintptr_t>(stats_out) & 7) == 0, "patch_conv: stats_out must be 8-byte aligned");
  if (B == 0) return 0;
  return patch_conv_tc(folded, wf, bias, pos, out, stats_out, B, Tf, Hf, Wf, Cf, nt, nh, nw, lo_t, lo_h, lo_w, N,
                       static_cast<cudaStream_t>(stream));
}

extern "C" int mvit_patch_conv_fwd(const void *folded, const void *wf, const float *bias, const void *pos, void *out,
                                   int B, int Tf, int Hf, int Wf, int Cf, int nt, int nh, int nw, int lo_t, int lo_h,
                                   int lo_w, int N, void *stream) {
  return patch_conv_entry(folded, wf, bias, pos, out, nullptr, B, Tf, Hf, Wf, Cf, nt, nh, nw, lo_t, lo_h, lo_w, N, stream);
}

extern "C" int mvit_patch_conv_stats_fwd(const void *folded, const void *wf, const float *bias, const void *pos, void *out,
                                         float *stats_out, int B, int Tf, int Hf, int Wf, int Cf, int nt, int nh, int nw,
                                         int lo_t, int lo_h, int lo_w, int N, void *stream) {
  MVIT_REQUIRE(stats_out, "patch_conv_stats: stats_out is NULL");
  return patch_conv_entry(folded, wf, bias, pos, out, stats_out, B, Tf, Hf, Wf, Cf, nt, nh, nw, lo_t, lo_h, lo_w, N, stream);
}
