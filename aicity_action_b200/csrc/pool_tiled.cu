// Tuned attention_pool: depthwise 3x3x3 Conv3d, stride (1,s,s), pad 1, head_dim 96, fused LayerNorm.
// (attention.py:172-212 pool_{q,k,v} + attention.py:66-67 norm_{q,k,v}; HBM-bound)
//
// A CTA (4 warps, two resident per SM) owns one (batch, head), a TH x 8 tile of output positions and a
// range of output frames, and marches through the input frames it needs ONCE: each frame's halo tile is
// brought into shared memory with 16-byte cp.async (zero-filled outside the image) through a 3-4 deep
// ring, so several frames per CTA are in flight and HBM latency is covered.
// A warp owns CPW consecutive output columns of one output row; lane L owns channels {2L, 2L+1, 64+L}
// (one 4-byte + one 2-byte shared load per input position, conflict-free).  The 27x3 filter taps of
// those channels live in registers.  An input frame t contributes to outputs t-1, t, t+1 through three
// rolling accumulator sets, so every input value is loaded from shared memory once per row it
// touches and reused across the kw taps of neighbouring columns.  FMAs on the channel pair use the
// packed fp32x2 FMA of sm_100 (bit-identical to two fmaf).  When a frame completes, the warp
// LayerNorms the 96 channels of each of its columns (fp32 statistics, shuffles) and stores 192
// contiguous bytes per token*head.
#include "pool_tile_common.cuh"

namespace mvit {
namespace ptile {

template <typename T, int S, int CPW>
__global__ void __launch_bounds__(kThreads, 2)
pool_tiled_kernel(const T *__restrict__ in, const float *__restrict__ weight, const float *__restrict__ gamma,
                  const float *__restrict__ beta, T *__restrict__ out, PoolParams p, int tiles_w, int t_per_cta) {
  using G = Geo<S, CPW>;
  constexpr int PITCH = IO<T>::kPitch;
  constexpr int CHUNKS = PITCH / 16;
  constexpr int kStages = Ring<T>::kStages;
  constexpr int FRAME = G::NPOS * PITCH;
  extern __shared__ __align__(16) uint8_t smem[];
  int *offs = reinterpret_cast<int *>(smem + kStages * FRAME);
  float *w_s = reinterpret_cast<float *>(offs + G::NPOS_PAD);   // [96][27] staged copy of the filter
  float *gb_s = w_s + 96 * 27;                               // gamma[96] | beta[96]
  float *stg = gb_s + 192 + (threadIdx.x >> 5) * (CPW * kStgPitch);   // this warp's [CPW][96 (+pad)] fp32 LN staging

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile_h = blockIdx.x / tiles_w, tile_w = blockIdx.x % tiles_w;
  const int b = blockIdx.y / p.heads, head = blockIdx.y % p.heads;
  const int to0 = blockIdx.z * t_per_cta, to1 = min(to0 + t_per_cta, p.To);   // output frames [to0, to1)
  const int ho0 = tile_h * G::TH, wo0 = tile_w * TW;

  // per-position element offset inside one input frame (or -1 outside the image)
  for (int i = threadIdx.x; i < G::NPOS; i += kThreads) {
    const int r = i / G::NC, c = i % G::NC;
    const int hin = ho0 * S - 1 + (S < 3 ? r : (r / 3) * S + r % 3);
    const int win = wo0 * S - 1 + (S < 3 ? c : (c / 3) * S + c % 3);
    offs[i] = (hin >= 0 && hin < p.H && win >= 0 && win < p.W) ? (int)((hin * p.W + win) * p.in_ls) : -1;
  }
  for (int i = threadIdx.x; i < 96 * 27; i += kThreads) w_s[i] = __ldg(weight + i);   // coalesced
  for (int i = threadIdx.x; i < 192; i += kThreads)
    gb_s[i] = p.has_ln ? (i < 96 ? __ldg(gamma + i) : __ldg(beta + i - 96)) : (i < 96 ? 1.f : 0.f);
  __syncthreads();
  // filter taps of this lane's channels: w[tap] = {c=2L, c=2L+1}, wz[tap] = c=64+L
  float2 wxy[27];
  float wz[27];
#pragma unroll
  for (int tap = 0; tap < 27; ++tap) {
    wxy[tap].x = w_s[(2 * lane) * 27 + tap];
    wxy[tap].y = w_s[(2 * lane + 1) * 27 + tap];
    wz[tap] = w_s[(64 + lane) * 27 + tap];
  }
  __syncthreads();

  const T *src_bh = in + (int64_t)b * p.in_bs + (int64_t)head * p.in_hs;
  const int64_t frame_elems = (int64_t)p.H * p.W * p.in_ls;
  auto load_frame = [&](int t, uint8_t *dst) {
    const T *base = src_bh + t * frame_elems;
    for (int i = threadIdx.x; i < G::NPOS * CHUNKS; i += kThreads) {
      const int pos = i / CHUNKS, ch = i - pos * CHUNKS;
      const int off = offs[pos];
      const T *src = off >= 0 ? base + off + ch * (16 / (int)sizeof(T)) : base;
      cp_async16(dst + pos * PITCH + ch * 16, src, off >= 0 ? 16 : 0);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  const int hl = CPW == 8 ? warp : (warp >> 1);          // output row of this warp inside the tile
  const int cl0 = CPW == 8 ? 0 : (warp & 1) * 4;         // first output column of this warp
  // rolling accumulators: a[0] -> output frame t-1, a[1] -> t, a[2] -> t+1 while input frame t is processed
  // axy: channel pair {2L, 2L+1} of column j; az2: channel 64+L of the column PAIR (2jp, 2jp+1) so that it, too,
  // is updated with packed FMAs
  float2 axy[3][CPW];
  float2 az2[3][CPW / 2];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
#pragma unroll
    for (int j = 0; j < CPW; ++j) axy[k][j] = make_float2(0.f, 0.f);
#pragma unroll
    for (int j = 0; j < CPW / 2; ++j) az2[k][j] = make_float2(0.f, 0.f);
  }

  const int t_first = max(to0 - 1, 0), t_last = min(to1, p.T - 1);   // input frames needed (st == 1)
  // ring prologue: frames t_first .. t_first+kStages-2 in flight (one commit group per frame, empty if past the end)
#pragma unroll
  for (int i = 0; i < kStages - 1; ++i) {
    if (t_first + i <= t_last) load_frame(t_first + i, smem + i * FRAME);
    else asm volatile("cp.async.commit_group;" ::: "memory");
  }
  int slot = 0;
  for (int t = t_first; t <= t_last; ++t) {
    // groups committed so far: (t - t_first) + kStages - 1; frame t is the oldest that may be pending
    asm volatile("cp.async.wait_group %0;" ::"n"(kStages - 2) : "memory");
    __syncthreads();   // frame t visible to all warps; everyone finished frame t-1 -> its slot is free
    {
      const int tp = t + kStages - 1, sp = (slot + kStages - 1) % kStages;
      if (tp <= t_last) load_frame(tp, smem + sp * FRAME);
      else asm volatile("cp.async.commit_group;" ::: "memory");
    }
    uint8_t *bufc = smem + slot * FRAME;
    // ---- accumulate this input frame into the three output frames it touches
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const uint8_t *rowp = bufc + ((hl * G::MS + kh) * G::NC + cl0 * G::MS) * PITCH;
      float2 xy[G::WC];
      float z[G::WC];
#pragma unroll
      for (int cc = 0; cc < G::WC; ++cc) IO<T>::load3(rowp + cc * PITCH, lane, xy[cc], z[cc]);
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        // output column j sees compact input column cin(j, kw) through tap kw
#pragma unroll
        for (int j = 0; j < CPW; ++j) {
          const int cin = j * G::MS + kw;
#pragma unroll
          for (int kt = 0; kt < 3; ++kt) {
            // input frame t is tap kt of output frame t + 1 - kt  -> accumulator slot 2 - kt
            axy[2 - kt][j] = __ffma2_rn(xy[cin], wxy[(kt * 3 + kh) * 3 + kw], axy[2 - kt][j]);
          }
        }
#pragma unroll
        for (int jp = 0; jp < CPW / 2; ++jp) {
          const float2 zz = make_float2(z[(2 * jp) * G::MS + kw], z[(2 * jp + 1) * G::MS + kw]);
#pragma unroll
          for (int kt = 0; kt < 3; ++kt) {
            const float w1 = wz[(kt * 3 + kh) * 3 + kw];
            az2[2 - kt][jp] = __ffma2_rn(zz, make_float2(w1, w1), az2[2 - kt][jp]);
          }
        }
      }
    }
    // ---- output frame t-1 is complete (and frame t too when t is the last input frame).
    // LayerNorm + store through a per-warp fp32 staging tile: the conv layout (lane = 3 channels of every column)
    // is turned into "LPC lanes per column, 96/LPC contiguous channels each", so the two reductions cost
    // log2(LPC) shuffles for ALL columns at once and every lane stores one contiguous, vectorised piece of a row.
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
      const int to = pass == 0 ? t - 1 : t;
      const bool emit = to >= to0 && to < to1 && (pass == 0 || t == p.T - 1);
      if (emit) {
#pragma unroll
        for (int j = 0; j < CPW; ++j) {
          *reinterpret_cast<float2 *>(stg + j * kStgPitch + 2 * lane) = pass == 0 ? axy[0][j] : axy[1][j];
          const float2 zz = pass == 0 ? az2[0][j / 2] : az2[1][j / 2];
          stg[j * kStgPitch + 64 + lane] = (j & 1) ? zz.y : zz.x;
        }
        __syncwarp();
        constexpr int LPC = 32 / CPW;            // lanes per column: 4 or 8
        constexpr int CH = 96 / LPC;             // channels per lane: 24 or 12
        const int jc = lane / LPC, part = lane % LPC;
        float x[CH];
#pragma unroll
        for (int k = 0; k < CH / 4; ++k) {
          const float4 v = *reinterpret_cast<const float4 *>(stg + jc * kStgPitch + part * CH + 4 * k);
          x[4 * k] = v.x; x[4 * k + 1] = v.y; x[4 * k + 2] = v.z; x[4 * k + 3] = v.w;
        }
        const int ho = ho0 + hl, wo = wo0 + cl0 + jc;
        const bool inside = ho < p.Ho && wo < p.Wo;
        if (p.pre_out != nullptr && inside) {      // training: keep the conv output for the LayerNorm backward
          T *prow = static_cast<T *>(p.pre_out) +
                    (((int64_t)blockIdx.y * p.To + to) * p.Ho * p.Wo + (int64_t)ho * p.Wo + wo) * 96 + part * CH;
          if constexpr (sizeof(T) == 2) {
#pragma unroll
            for (int k = 0; k < CH / 4; ++k) {
              __nv_bfloat162 h0 = __floats2bfloat162_rn(x[4 * k], x[4 * k + 1]), h1 = __floats2bfloat162_rn(x[4 * k + 2], x[4 * k + 3]);
              *reinterpret_cast<uint2 *>(prow + 4 * k) = make_uint2(*reinterpret_cast<uint32_t *>(&h0), *reinterpret_cast<uint32_t *>(&h1));
            }
          } else {
#pragma unroll
            for (int k = 0; k < CH / 4; ++k)
              *reinterpret_cast<float4 *>(prow + 4 * k) = make_float4(x[4 * k], x[4 * k + 1], x[4 * k + 2], x[4 * k + 3]);
          }
        }
        if (p.has_ln) {
          float sum = 0.f;
#pragma unroll
          for (int k = 0; k < CH; ++k) sum += x[k];
#pragma unroll
          for (int o = LPC / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
          const float mean = sum * (1.0f / 96.0f);
          float ss = 0.f;
#pragma unroll
          for (int k = 0; k < CH; ++k) { x[k] -= mean; ss = fmaf(x[k], x[k], ss); }
#pragma unroll
          for (int o = LPC / 2; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
          const float rstd = rsqrtf(ss * (1.0f / 96.0f) + p.eps);
#pragma unroll
          for (int k = 0; k < CH / 4; ++k) {
            const float4 g = *reinterpret_cast<const float4 *>(gb_s + part * CH + 4 * k);
            const float4 bt = *reinterpret_cast<const float4 *>(gb_s + 96 + part * CH + 4 * k);
            x[4 * k] = fmaf(x[4 * k] * rstd, g.x, bt.x);
            x[4 * k + 1] = fmaf(x[4 * k + 1] * rstd, g.y, bt.y);
            x[4 * k + 2] = fmaf(x[4 * k + 2] * rstd, g.z, bt.z);
            x[4 * k + 3] = fmaf(x[4 * k + 3] * rstd, g.w, bt.w);
          }
        }
        if (inside) {
          T *row = out + (int64_t)b * p.out_bs + (int64_t)((to * p.Ho + ho) * p.Wo + wo) * p.out_ls +
                   (int64_t)head * p.out_hs + part * CH;
          if constexpr (sizeof(T) == 2) {
            uint32_t w[CH / 2];
#pragma unroll
            for (int k = 0; k < CH / 2; ++k) {
              __nv_bfloat162 h = __floats2bfloat162_rn(x[2 * k], x[2 * k + 1]);
              w[k] = *reinterpret_cast<uint32_t *>(&h);
            }
            if constexpr (CH == 24) {
#pragma unroll
              for (int k = 0; k < 3; ++k)
                *reinterpret_cast<uint4 *>(row + 8 * k) = make_uint4(w[4 * k], w[4 * k + 1], w[4 * k + 2], w[4 * k + 3]);
            } else {
#pragma unroll
              for (int k = 0; k < 3; ++k) *reinterpret_cast<uint2 *>(row + 4 * k) = make_uint2(w[2 * k], w[2 * k + 1]);
            }
          } else {
#pragma unroll
            for (int k = 0; k < CH / 4; ++k)
              *reinterpret_cast<float4 *>(row + 4 * k) = make_float4(x[4 * k], x[4 * k + 1], x[4 * k + 2], x[4 * k + 3]);
          }
        }
        __syncwarp();   // staging is rewritten by the next emit
      }
    }
    // rotate: slot0 <- slot1 <- slot2 <- 0
#pragma unroll
    for (int j = 0; j < CPW; ++j) {
      axy[0][j] = axy[1][j];
      axy[1][j] = axy[2][j];
      axy[2][j] = make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int j = 0; j < CPW / 2; ++j) {
      az2[0][j] = az2[1][j];
      az2[1][j] = az2[2][j];
      az2[2][j] = make_float2(0.f, 0.f);
    }
    slot = (slot + 1) % kStages;
  }
}

template <typename T, int S, int CPW>
static int launch(const void *in, const float *w, const float *g, const float *b, void *out, const PoolParams &p,
                  cudaStream_t st) {
  using G = Geo<S, CPW>;
  const size_t smem = Ring<T>::kStages * (size_t)G::NPOS * IO<T>::kPitch + (size_t)G::NPOS_PAD * sizeof(int) +
                      (96 * 27 + 192 + kWarps * CPW * kStgPitch) * sizeof(float);
  MVIT_SMEM_OPT_IN((pool_tiled_kernel<T, S, CPW>), smem);
  const int tiles_h = (p.Ho + G::TH - 1) / G::TH, tiles_w = (p.Wo + TW - 1) / TW;
  const int bh = p.B * p.heads;
  // split the frame axis until the grid covers the GPU about twice (each split re-reads one halo frame)
  int t_per_cta = p.To;
  while (t_per_cta > 2 && (int64_t)tiles_h * tiles_w * bh * ((p.To + t_per_cta - 1) / t_per_cta) < 2 * num_sms())
    t_per_cta = (t_per_cta + 1) / 2;
  dim3 grid(tiles_h * tiles_w, bh, (p.To + t_per_cta - 1) / t_per_cta);
  pool_tiled_kernel<T, S, CPW><<<grid, kThreads, smem, st>>>(static_cast<const T *>(in), w, g, b, static_cast<T *>(out),
                                                             p, tiles_w, t_per_cta);
  MVIT_LAUNCH_OK("attention_pool(tiled)");
  return 0;
}

template <typename T>
static int dispatch(const void *in, const float *w, const float *g, const float *b, void *out, const PoolParams &p,
                    cudaStream_t st) {
  switch (p.sh) {
    case 1: return launch<T, 1, 8>(in, w, g, b, out, p, st);
    case 2: return launch<T, 2, 4>(in, w, g, b, out, p, st);
    case 4: return launch<T, 4, 4>(in, w, g, b, out, p, st);
    case 8: return launch<T, 8, 4>(in, w, g, b, out, p, st);
  }
  return 1;
}

}  // namespace ptile

int pool_tiled_try(const void *in, const float *w, const float *g, const float *b, void *out, const PoolParams &p,
                   int mode, int dtype, cudaStream_t st) {
  if (mode != MVIT_POOL_CONV || p.d != 96 || p.has_cls) return 1;
  if (p.kt != 3 || p.kh != 3 || p.kw != 3 || p.st != 1 || p.sh != p.sw) return 1;
  if (p.sh != 1 && p.sh != 2 && p.sh != 4 && p.sh != 8) return 1;
  if ((int64_t)p.B * p.heads >= 65536) return 1;
  if ((int64_t)p.H * p.W * p.in_ls >= ((int64_t)1 << 31)) return 1;
  // 16-byte cp.async needs aligned rows: every stride and the base must be multiples of 16 bytes
  const int64_t es = dtype == MVIT_BF16 ? 2 : 4;
  auto al = [&](int64_t elems) { return (elems * es) % 16 == 0; };
  if ((reinterpret_cast<uintptr_t>(in) & 15) || !al(p.in_bs) || !al(p.in_ls) || !al(p.in_hs)) return 1;
  if ((reinterpret_cast<uintptr_t>(out) & 15) || (p.out_bs * es) % 16 || (p.out_ls * es) % 16 || (p.out_hs * es) % 16) return 1;
  if (dtype == MVIT_BF16) return ptile::dispatch<bf16>(in, w, g, b, out, p, st);
  return ptile::dispatch<float>(in, w, g, b, out, p, st);
}

}  // namespace mvit
