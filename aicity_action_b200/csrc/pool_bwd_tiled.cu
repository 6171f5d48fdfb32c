// Tuned weight gradient of the attention_pool depthwise Conv3d (3x3x3, stride (1,s,s), pad 1, head_dim 96):
//   dW[c][kt,kh,kw] += sum over (batch, head, output position) x[input position of that tap][c] * dy[output][c]
// (autograd of attention.py:172-212 pool_{q,k,v}).  Same data movement as the forward kernel (pool_tiled.cu): a CTA
// marches through the input frames of a 2 x 8 output tile with a cp.async ring of halo tiles and lane L owns channels
// {2L, 2L+1, 64+L}; here the 81 registers that hold the filter taps in the forward hold the 27 x 3 partial sums
// instead, and the three output frames an input frame touches supply dy through a rolling register window.  A CTA
// works through several tiles before it reduces its partial sums (shared memory, then one fp32 atomic per weight), so
// the 2592 weight addresses see a few hundred atomics per launch instead of one per output position.
#include "pool_tile_common.cuh"

namespace mvit {
namespace ptile {

constexpr int kWgCPW = 4;

template <typename T, int S>
__global__ void __launch_bounds__(kThreads, 2)
pool_wgrad_tiled_kernel(const T *__restrict__ in, const T *__restrict__ dy, float *__restrict__ dw, PoolParams p, int tiles_w,
                        int tiles, int tiles_per_cta, int t_per_cta) {
  constexpr int CPW = kWgCPW;
  using G = Geo<S, CPW>;
  constexpr int PITCH = IO<T>::kPitch;
  constexpr int CHUNKS = PITCH / 16;
  constexpr int kStages = Ring<T>::kStages;
  constexpr int FRAME = G::NPOS * PITCH;
  extern __shared__ __align__(16) uint8_t smem[];
  int *offs = reinterpret_cast<int *>(smem + kStages * FRAME);
  float *red = reinterpret_cast<float *>(offs + G::NPOS_PAD);   // [96 * 27] CTA-level reduction

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y / p.heads, head = blockIdx.y % p.heads;
  const int to0 = blockIdx.z * t_per_cta, to1 = min(to0 + t_per_cta, p.To);
  const int hl = warp >> 1, cl0 = (warp & 1) * 4;

  float2 axy[27];
  float az[27];
#pragma unroll
  for (int i = 0; i < 27; ++i) { axy[i] = make_float2(0.f, 0.f); az[i] = 0.f; }

  const T *src_bh = in + (int64_t)b * p.in_bs + (int64_t)head * p.in_hs;
  const int64_t frame_elems = (int64_t)p.H * p.W * p.in_ls;
  const T *dy_bh = dy + (int64_t)blockIdx.y * p.To * p.Ho * p.Wo * 96;

  const int tile_begin = blockIdx.x * tiles_per_cta, tile_end = min(tiles, tile_begin + tiles_per_cta);
  for (int tile = tile_begin; tile < tile_end; ++tile) {
    const int tile_h = tile / tiles_w, tile_w = tile % tiles_w;
    const int ho0 = tile_h * G::TH, wo0 = tile_w * TW;
    __syncthreads();                                   // previous tile's frames / offsets no longer in use
    for (int i = threadIdx.x; i < G::NPOS; i += kThreads) {
      const int r = i / G::NC, c = i % G::NC;
      const int hin = ho0 * S - 1 + (S < 3 ? r : (r / 3) * S + r % 3);
      const int win = wo0 * S - 1 + (S < 3 ? c : (c / 3) * S + c % 3);
      offs[i] = (hin >= 0 && hin < p.H && win >= 0 && win < p.W) ? (int)((hin * p.W + win) * p.in_ls) : -1;
    }
    __syncthreads();
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
    // dy of output frame `to` at this warp's row / 4 columns (zero outside the CTA's frame range or the image)
    const int ho = ho0 + hl;
    auto load_dy = [&](int to, float2 (&gxy)[CPW], float (&gz)[CPW]) {
      const bool ok = to >= to0 && to < to1 && ho < p.Ho;
#pragma unroll
      for (int j = 0; j < CPW; ++j) {
        const int wo = wo0 + cl0 + j;
        if (ok && wo < p.Wo) {
          IO<T>::load3(reinterpret_cast<const uint8_t *>(dy_bh + (int64_t)((to * p.Ho + ho) * p.Wo + wo) * 96), lane, gxy[j],
                       gz[j]);
        } else {
          gxy[j] = make_float2(0.f, 0.f);
          gz[j] = 0.f;
        }
      }
    };

    const int t_first = max(to0 - 1, 0), t_last = min(to1, p.T - 1);
#pragma unroll
    for (int i = 0; i < kStages - 1; ++i) {
      if (t_first + i <= t_last) load_frame(t_first + i, smem + i * FRAME);
      else asm volatile("cp.async.commit_group;" ::: "memory");
    }
    float2 gxy[3][CPW];
    float gz[3][CPW];
    load_dy(t_first - 1, gxy[0], gz[0]);
    load_dy(t_first, gxy[1], gz[1]);
    int slot = 0;
    for (int t = t_first; t <= t_last; ++t) {
      load_dy(t + 1, gxy[2], gz[2]);
      asm volatile("cp.async.wait_group %0;" ::"n"(kStages - 2) : "memory");
      __syncthreads();
      {
        const int tp = t + kStages - 1, sp = (slot + kStages - 1) % kStages;
        if (tp <= t_last) load_frame(tp, smem + sp * FRAME);
        else asm volatile("cp.async.commit_group;" ::: "memory");
      }
      const uint8_t *bufc = smem + slot * FRAME;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const uint8_t *rowp = bufc + ((hl * G::MS + kh) * G::NC + cl0 * G::MS) * PITCH;
        float2 xy[G::WC];
        float z[G::WC];
#pragma unroll
        for (int cc = 0; cc < G::WC; ++cc) IO<T>::load3(rowp + cc * PITCH, lane, xy[cc], z[cc]);
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
#pragma unroll
          for (int j = 0; j < CPW; ++j) {
            const int cin = j * G::MS + kw;
#pragma unroll
            for (int kt = 0; kt < 3; ++kt) {
              // input frame t is tap kt of output frame t + 1 - kt, whose dy sits in window slot 2 - kt
              const int tap = (kt * 3 + kh) * 3 + kw;
              axy[tap] = __ffma2_rn(xy[cin], gxy[2 - kt][j], axy[tap]);
              az[tap] = fmaf(z[cin], gz[2 - kt][j], az[tap]);
            }
          }
        }
      }
#pragma unroll
      for (int j = 0; j < CPW; ++j) {
        gxy[0][j] = gxy[1][j]; gxy[1][j] = gxy[2][j];
        gz[0][j] = gz[1][j]; gz[1][j] = gz[2][j];
      }
      slot = (slot + 1) % kStages;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  }

  // ---- CTA reduction: warps add their 2592 partial sums in turn, then one atomic per weight
  __syncthreads();
  for (int w = 0; w < kWarps; ++w) {
    if (warp == w) {
#pragma unroll
      for (int tap = 0; tap < 27; ++tap) {
        float *r0 = red + (2 * lane) * 27 + tap, *r1 = red + (2 * lane + 1) * 27 + tap, *r2 = red + (64 + lane) * 27 + tap;
        if (w == 0) { *r0 = axy[tap].x; *r1 = axy[tap].y; *r2 = az[tap]; }
        else { *r0 += axy[tap].x; *r1 += axy[tap].y; *r2 += az[tap]; }
      }
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < 96 * 27; i += kThreads) atomicAdd(&dw[i], red[i]);
}

template <typename T, int S>
static int launch_wgrad(const void *in, const void *dy, float *dw, const PoolParams &p, cudaStream_t st) {
  using G = Geo<S, kWgCPW>;
  const size_t smem = Ring<T>::kStages * (size_t)G::NPOS * IO<T>::kPitch + (size_t)G::NPOS_PAD * sizeof(int) +
                      96 * 27 * sizeof(float);
  MVIT_SMEM_OPT_IN((pool_wgrad_tiled_kernel<T, S>), smem);
  const int tiles_h = (p.Ho + G::TH - 1) / G::TH, tiles_w = (p.Wo + TW - 1) / TW, tiles = tiles_h * tiles_w;
  const int bh = p.B * p.heads;
  const int target = 4 * num_sms();                    // CTAs wanted (2 resident per SM, two waves)
  int t_per_cta = p.To;
  while (t_per_cta > 2 && (int64_t)tiles * bh * ((p.To + t_per_cta - 1) / t_per_cta) < target) t_per_cta = (t_per_cta + 1) / 2;
  const int tsplits = (p.To + t_per_cta - 1) / t_per_cta;
  int tiles_per_cta = (int)std::max<int64_t>(1, ((int64_t)tiles * bh * tsplits) / target);
  tiles_per_cta = std::min(tiles_per_cta, tiles);
  dim3 grid((tiles + tiles_per_cta - 1) / tiles_per_cta, bh, tsplits);
  pool_wgrad_tiled_kernel<T, S><<<grid, kThreads, smem, st>>>(static_cast<const T *>(in), static_cast<const T *>(dy), dw, p,
                                                              tiles_w, tiles, tiles_per_cta, t_per_cta);
  MVIT_LAUNCH_OK("attention_pool_bwd(wgrad, tiled)");
  return 0;
}

template <typename T>
static int dispatch_wgrad(const void *in, const void *dy, float *dw, const PoolParams &p, cudaStream_t st) {
  switch (p.sh) {
    case 1: return launch_wgrad<T, 1>(in, dy, dw, p, st);
    case 2: return launch_wgrad<T, 2>(in, dy, dw, p, st);
    case 4: return launch_wgrad<T, 4>(in, dy, dw, p, st);
    case 8: return launch_wgrad<T, 8>(in, dy, dw, p, st);
  }
  return 1;
}

}  // namespace ptile

// returns 1 if the tuned kernel does not apply, 0 on launch, <0 on error
int pool_wgrad_tiled_try(const void *in, const void *dy, float *dw, const PoolParams &p, int dtype, cudaStream_t st) {
  if (p.d != 96 || p.kt != 3 || p.kh != 3 || p.kw != 3 || p.st != 1 || p.sh != p.sw) return 1;
  if (p.sh != 1 && p.sh != 2 && p.sh != 4 && p.sh != 8) return 1;
  if ((int64_t)p.B * p.heads >= 65536) return 1;
  if ((int64_t)p.H * p.W * p.in_ls >= ((int64_t)1 << 31)) return 1;
  const int64_t es = dtype == MVIT_BF16 ? 2 : 4;
  auto al = [&](int64_t elems) { return (elems * es) % 16 == 0; };
  if ((reinterpret_cast<uintptr_t>(in) & 15) || !al(p.in_bs) || !al(p.in_ls) || !al(p.in_hs)) return 1;
  if (reinterpret_cast<uintptr_t>(dy) & 3) return 1;
  if (dtype == MVIT_BF16) return ptile::dispatch_wgrad<bf16>(in, dy, dw, p, st);
  return ptile::dispatch_wgrad<float>(in, dy, dw, p, st);
}

}  // namespace mvit
