// attention_pool for the shipped MViTv2 configuration, bf16: the three depthwise 3x3x3 Conv3d poolings of Q, K and V
// (stride (1,s,s), pad 1, head_dim 96) fused with their LayerNorms, reading the qkv GEMM output [B, N, 3, heads, 96]
// in place (attention.py:172-212 pool_{q,k,v} + attention.py:66-67 norm_{q,k,v}).  HBM-bound overall;
// the stride-1 pools are co-limited by the fp32 FMA pipe (27 FMA per output element), so the design goal is that the
// warps which own the FMA pipe issue (almost) nothing but FMAs.
//
//  * PERSISTENT CTAs (two per SM) walk a static round-robin list of work items = (q|k|v, batch*head, 4x8 or 2x8 output
//    tile, frame range); q, k and v of equal stride share ONE launch (blockIdx.y picks the tensor, so the 81 filter taps a
//    lane owns stay in registers for the CTA's whole life).
//  * A PRODUCER WARP feeds a ring of halo tiles with one 5-D TMA load per input frame (tensor map over the qkv tensor:
//    [channel, w, h, t, b]; the out-of-image halo is zero-filled by the TMA unit = the conv's zero padding), completion on
//    full/empty mbarriers — no per-thread cp.async issue, no offset table, no __syncthreads in the frame loop.
//  * FOUR CONVOLUTION WARPS (warpgroup 0): a warp owns CPW consecutive output columns of one output row.  Every lane owns
//    the channel pair {2L, 2L+1} of all CPW columns, and the channel pair {64+2(L%16), 65+2(L%16)} of the half of the
//    columns its half-warp is responsible for — so ALL multiply-adds are packed fp32x2 FMAs on natural channel pairs
//    (bit-identical to two fmaf), 324 per warp and frame, with no register shuffling to form pairs.  An input frame t feeds
//    output frames t-1, t, t+1 through three rolling accumulator sets; the frame loop is unrolled by three so the rotation
//    is a renaming, not register moves.  A finished output frame leaves the warp as fp32 through a double-buffered staging
//    tile in shared memory (12 stores per lane + one mbarrier arrive; three tiles per warp) — the convolution warps never run the LayerNorm.
//  * THREE LAYERNORM WARPS (the producer's warpgroup) drain the staging tiles round-robin: 4 (8) lanes per column, 24 (12)
//    channels per lane, one sweep for sum and sum of squares, packed arithmetic, bf16 rows stored straight to
//    [B, heads, L', 96].  ncu on the previous
//    version (LayerNorm inside the convolution warps) showed them spending as many stall samples in the latency-bound
//    LayerNorm / store tail as in the convolution, with the FMA pipe 44 % busy; now the tail overlaps the FMAs of all
//    four convolution warps.
//  * REGISTERS: the CTA is two warpgroups launched at 128 registers per thread (two CTAs per SM); the producer / LayerNorm
//    warpgroup gives registers back (setmaxnreg.dec 80) and the convolution warpgroup takes them (setmaxnreg.inc 176): the
//    54 filter taps + 72 accumulators + a row of inputs live in registers without spills at full occupancy.
#include <algorithm>
#include <type_traits>

#include "pool.cuh"
#include "tc_common.cuh"

namespace mvit {
namespace ptma {

using namespace tc;

constexpr int kConvWarps = 4;
constexpr int kLnWarps = 3;        // warps 5-7; warp 4 is the TMA producer
constexpr int kStgBufs = 3;        // staging tiles per convolution warp: emit n uses tile n % 3, which makes LayerNorm warp
                                   // (n + w) % 3 the ONLY consumer of tile (w, n % 3) — a waiter that could run two phases
                                   // ahead of an mbarrier would see its parity test pass falsely
constexpr int kThreads = 256;
constexpr int TW = 8;
constexpr int kStgPitch = 100;   // floats per staged column (96 + 4 pad)
constexpr int kPitch = 192;      // bytes per position (96 bf16)

template <int S> struct Geo {
  static constexpr int CPW = S == 1 ? 8 : 4;                            // output columns per warp
  static constexpr int TH = S == 1 ? kConvWarps : kConvWarps / 2;       // one warp per row, or two warps per row
  static constexpr int NR = (TH - 1) * S + 3;                           // input rows / cols of the halo tile
  static constexpr int NC = (TW - 1) * S + 3;
  static constexpr int WC = (CPW - 1) * S + 3;                          // input cols one warp touches
  static constexpr int kFrameBytes = NR * NC * kPitch;
  static constexpr int kStageBytes = (kFrameBytes + 127) & ~127;        // TMA destinations stay 128-byte aligned
  static constexpr int kStages = S == 1 ? 6 : 5;
  static constexpr int kStgFloats = CPW * kStgPitch;                    // one staging tile
};

struct Stream {              // one of q / k / v
  const float *w, *gamma, *beta;
  bf16 *out, *pre;           // [B, heads, L', 96] contiguous; pre = conv output before the LayerNorm (training) or NULL
  int c0;                    // first channel of this tensor inside a token row of qkv: which * heads * 96
};
struct Params {
  Stream s[3];
  int heads, T, To, Ho, Wo, tiles_h, tiles_w, t_per_item, t_splits, items;
  float eps;
};
struct StgMeta {             // where a staged tile goes: written by the convolution warp, read by the LayerNorm warp
  long long off;             // element offset of (column 0, channel 0) in out / pre
  int ncols, pad;            // live columns (0: nothing to store — halo frame of a split item, or a row below the image)
};


// mbarrier operations on a raw shared-window address: the convolution warps keep their barrier addresses in registers
// (re-deriving them from generic pointers cost ~40 dependent instructions per frame, S2R / S2UR reads included)
__device__ __forceinline__ void bar_wait(uint32_t addr, uint32_t parity) {
  if (mbar_try_wait(addr, parity)) return;
  uint32_t spins = 0;
  while (!mbar_try_wait(addr, parity)) {
    if ((++spins & 1023u) == 0) {
      if (*reinterpret_cast<volatile unsigned int *>(&g_tc_fault) != 0) return;
      if (spins > (1u << 26)) {
        atomicExch(&g_tc_fault, 1u);
        return;
      }
    }
  }
}
__device__ __forceinline__ void bar_arrive(uint32_t addr) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory");
}

template <int S>
__global__ void __launch_bounds__(kThreads, 2)
pool_tma_kernel(const __grid_constant__ CUtensorMap tmap, const Params p) {
  using G = Geo<S>;
  constexpr int CPW = G::CPW;
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t *stages = smem;
  uint64_t *full = reinterpret_cast<uint64_t *>(smem + G::kStages * G::kStageBytes);
  uint64_t *empty = full + G::kStages;
  uint64_t *stg_full = empty + G::kStages;                              // [conv warp][kStgBufs]
  uint64_t *stg_empty = stg_full + kStgBufs * kConvWarps;
  StgMeta *meta = reinterpret_cast<StgMeta *>(stg_empty + kStgBufs * kConvWarps);   // [conv warp][kStgBufs]
  float *gb_s = reinterpret_cast<float *>(meta + kStgBufs * kConvWarps);   // gamma[96] | beta[96]
  float2 *wz_s = reinterpret_cast<float2 *>(gb_s + 192);                // [27 taps][16 channel pairs of 64..95]
  float *stg_all = reinterpret_cast<float *>(wz_s + 27 * 16);           // [conv warp][kStgBufs][CPW][kStgPitch]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const Stream &sm = p.s[blockIdx.y];
  if (threadIdx.x == 0) {
    for (int i = 0; i < G::kStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], kConvWarps);
    }
    for (int i = 0; i < kStgBufs * kConvWarps; ++i) {
      mbar_init(&stg_full[i], 1);
      mbar_init(&stg_empty[i], 1);
    }
    fence_barrier_init();
  }
  pdl_wait();                 // the qkv GEMM has completed (see common.cuh)
  pdl_launch_dependents();
  for (int i = threadIdx.x; i < 192; i += kThreads)
    gb_s[i] = sm.gamma ? (i < 96 ? __ldg(sm.gamma + i) : __ldg(sm.beta + i - 96)) : (i < 96 ? 1.f : 0.f);
  for (int i = threadIdx.x; i < 27 * 16; i += kThreads) {
    const int tap = i >> 4, zl = i & 15;
    wz_s[i] = make_float2(__ldg(sm.w + (64 + 2 * zl) * 27 + tap), __ldg(sm.w + (65 + 2 * zl) * 27 + tap));
  }
  __syncthreads();

  const int tiles = p.tiles_h * p.tiles_w;
  // item -> (bh, tile, frame range); consecutive items (= CTAs working at the same time) are neighbouring tiles of one
  // (batch, head): their halos hit L2
  auto decode = [&](int item, int &bh, int &tile_h, int &tile_w, int &to0, int &to1) {
    const int ts = item % p.t_splits;
    const int r = item / p.t_splits;
    const int tile = r % tiles;
    bh = r / tiles;
    tile_h = tile / p.tiles_w;
    tile_w = tile - tile_h * p.tiles_w;
    to0 = ts * p.t_per_item;
    to1 = min(to0 + p.t_per_item, p.To);
  };
  // frames an item reads / steps it takes: one step per input frame, plus a flush step after the last frame of the clip
  auto frame_range = [&](int to0, int to1, int &t_first, int &t_last, int &t_end) {
    t_first = max(to0 - 1, 0);
    t_last = min(to1, p.T - 1);
    t_end = t_last + (to1 == p.T ? 1 : 0);
  };

  if (warp >= kConvWarps) {
    // ------------------------------------------------------------------ producer / LayerNorm warpgroup
    asm volatile("setmaxnreg.dec.sync.aligned.u32 80;");
    if (warp == kConvWarps) {
      if (lane == 0) {                                                  // one lane issues every TMA load
        tma_prefetch_desc(&tmap);
        uint32_t it = 0;
        for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
          int bh, tile_h, tile_w, to0, to1, t_first, t_last, t_end;
          decode(item, bh, tile_h, tile_w, to0, to1);
          frame_range(to0, to1, t_first, t_last, t_end);
          const int b = bh / p.heads, head = bh - b * p.heads;
          const int h0 = tile_h * G::TH * S - 1, w0 = tile_w * TW * S - 1;
          for (int t = t_first; t <= t_last; ++t, ++it) {
            const uint32_t slot = it % G::kStages, ph = (it / G::kStages) & 1;
            mbar_wait(&empty[slot], ph ^ 1);
            mbar_arrive_expect_tx(&full[slot], G::kFrameBytes);
            tma_load_5d(stages + slot * G::kStageBytes, &tmap, &full[slot], sm.c0 + head * 96, w0, h0, t, b);
          }
        }
      }
      return;
    }
    // ---- LayerNorm warps: staged tiles are numbered idx = 4 * n + w (n-th emit of convolution warp w); warp l takes
    // idx = l, l + 3, ... in increasing order (every wait is for a tile older than anything that could wait on us)
    const int l = warp - kConvWarps - 1;
    constexpr int LPC = 32 / CPW;                                       // lanes per column: 4 (stride 1) or 8
    constexpr int CH = 96 / LPC;                                        // channels per lane: 24 or 12
    const int jc = lane / LPC, part = lane % LPC;                       // column of the tile, channel slice
    uint32_t n = 0;                                                     // emits per convolution warp so far
    uint32_t kb = 0, kph = 0;                                           // n % 3 and (n / 3) & 1
    int next = l;                                                       // next idx of this warp, relative to 4 * n
    for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
      int bh, tile_h, tile_w, to0, to1, t_first, t_last, t_end;
      decode(item, bh, tile_h, tile_w, to0, to1);
      frame_range(to0, to1, t_first, t_last, t_end);
      for (int t = t_first; t <= t_end; ++t, ++n) {
        for (; next < kConvWarps; next += kLnWarps) {
          const int w = next;
          const uint32_t bsel = kb;
          mbar_wait(&stg_full[w * kStgBufs + bsel], kph);
          const StgMeta m = meta[w * kStgBufs + bsel];
          if (m.ncols > 0) {                                            // warp-uniform
            const float *src = stg_all + (w * kStgBufs + bsel) * G::kStgFloats + jc * kStgPitch + part * CH;
            float2 x[CH / 2];
#pragma unroll
            for (int k = 0; k < CH / 4; ++k) {
              const float4 v = *reinterpret_cast<const float4 *>(src + 4 * k);
              x[2 * k] = make_float2(v.x, v.y);
              x[2 * k + 1] = make_float2(v.z, v.w);
            }
            const bool live = jc < m.ncols;
            auto store_row = [&](bf16 *base) {
              uint32_t wd[CH / 2];
#pragma unroll
              for (int k = 0; k < CH / 2; ++k) {
                const __nv_bfloat162 h = __floats2bfloat162_rn(x[k].x, x[k].y);
                wd[k] = *reinterpret_cast<const uint32_t *>(&h);
              }
              bf16 *row = base + m.off + jc * 96 + part * CH;
              if constexpr (CH == 24) {
#pragma unroll
                for (int k = 0; k < 3; ++k)
                  *reinterpret_cast<uint4 *>(row + 8 * k) = make_uint4(wd[4 * k], wd[4 * k + 1], wd[4 * k + 2], wd[4 * k + 3]);
              } else {
#pragma unroll
                for (int k = 0; k < 3; ++k) *reinterpret_cast<uint2 *>(row + 4 * k) = make_uint2(wd[2 * k], wd[2 * k + 1]);
              }
            };
            if (sm.pre != nullptr && live) store_row(sm.pre);           // training: keep the conv output for the LN backward
            if (sm.gamma != nullptr) {
              // one sweep: sum and sum of squares on independent accumulators, reduced together (fp32; the variance of 96
              // O(1) values loses nothing to the E[x^2] - mean^2 form), then y = x * (rstd*gamma) + (beta - mean*rstd*gamma)
              float2 sa = x[0], sb = x[1], qa = __fmul2_rn(x[0], x[0]), qb = __fmul2_rn(x[1], x[1]);
#pragma unroll
              for (int k = 2; k < CH / 2; k += 2) {
                sa = __fadd2_rn(sa, x[k]);
                sb = __fadd2_rn(sb, x[k + 1]);
                qa = __ffma2_rn(x[k], x[k], qa);
                qb = __ffma2_rn(x[k + 1], x[k + 1], qb);
              }
              const float2 s2 = __fadd2_rn(sa, sb), q2 = __fadd2_rn(qa, qb);
              float sum = s2.x + s2.y, ss = q2.x + q2.y;
#pragma unroll
              for (int o = LPC / 2; o > 0; o >>= 1) {
                sum += __shfl_xor_sync(0xffffffffu, sum, o);
                ss += __shfl_xor_sync(0xffffffffu, ss, o);
              }
              const float mean = sum * (1.0f / 96.0f);
              const float rstd = rsqrtf(fmaxf(ss * (1.0f / 96.0f) - mean * mean, 0.f) + p.eps);
              const float2 r2 = make_float2(rstd, rstd), nm = make_float2(-mean * rstd, -mean * rstd);
#pragma unroll
              for (int k = 0; k < CH / 4; ++k) {
                const float4 g = *reinterpret_cast<const float4 *>(gb_s + part * CH + 4 * k);
                const float4 bt = *reinterpret_cast<const float4 *>(gb_s + 96 + part * CH + 4 * k);
                const float2 g0 = make_float2(g.x, g.y), g1 = make_float2(g.z, g.w);
                x[2 * k] = __ffma2_rn(x[2 * k], __fmul2_rn(r2, g0), __ffma2_rn(nm, g0, make_float2(bt.x, bt.y)));
                x[2 * k + 1] = __ffma2_rn(x[2 * k + 1], __fmul2_rn(r2, g1), __ffma2_rn(nm, g1, make_float2(bt.z, bt.w)));
              }
            }
            if (live) store_row(sm.out);
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&stg_empty[w * kStgBufs + bsel]);  // the convolution warp may refill the tile
        }
        next -= kConvWarps;
        if (++kb == kStgBufs) { kb = 0; kph ^= 1; }
      }
    }
    return;
  }

  // -------------------------------------------------------------------- convolution warpgroup
  asm volatile("setmaxnreg.inc.sync.aligned.u32 176;");
  // filter taps of the channel pair {2L, 2L+1} in registers; those of the second pair come from shared memory (wz_s)
  float2 wxy[27];
#pragma unroll
  for (int tap = 0; tap < 27; ++tap) {
    wxy[tap].x = __ldg(sm.w + (2 * lane) * 27 + tap);
    wxy[tap].y = __ldg(sm.w + (2 * lane + 1) * 27 + tap);
  }
  const int hl = CPW == 8 ? warp : (warp >> 1);          // output row of this warp inside the tile
  const int cl0 = CPW == 8 ? 0 : (warp & 1) * 4;         // first output column of this warp
  constexpr int ZC = CPW / 2;                            // columns whose channels 64..95 this half-warp accumulates
  const int zl = lane & 15, jz0 = (lane >> 4) * ZC;
  // rolling accumulators; axy: channel pair {2L, 2L+1} of column j; az2: channel pair {64+2zl, 65+2zl} of column jz0 + j
  float2 axy[3][CPW];
  float2 az2[3][ZC];

  int to0 = 0, to1 = 0;
  // control state kept incremental (no divisions, no address re-derivation in the frame loop)
  const uint32_t full0 = smem_u32(full), empty0 = smem_u32(empty);
  const uint32_t stgf0 = smem_u32(stg_full) + warp * kStgBufs * 8, stge0 = smem_u32(stg_empty) + warp * kStgBufs * 8;
  uint32_t slot = 0, slot_ph = 0;                      // TMA ring position
  uint32_t e_buf = 0, e_ph = 0;                        // staging tile of the next emit: count % 3, (count / 3) & 1
  long long out_off = 0;                               // element offset of this warp's first column in the frame being emitted
  int row_cols = 0;                                    // live columns of this warp's row (0 below the image)
  const long long frame_elems = (long long)p.Ho * p.Wo * 96;

  // accumulate input frame `buf` into the three output frames it touches: tap kt of output frame t + 1 - kt lives in
  // accumulator set (R + 2 - kt) % 3
  auto accumulate = [&](const uint8_t *bufc, auto Rc) {
    constexpr int R = decltype(Rc)::value;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const uint8_t *rowp = bufc + ((hl * S + kh) * G::NC + cl0 * S) * kPitch;
      constexpr int ZW = (ZC - 1) * S + 3;               // input columns the half-warp's second channel pair touches
      float2 xy[G::WC], z2[ZW];
#pragma unroll
      for (int cc = 0; cc < G::WC; ++cc) {
        const uint32_t u = *reinterpret_cast<const uint32_t *>(rowp + cc * kPitch + 4 * lane);
        xy[cc] = make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
      }
#pragma unroll
      for (int cc = 0; cc < ZW; ++cc) {
        const uint32_t u = *reinterpret_cast<const uint32_t *>(rowp + (jz0 * S + cc) * kPitch + 128 + 4 * zl);
        z2[cc] = make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
      }
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
#pragma unroll
        for (int kt = 0; kt < 3; ++kt) {
          const int tap = (kt * 3 + kh) * 3 + kw;
          const float2 wz = wz_s[tap * 16 + zl];
#pragma unroll
          for (int j = 0; j < CPW; ++j)
            axy[(R + 2 - kt) % 3][j] = __ffma2_rn(xy[j * S + kw], wxy[tap], axy[(R + 2 - kt) % 3][j]);
#pragma unroll
          for (int j = 0; j < ZC; ++j)
            az2[(R + 2 - kt) % 3][j] = __ffma2_rn(z2[j * S + kw], wz, az2[(R + 2 - kt) % 3][j]);
        }
      }
    }
  };

  // hand output frame `to` (accumulator set A) to the LayerNorm warps, then clear the set.  Straight-line on purpose:
  // frames outside [to0, to1) (the halo frame of a split item) are staged too, with ncols = 0 — a branch here costs
  // register shuffles at the merge in the unrolled loop, and the emit count stays in step with the LayerNorm warps'.
  auto emit = [&](int to, auto Ac) {
    constexpr int A = decltype(Ac)::value;
    bar_wait(stge0 + e_buf * 8, e_ph ^ 1);                              // the tile's previous contents (3 emits ago) are drained
    float *stg = stg_all + (warp * kStgBufs + e_buf) * G::kStgFloats;
#pragma unroll
    for (int j = 0; j < CPW; ++j) *reinterpret_cast<float2 *>(stg + j * kStgPitch + 2 * lane) = axy[A][j];
#pragma unroll
    for (int j = 0; j < ZC; ++j) *reinterpret_cast<float2 *>(stg + (jz0 + j) * kStgPitch + 64 + 2 * zl) = az2[A][j];
    if (lane == 0) {
      StgMeta m;
      m.off = out_off;
      m.ncols = (to >= to0 && to < to1) ? row_cols : 0;
      m.pad = 0;
      meta[warp * kStgBufs + e_buf] = m;
    }
    out_off += frame_elems;
    __syncwarp();
    if (lane == 0) bar_arrive(stgf0 + e_buf * 8);
    if (++e_buf == kStgBufs) { e_buf = 0; e_ph ^= 1; }
#pragma unroll
    for (int j = 0; j < CPW; ++j) axy[A][j] = make_float2(0.f, 0.f);
#pragma unroll
    for (int j = 0; j < ZC; ++j) az2[A][j] = make_float2(0.f, 0.f);
  };

  // one step: input frame t (if there is one) is accumulated, then output frame t-1 is complete.  R = (t - t_first) % 3
  // names the accumulator set that holds output frame t-1.
  auto step = [&](int t, int t_last, auto Rc) {
    constexpr int R = decltype(Rc)::value;
    if (t <= t_last) {
      bar_wait(full0 + slot * 8, slot_ph);
      accumulate(stages + slot * G::kStageBytes, Rc);
      __syncwarp();
      if (lane == 0) bar_arrive(empty0 + slot * 8);      // this warp is done reading the halo tile
      if (++slot == G::kStages) { slot = 0; slot_ph ^= 1; }
    }
    emit(t - 1, std::integral_constant<int, R>{});
  };

  for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
    int bh, tile_h, tile_w, t_first, t_last, t_end;
    decode(item, bh, tile_h, tile_w, to0, to1);
    frame_range(to0, to1, t_first, t_last, t_end);
    {
      const int ho = tile_h * G::TH + hl, wo = tile_w * TW + cl0;
      row_cols = ho < p.Ho ? max(0, min(CPW, p.Wo - wo)) : 0;
      // the first emit of an item is output frame t_first - 1
      out_off = (((long long)bh * p.To + (t_first - 1)) * p.Ho * p.Wo + (long long)ho * p.Wo + wo) * 96;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
      for (int j = 0; j < CPW; ++j) axy[k][j] = make_float2(0.f, 0.f);
#pragma unroll
      for (int j = 0; j < ZC; ++j) az2[k][j] = make_float2(0.f, 0.f);
    }
    int t = t_first;
    while (true) {
      step(t, t_last, std::integral_constant<int, 0>{});
      if (++t > t_end) break;
      step(t, t_last, std::integral_constant<int, 1>{});
      if (++t > t_end) break;
      step(t, t_last, std::integral_constant<int, 2>{});
      if (++t > t_end) break;
    }
  }
}

template <int S>
static int launch(const void *qkv, int B, int heads, int T, int H, int W, const Stream *streams, int n_streams, float eps,
                  cudaStream_t st) {
  using G = Geo<S>;
  Params p{};
  for (int i = 0; i < n_streams; ++i) p.s[i] = streams[i];
  p.heads = heads;
  p.T = T;
  p.To = T;
  p.Ho = (H + 2 - 3) / S + 1;
  p.Wo = (W + 2 - 3) / S + 1;
  p.tiles_h = (p.Ho + G::TH - 1) / G::TH;
  p.tiles_w = (p.Wo + TW - 1) / TW;
  p.eps = eps;
  const int ctas_per_stream = std::max(1, 2 * num_sms() / n_streams);
  // split the frame axis while the items do not cover the CTAs about twice (each split re-reads one or two halo frames)
  const int64_t spatial = (int64_t)p.tiles_h * p.tiles_w * B * heads;
  p.t_per_item = p.To;
  while (p.t_per_item > 2 && spatial * ((p.To + p.t_per_item - 1) / p.t_per_item) < 2 * ctas_per_stream)
    p.t_per_item = (p.t_per_item + 1) / 2;
  p.t_splits = (p.To + p.t_per_item - 1) / p.t_per_item;
  const int64_t items = spatial * p.t_splits;
  MVIT_REQUIRE(items < (1ll << 31), "attention_pool_qkv: too many tiles");
  p.items = (int)items;

  CUtensorMap tmap;
  const uint64_t row = (uint64_t)3 * heads * 96;                              // elements per token of the qkv tensor
  const uint64_t dims[5] = {row, (uint64_t)W, (uint64_t)H, (uint64_t)T, (uint64_t)B};
  const uint64_t strides[4] = {row * 2, row * 2 * W, row * 2 * W * H, row * 2 * W * H * T};
  const uint32_t box[5] = {96, (uint32_t)G::NC, (uint32_t)G::NR, 1, 1};
  int r = encode_tmap_bf16(&tmap, qkv, 5, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE);
  if (r) return r;
  const size_t smem = (size_t)G::kStages * G::kStageBytes + (2 * G::kStages + 2 * kStgBufs * kConvWarps) * sizeof(uint64_t) +
                      kStgBufs * kConvWarps * sizeof(StgMeta) +
                      (192 + 27 * 16 * 2 + kConvWarps * kStgBufs * G::kStgFloats) * sizeof(float);
  MVIT_SMEM_OPT_IN(pool_tma_kernel<S>, smem);
  dim3 grid((unsigned)std::min<int64_t>(items, ctas_per_stream), (unsigned)n_streams);
  // Launched WITHOUT the PDL attribute (the kernel's pdl_wait() is then a no-op): started early, the q launch's persistent
  // CTAs take every SM slot before the k/v launch on the side stream becomes eligible and the two no longer share the SMs
  // (measured: 652 -> 638 clips/s with it, against +1 % for the GEMM / attention / fused-MLP launches).
  pool_tma_kernel<S><<<grid, kThreads, smem, st>>>(tmap, p);
  MVIT_LAUNCH_OK("attention_pool_qkv(tma)");
  return 0;
}

}  // namespace ptma

int pool_tma_fault_take() { return tc::tc_fault_take(); }

}  // namespace mvit

// q / k / v of one block.  Tensors whose stride is 1 or 2 go through the TMA kernel (equal strides share a launch); strides
// 4 and 8 (the K/V pools of the first two stages, which touch only (3/s)^2 of the tokens) keep the cp.async kernel of
// pool_tiled.cu that gathers just the touched columns.
extern "C" int mvit_attention_pool_qkv_fwd(const void *qkv, int B, int heads, int T, int H, int W, const float *const *weights,
                                           const float *const *gammas, const float *const *betas, const int *strides_hw,
                                           void *const *outs, void *const *pre_outs, float eps, int dtype, void *stream) {
  using namespace mvit;
  MVIT_REQUIRE(qkv && weights && gammas && betas && strides_hw && outs, "attention_pool_qkv: null pointer");
  MVIT_REQUIRE(dtype == MVIT_BF16, "attention_pool_qkv: bf16 only (fp32 tensors use mvit_attention_pool_fwd)");
  MVIT_REQUIRE(B > 0 && heads > 0 && T > 0 && H > 0 && W > 0, "attention_pool_qkv: bad shape");
  MVIT_REQUIRE((reinterpret_cast<uintptr_t>(qkv) & 15) == 0, "attention_pool_qkv: qkv must be 16-byte aligned");
  MVIT_REQUIRE((int64_t)B * heads < 65536, "attention_pool_qkv: B*heads too large");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  for (int i = 0; i < 3; ++i) {
    if (strides_hw[i] == 0) continue;                       // tensor not pooled by this call
    MVIT_REQUIRE(weights[i] && outs[i], "attention_pool_qkv: tensor %d needs a weight and an output", i);
    MVIT_REQUIRE((gammas[i] == nullptr) == (betas[i] == nullptr), "attention_pool_qkv: gamma/beta must both be set or NULL");
    const int s = strides_hw[i];
    MVIT_REQUIRE(s == 1 || s == 2 || s == 4 || s == 8, "attention_pool_qkv: stride (1,%d,%d) unsupported", s, s);
    MVIT_REQUIRE((reinterpret_cast<uintptr_t>(outs[i]) & 15) == 0, "attention_pool_qkv: outputs must be 16-byte aligned");
  }
  bool done[3] = {strides_hw[0] == 0, strides_hw[1] == 0, strides_hw[2] == 0};
  for (int i = 0; i < 3; ++i) {
    if (done[i]) continue;
    const int s = strides_hw[i];
    if (s <= 2) {
      ptma::Stream group[3];
      int n = 0;
      for (int j = i; j < 3; ++j)
        if (!done[j] && strides_hw[j] == s) {
          group[n++] = ptma::Stream{weights[j], gammas[j], betas[j], static_cast<bf16 *>(outs[j]),
                                    pre_outs ? static_cast<bf16 *>(pre_outs[j]) : nullptr, j * heads * 96};
          done[j] = true;
        }
      const int r = s == 1 ? ptma::launch<1>(qkv, B, heads, T, H, W, group, n, eps, st)
                           : ptma::launch<2>(qkv, B, heads, T, H, W, group, n, eps, st);
      if (r) return r;
    } else {
      PoolParams p;
      const int64_t row = (int64_t)3 * heads * 96;
      p.in_bs = (int64_t)T * H * W * row; p.in_ls = row; p.in_hs = 96;
      p.B = B; p.heads = heads; p.d = 96; p.T = T; p.H = H; p.W = W;
      p.kt = p.kh = p.kw = 3; p.st = 1; p.sh = p.sw = s; p.pt = p.ph = p.pw = 1;
      p.To = T; p.Ho = (H + 2 - 3) / s + 1; p.Wo = (W + 2 - 3) / s + 1;
      const int64_t Lo = (int64_t)p.To * p.Ho * p.Wo;
      p.out_bs = heads * Lo * 96; p.out_ls = 96; p.out_hs = Lo * 96;
      p.has_cls = 0; p.has_ln = gammas[i] ? 1 : 0; p.eps = eps;
      p.pre_out = pre_outs ? pre_outs[i] : nullptr;
      const bf16 *in = static_cast<const bf16 *>(qkv) + (int64_t)i * heads * 96;
      const int r = pool_tiled_try(in, weights[i], gammas[i], betas[i], outs[i], p, MVIT_POOL_CONV, MVIT_BF16, st);
      MVIT_REQUIRE(r <= 0, "attention_pool_qkv: tensor %d (stride %d) rejected by the tiled kernel (alignment?)", i, s);
      if (r < 0) return r;
      done[i] = true;
    }
  }
  return 0;
}

