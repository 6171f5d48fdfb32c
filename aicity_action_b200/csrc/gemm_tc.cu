// tcgen05 Linear for sm_100a:  y = epi(x·wᵀ + bias)·row_scale + residual      (bf16 in/out, fp32 accumulate)
//
// Replaces the cuBLAS GEMMs + separate bias / GELU / residual-add / DropPath kernels behind
// attention.py:231,281,426,443 and common.py:27-31 with one persistent, warp-specialised kernel:
//   warp 0      TMA producer: x tile [128 x 64] and w tile [BN x 64] (128B-swizzled, K-major) into a
//               kStages-deep shared-memory ring, completion on `full` mbarriers;
//   warp 1      MMA issuer: one elected lane issues tcgen05.mma (M=128, N=BN, K=16) x 4 per stage into a
//               double-buffered fp32 accumulator in TMEM; tcgen05.commit releases the smem slot / signals
//               the epilogue;
//   warp 2      TMEM allocator;
//   warp 3      residual producer: TMA-loads the residual tile (64B-swizzled 32-column boxes) into the
//               output ring kNB tiles ahead, so the skip-connection read never stalls the epilogue;
//   warps 4-11  epilogue (2 warps per TMEM lane quarter, half of the columns each): tcgen05.ld (lane = row)
//               -> +bias -> GELU -> *row_scale -> +residual (read in place from the output ring) -> bf16 ->
//               written back in place -> one elected thread TMA-stores the tile.
// All global traffic is TMA (fully coalesced, deep memory-level parallelism); the accumulator of tile
// i+1 is produced while tile i is drained.
#include <stdlib.h>

#include "linear.cuh"
#include "tc_common.cuh"

namespace mvit {
using namespace tc;

namespace gemm {
constexpr int BM = 128;
constexpr int BK = 64;                     // 64 bf16 = one 128-byte swizzle atom row
constexpr int UMMA_K = 16;
constexpr int kEpiWarp0 = 4;                             // warps 0-3: TMA, MMA, TMEM alloc, residual producer
constexpr int kBoxCols = 32;               // output / residual boxes: 128 rows x 32 cols (64 B), SWIZZLE_64B
constexpr int kBoxBytes = BM * kBoxCols * 2;

// Three tile configurations:
//   <96,  false> "narrow": memory-bound shapes (K <= 192): 4 operand stages, 4 in-place residual/output buffers;
//   <128, false> "wide":   compute-bound shapes whose N is not a multiple of 192;
//   <192, true>  "pair":   compute-bound shapes: a 2-CTA cluster computes a 256 x 192 tile with
//                tcgen05.mma.cta_group::2 — each CTA stages its own 128 rows of x and HALF (96 rows) of the w
//                tile, so operand traffic from L2 per FLOP drops by 1.75x (the 1-CTA kernel is L2->SM bound at
//                ~64 FLOP/B); accumulators live in both CTAs' TMEM, each CTA drains/stores its own 128 rows.
template <int BN, bool PAIR> struct Cfg {
  static constexpr int kStages = 4;
  static constexpr int kNB = (PAIR || BN == 128) ? 2 : 4;  // in-place residual/output tile buffers
  static constexpr int kABytes = BM * BK * 2;              // 16 KB
  static constexpr int kBRows = PAIR ? BN / 2 : BN;        // w rows staged by one CTA
  static constexpr int kBBytes = kBRows * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBoxes = BN / kBoxCols;
  static constexpr int kOutBytes = kBoxes * kBoxBytes;
  static constexpr int kTmemCols = BN == 192 ? 512 : 256;  // 2 accumulators of BN columns, power of two
  // Epilogue: TWO independent groups of 8 warps (4 TMEM lane quarters x 2 column halves).  Group g drains accumulator g,
  // i.e. every second tile of the CTA, with its own named barriers, bias slices, output buffers and store-issuing thread.
  // While one group sits in a latency phase (barrier, tcgen05.ld, waiting for its TMA store to release a buffer) the other
  // one computes: measured with ONE group of 12-16 warps in lock step the epilogue-bound layers (fc1 + GELU, the K <= 192
  // ones) issued on only 40-50 % of the cycles, a third of them real epilogue math.
  static constexpr int kGroups = 2;
  static constexpr int kGroupWarps = 8;
  static constexpr int kGroupThreads = kGroupWarps * 32;
  static constexpr int kEpiWarps = kGroups * kGroupWarps;
  static constexpr int kThreads = (kEpiWarp0 + kEpiWarps) * 32;   // 640
  static constexpr int kCols = BN / 2;                     // columns of one epilogue thread: 48 / 64 / 96
  static constexpr int kBatch = BN == 128 ? 32 : 48;       // columns held in registers at a time
  static constexpr int kBatches = kCols / kBatch;          // 1 / 2 / 2
  static constexpr int kBufsPerGroup = kNB / kGroups;      // output buffers a group rotates through: 2 (BN 96) or 1
  // bias | LayerNorm column-sum slices (per group, double-buffered) + per-half row statistics staging
  static constexpr int kVecBytes = kGroups * 4 * BN * 4;
  static constexpr int kStatBytes = kGroups * 2 * BM * 8;
  static constexpr int kLnMaxParts = 4;                    // N tiles of the producer (C <= 768 at 192 columns per tile)
  static constexpr int kLnBytes = kGroups * kLnMaxParts * BM * 8;   // kLnIn: input-row statistics, one slot per group
  // (the statistics staging of kLnStats and the statistics slots of kLnIn share their bytes: a launch runs one mode)
  static constexpr int kSmemBytes = kStages * kStageBytes + kNB * kOutBytes + kVecBytes + (kStatBytes > kLnBytes ? kStatBytes : kLnBytes) + 512 + 1024;
  static_assert(kSmemBytes <= 232448, "exceeds the 227 KB of dynamic shared memory a CTA can opt into");
};

// LayerNorm folding (eval path).  The LayerNorm that FOLLOWS a block-stream GEMM (proj+residual, fc2+residual, patch embed)
// and PRECEDES the next GEMM (qkv, fc1) is never run as a kernel:
//   kLnStats: the producer's epilogue also emits, per output row and N tile, (sum, sum of squares) of the bf16 values it
//             stores ([n_tiles][M][2] fp32, fixed summation order: reproducible);
//   kLnIn:    the consumer reads the raw rows x, multiplies by w' = w.diag(gamma) (folded by the host) and its epilogue
//             finishes the normalisation:  LN(x).w^T + b = rstd.(x.w'^T - mean.colsum(w')) + (b + w.beta).
constexpr int kLnNone = 0, kLnIn = 1, kLnStats = 2;

// Implicit-GEMM convolution mode (patch embedding): the A operand is a folded clip [B, Tf, Hf, Wf, Cf] read through a
// 5-D tensor map; an M tile is an 8(w) x 8(h) x 2(t) patch of output tokens and k-block kb = (tap, 64-channel slice):
// the TMA box of tap (dt, dh, dw) is the same patch shifted by the tap offset, zero-filled outside the clip.
struct ConvGeom {
  int enabled, Tf, Hf, Wf, cblocks, nt, nh, nw, lo_t, lo_h, lo_w;
};
struct Params {
  const float *bias, *row_scale;
  int64_t M, rows_per_sample, res_period;
  int N, K, epilogue, has_residual;
  ConvGeom conv;
  const float *colsum;        // kLnIn: sum_k w'[n, k]
  const float2 *ln_stats;     // kLnIn: [ln_parts][M] (sum, sum of squares) of the input rows
  float2 *stats_out;          // kLnStats: [n_tiles][M]
  int ln_parts;
  float ln_inv_c, ln_eps;
};
struct TileCoord { int b, t0, h0, w0; };
// (m tile, n tile) of a persistent CTA's tile sequence t = tile0, tile0 + step, ...; t = m * n_tiles + n.  Advanced
// incrementally: the 64-bit divisions of the closed form sat on every epilogue warp's critical path once per tile.
struct TileWalk {
  int m, n, dm, dn, n_tiles;
  __device__ __forceinline__ TileWalk(int64_t tile0, int64_t step, int n_tiles_) : n_tiles(n_tiles_) {
    m = (int)(tile0 / n_tiles_);
    n = (int)(tile0 % n_tiles_);
    dm = (int)(step / n_tiles_);
    dn = (int)(step % n_tiles_);
  }
  __device__ __forceinline__ void next() {
    m += dm;
    n += dn;
    if (n >= n_tiles) { n -= n_tiles; ++m; }
  }
};
__device__ __forceinline__ TileCoord conv_tile(const ConvGeom &g, int mt) {
  TileCoord c;
  const int wq = g.Wf / 8, hq = g.Hf / 8, tq = g.Tf / 2;
  c.w0 = (int)(mt % wq) * 8; mt /= wq;
  c.h0 = (int)(mt % hq) * 8; mt /= hq;
  c.t0 = (int)(mt % tq) * 2;
  c.b = (int)(mt / tq);
  return c;
}

// forward GELU: tc::gelu_tanh2 (tc_common.cuh)
__device__ __forceinline__ float2 gelu_fast2(float2 x) { return gelu_tanh2(x); }

// d/dx GELU(x) = Phi(x) + x*phi(x): Phi from the same polynomial, phi(x) = exp(-x^2/2)/sqrt(2 pi) with one MUFU.EX2.
__device__ __forceinline__ float2 gelu_grad_fast2(float2 x) {
  const float2 xc = make_float2(fminf(fmaxf(x.x, -4.f), 4.f), fminf(fmaxf(x.y, -4.f), 4.f));
  const float2 v = __fmul2_rn(xc, xc);
  float2 r = make_float2(-1.5806889130942636e-09f, -1.5806889130942636e-09f);
  r = __ffma2_rn(r, v, make_float2(1.2170519880783104e-07f, 1.2170519880783104e-07f));
  r = __ffma2_rn(r, v, make_float2(-4.100723799638217e-06f, -4.100723799638217e-06f));
  r = __ffma2_rn(r, v, make_float2(8.066566078923643e-05f, 8.066566078923643e-05f));
  r = __ffma2_rn(r, v, make_float2(-0.0010481934295967221f, -0.0010481934295967221f));
  r = __ffma2_rn(r, v, make_float2(0.009664841927587986f, 0.009664841927587986f));
  r = __ffma2_rn(r, v, make_float2(-0.06617535650730133f, -0.06617535650730133f));
  r = __ffma2_rn(r, v, make_float2(0.3988475501537323f, 0.3988475501537323f));
  const float2 cdf = __ffma2_rn(r, xc, make_float2(0.5f, 0.5f));
  const float2 u = __fmul2_rn(__fmul2_rn(x, x), make_float2(-0.72134752044448170368f, -0.72134752044448170368f));
  float ex, ey;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(u.x));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ey) : "f"(u.y));
  const float2 xpdf = __fmul2_rn(x, make_float2(0.39894228040143267794f * ex, 0.39894228040143267794f * ey));
  return __fadd2_rn(cdf, xpdf);
}

template <int BN, bool PAIR, int LNM>
__global__ void __launch_bounds__((Cfg<BN, PAIR>::kThreads), 1)
linear_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                 const __grid_constant__ CUtensorMap tmap_y, const __grid_constant__ CUtensorMap tmap_r, Params p) {
  using C = Cfg<BN, PAIR>;
  extern __shared__ uint8_t smem_raw[];
  // 1 KB alignment by an OFFSET in the shared window: the pointer keeps its address space, so every access below compiles
  // to LDS / STS instead of generic LD / ST
  uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t *sA = smem;
  uint8_t *sB = smem + C::kStages * C::kABytes;
  uint8_t *sOut = smem + C::kStages * C::kStageBytes;
  float *sVec = reinterpret_cast<float *>(sOut + C::kNB * C::kOutBytes);    // per group: [2][BN] bias, [2][BN] colsum
  float2 *sStatAll = reinterpret_cast<float2 *>(sVec + C::kGroups * 4 * BN); // per group: [2 halves][BM]
  float2 *sLnAll = sStatAll;                                                 // per group: [kLnMaxParts][BM] (other mode)
  uint64_t *bars = reinterpret_cast<uint64_t *>(sStatAll + (C::kStatBytes > C::kLnBytes ? C::kStatBytes : C::kLnBytes) / 8);
  uint64_t *full = bars, *empty = full + C::kStages, *tfull = empty + C::kStages, *tempty = tfull + 2;
  uint64_t *res_full = tempty + 2, *buf_free = res_full + C::kNB;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(buf_free + C::kNB);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // PAIR: the two CTAs of a cluster walk the same sequence of 256-row tiles; CTA `rank` owns rows [128*rank, +128)
  constexpr int TM = PAIR ? 2 * BM : BM;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0;
  const int64_t tile0 = PAIR ? blockIdx.x / 2 : blockIdx.x, tile_step = PAIR ? gridDim.x / 2 : gridDim.x;
  const int n_tiles = (p.N + BN - 1) / BN;
  const int64_t m_tiles = (p.M + TM - 1) / TM;
  const int64_t tiles = m_tiles * n_tiles;
  const int k_blocks = (p.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_w);
    tma_prefetch_desc(&tmap_y);
    if (p.has_residual) tma_prefetch_desc(&tmap_r);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < C::kStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], PAIR ? 2 * C::kGroupWarps : C::kGroupWarps);   // PAIR: both CTAs' groups release the leader
    }
    for (int i = 0; i < C::kNB; ++i) {
      mbar_init(&res_full[i], 1);
      mbar_init(&buf_free[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    if constexpr (PAIR) {
      tmem_alloc_pair(tmem_slot, C::kTmemCols);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(tmem_slot, C::kTmemCols);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all();   // barriers of BOTH CTAs are initialised before any remote arrive
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                 // the producer of x / residual / statistics has completed (see common.cuh)
  pdl_launch_dependents();

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (operands)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      TileWalk tw(tile0, tile_step, n_tiles);
      for (int64_t t = tile0; t < tiles; t += tile_step, tw.next()) {
        const int m0 = tw.m * TM + (int)rank * BM, n0 = tw.n * BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          if constexpr (PAIR) {
            // both CTAs' bytes land on the leader's barrier; the leader arms it for the whole 2-CTA stage
            if (rank == 0) mbar_arrive_expect_tx(&full[stage], 2 * C::kStageBytes);
            tma_load_2d_pair(sA + stage * C::kABytes, &tmap_x, &full[stage], kb * BK, m0);
            tma_load_2d_pair(sB + stage * C::kBBytes, &tmap_w, &full[stage], kb * BK, n0 + (int)rank * C::kBRows);
          } else {
            mbar_arrive_expect_tx(&full[stage], C::kStageBytes);
            if (p.conv.enabled) {
              const TileCoord tc = conv_tile(p.conv, tw.m);
              const int tap = kb / p.conv.cblocks, kc = kb - tap * p.conv.cblocks;
              const int dw = p.conv.lo_w + tap % p.conv.nw, dh = p.conv.lo_h + (tap / p.conv.nw) % p.conv.nh,
                        dt = p.conv.lo_t + tap / (p.conv.nw * p.conv.nh);
              tma_load_5d(sA + stage * C::kABytes, &tmap_x, &full[stage], kc * BK, tc.w0 + dw, tc.h0 + dh, tc.t0 + dt, tc.b);
            } else {
              tma_load_2d(sA + stage * C::kABytes, &tmap_x, &full[stage], kb * BK, m0);
            }
            tma_load_2d(sB + stage * C::kBBytes, &tmap_w, &full[stage], kb * BK, n0);
          }
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (PAIR: leader CTA only)
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(TM, BN, 0, 0);
      // descriptors of stage 0, built once; each MMA only advances the start-address field
      const uint64_t dsc_a = make_smem_desc(smem_u32(sA), 16, 1024, SWZ_128B);
      const uint64_t dsc_b = make_smem_desc(smem_u32(sB), 16, 1024, SWZ_128B);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int64_t t = tile0; t < tiles; t += tile_step) {
        mbar_wait(&tempty[acc], acc_phase ^ 1);       // epilogue(s) have drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint64_t da0 = desc_advance(dsc_a, stage * C::kABytes), db0 = desc_advance(dsc_b, stage * C::kBBytes);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t da = desc_advance(da0, k * UMMA_K * 2), db = desc_advance(db0, k * UMMA_K * 2);
            if constexpr (PAIR) umma_ss_pair(d_tmem, da, db, idesc, (kb | k) != 0);
            else umma_ss(d_tmem, da, db, idesc, (kb | k) != 0);
          }
          // smem slot free (in both CTAs) once these MMAs have read it
          if constexpr (PAIR) umma_commit_pair(&empty[stage], 0b11);
          else umma_commit(&empty[stage]);
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
        if constexpr (PAIR) umma_commit_pair(&tfull[acc], 0b11);   // accumulator complete (both halves)
        else umma_commit(&tfull[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp == 3) {
    // ------------------------------------------------------------ residual producer / output-ring gatekeeper
    if (lane == 0) {
      int buf = 0;
      uint32_t phase = 0;
      TileWalk tw(tile0, tile_step, n_tiles);
      for (int64_t t = tile0; t < tiles; t += tile_step, tw.next()) {
        mbar_wait(&buf_free[buf], phase ^ 1);         // the TMA store that last used this buffer has read it
        if (p.has_residual) {
          const int64_t mrow = (int64_t)tw.m * TM + rank * BM;
          const int m0 = (int)(p.res_period ? mrow % p.res_period : mrow), n0 = tw.n * BN;
          mbar_arrive_expect_tx(&res_full[buf], C::kOutBytes);
          if (p.conv.enabled) {                         // positional table [Tf, Hf, Wf, N], same patch, no batch axis
            const TileCoord tc = conv_tile(p.conv, tw.m);
#pragma unroll
            for (int bx = 0; bx < C::kBoxes; ++bx)
              tma_load_4d(sOut + buf * C::kOutBytes + bx * kBoxBytes, &tmap_r, &res_full[buf], n0 + bx * kBoxCols, tc.w0,
                          tc.h0, tc.t0);
          } else {
#pragma unroll
            for (int bx = 0; bx < C::kBoxes; ++bx)
              tma_load_2d(sOut + buf * C::kOutBytes + bx * kBoxBytes, &tmap_r, &res_full[buf], n0 + bx * kBoxCols, m0);
          }
        } else {
          mbar_arrive(&res_full[buf]);
        }
        if (++buf == C::kNB) { buf = 0; phase ^= 1; }
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ------------------------------------------------------------ epilogue: two groups, group g drains accumulator g
    const int e = warp - kEpiWarp0;
    const int g = e >> 3;                              // group
    const int q = e & 3;                               // TMEM lane quarter (== warp % 4)
    const int hf = (e >> 2) & 1;                       // column half of the tile
    const int gt = threadIdx.x - (kEpiWarp0 + g * C::kGroupWarps) * 32;   // 0..255 inside the group
    const int row = q * 32 + lane;                     // row inside the tile == TMEM lane
    const uint32_t swz = (uint32_t)((row >> 1) & 3);
    const int bar_a = 1 + 2 * g, bar_b = 2 + 2 * g;    // named barriers of this group
    float *sBias = sVec + g * 4 * BN, *sCol = sBias + 2 * BN;
    float2 *sStat = sStatAll + g * 2 * BM, *sLn = sLnAll + g * C::kLnMaxParts * BM;
    // this group's tiles: local iterations it = g, g + 2, ... of the CTA's sequence; k counts the group's own tiles
    const int64_t gstep = 2 * tile_step;
    TileWalk tw(tile0 + g * tile_step, gstep, n_tiles), tw_next = tw;
    tw_next.next();
    const int64_t t_first = tile0 + g * tile_step;
    // bias slice of the first tile; later slices are fetched one tile ahead (global latency off the critical path)
    if (gt < BN) {
      const int n = tw.n * BN + gt;
      sBias[gt] = (p.bias && t_first < tiles && n < p.N) ? p.bias[n] : 0.f;
      if constexpr (LNM == kLnIn) sCol[gt] = (t_first < tiles && n < p.N) ? p.colsum[n] : 0.f;
    }
    // kLnIn: (sum, sum of squares) of the 128 input rows of a tile, one slice per N tile of the producer, fetched a tile
    // ahead with cp.async by the group's first four warps (gt == row): no thread waits on the L2 round trip
    auto fetch_row_stats = [&](int m_tile) {
      if (gt < BM) {
        const int64_t m = min((int64_t)m_tile * TM + rank * BM + gt, p.M - 1);
        for (int i = 0; i < p.ln_parts; ++i)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(sLn + i * BM + gt)),
                       "l"(p.ln_stats + (int64_t)i * p.M + m) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if constexpr (LNM == kLnIn) {
      if (t_first < tiles) fetch_row_stats(tw.m);
    }
    int k = 0;
    for (int64_t t = t_first; t < tiles; t += gstep, ++k, tw.next(), tw_next.next()) {
      const int it = 2 * k + g;                        // position in the CTA's tile sequence
      const int buf = it % C::kNB;
      const uint32_t buf_phase = (uint32_t)(it / C::kNB) & 1, acc_phase = (uint32_t)k & 1;
      const int64_t m0 = (int64_t)tw.m * TM + rank * BM;
      const int n0 = tw.n * BN;
      const float *bias_s = sBias + (k & 1) * BN;
      const float *col_s = sCol + (k & 1) * BN;
      float bias_next = 0.f, col_next = 0.f;
      const int64_t tn = t + gstep;
      if (gt < BN && tn < tiles) {
        const int n = tw_next.n * BN + gt;
        if (n < p.N) {
          if (p.bias) bias_next = p.bias[n];
          if constexpr (LNM == kLnIn) col_next = p.colsum[n];
        }
      }
      float2 sum2 = make_float2(0.f, 0.f), sq2 = make_float2(0.f, 0.f);   // kLnStats
      if constexpr (LNM == kLnIn) asm volatile("cp.async.wait_group 0;" ::: "memory");   // this tile's statistics have landed
      float rs = 1.f;
      if (p.row_scale) rs = p.row_scale[min(m0 + row, p.M - 1) / p.rows_per_sample];
      // this tile's bias slice (and row statistics) are visible; the group's previous store was issued
      asm volatile("bar.sync %0, %1;" ::"r"(bar_a), "n"(C::kGroupThreads) : "memory");
      float ln_a = 1.f, ln_b = 0.f;                      // x_norm . w'^T = ln_a * acc + ln_b * colsum
      if constexpr (LNM == kLnIn) {
        float2 st = sLn[row];
        for (int i = 1; i < p.ln_parts; ++i) {           // fixed order: reproducible
          st.x += sLn[i * BM + row].x;
          st.y += sLn[i * BM + row].y;
        }
        const float mean = st.x * p.ln_inv_c;
        const float var = fmaxf(st.y * p.ln_inv_c - mean * mean, 0.f);
        ln_a = rsqrtf(var + p.ln_eps);
        ln_b = -mean * ln_a;
      }
      mbar_wait(&res_full[buf], buf_phase);            // buffer is ours (and holds the residual tile, if any)
      mbar_wait(&tfull[g], acc_phase);
      tc_fence_after();
      uint8_t *obuf = sOut + buf * C::kOutBytes;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + g * BN + hf * C::kCols;
#pragma unroll
      for (int bt = 0; bt < C::kBatches; ++bt) {
        uint32_t rr[C::kBatch / 16][16];
#pragma unroll
        for (int c = 0; c < C::kBatch / 16; ++c) tmem_ld16(taddr + bt * C::kBatch + c * 16, rr[c]);
        const uint32_t *r = &rr[0][0];
        tmem_ld_wait();
        if (bt == C::kBatches - 1) {
          // accumulator fully read -> hand it back to the MMA warp before the math of the last batch
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (PAIR && rank != 0) mbar_arrive_remote(&tempty[g], 0);   // the leader's MMA warp owns the accumulators
            else mbar_arrive(&tempty[g]);
          }
        }
#pragma unroll
        for (int v = 0; v < C::kBatch / 8; ++v) {
          const int col = hf * C::kCols + bt * C::kBatch + v * 8;   // first tile column of these 8 values
          const float4 b0 = *reinterpret_cast<const float4 *>(bias_s + col);
          const float4 b1 = *reinterpret_cast<const float4 *>(bias_s + col + 4);
          float2 x[4];
          if constexpr (LNM == kLnIn) {
            const float4 c0 = *reinterpret_cast<const float4 *>(col_s + col);
            const float4 c1 = *reinterpret_cast<const float4 *>(col_s + col + 4);
            const float2 m2 = make_float2(ln_b, ln_b);
            const float2 sh[4] = {__ffma2_rn(m2, make_float2(c0.x, c0.y), make_float2(b0.x, b0.y)),
                                  __ffma2_rn(m2, make_float2(c0.z, c0.w), make_float2(b0.z, b0.w)),
                                  __ffma2_rn(m2, make_float2(c1.x, c1.y), make_float2(b1.x, b1.y)),
                                  __ffma2_rn(m2, make_float2(c1.z, c1.w), make_float2(b1.z, b1.w))};
            // scalar FMAs on the accumulator words: pairing the tcgen05.ld destination registers for a packed FMA costs a
            // register move per element
#pragma unroll
            for (int j = 0; j < 4; ++j)
              x[j] = make_float2(fmaf(__uint_as_float(r[v * 8 + 2 * j]), ln_a, sh[j].x),
                                 fmaf(__uint_as_float(r[v * 8 + 2 * j + 1]), ln_a, sh[j].y));
          } else {
            x[0] = make_float2(__uint_as_float(r[v * 8 + 0]) + b0.x, __uint_as_float(r[v * 8 + 1]) + b0.y);
            x[1] = make_float2(__uint_as_float(r[v * 8 + 2]) + b0.z, __uint_as_float(r[v * 8 + 3]) + b0.w);
            x[2] = make_float2(__uint_as_float(r[v * 8 + 4]) + b1.x, __uint_as_float(r[v * 8 + 5]) + b1.y);
            x[3] = make_float2(__uint_as_float(r[v * 8 + 6]) + b1.z, __uint_as_float(r[v * 8 + 7]) + b1.w);
          }
          if (p.epilogue == MVIT_EPI_GELU) {
#pragma unroll
            for (int j = 0; j < 4; ++j) x[j] = gelu_fast2(x[j]);
          } else if (p.epilogue == MVIT_EPI_GELU_GRAD) {
#pragma unroll
            for (int j = 0; j < 4; ++j) x[j] = gelu_grad_fast2(x[j]);
          }
          if (p.row_scale) {
#pragma unroll
            for (int j = 0; j < 4; ++j) x[j] = __fmul2_rn(x[j], make_float2(rs, rs));
          }
          uint4 *slot = reinterpret_cast<uint4 *>(obuf + (col >> 5) * kBoxBytes + row * 64 +
                                                  ((((uint32_t)(col & 31) >> 3) ^ swz) << 4));
          if (p.has_residual) {
            const uint4 rv = *slot;
            const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
            if (p.epilogue == MVIT_EPI_GELU_GRAD) {       // the "residual" is the upstream gradient: multiply
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                x[j].x *= __uint_as_float(rw[j] << 16);
                x[j].y *= __uint_as_float(rw[j] & 0xffff0000u);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                x[j].x += __uint_as_float(rw[j] << 16);
                x[j].y += __uint_as_float(rw[j] & 0xffff0000u);
              }
            }
          }
          uint32_t o[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            __nv_bfloat162 h = __floats2bfloat162_rn(x[j].x, x[j].y);
            o[j] = *reinterpret_cast<uint32_t *>(&h);
            if constexpr (LNM == kLnStats) {                     // statistics of exactly the values the next GEMM reads
              const float2 f = make_float2(__uint_as_float(o[j] << 16), __uint_as_float(o[j] & 0xffff0000u));
              sum2 = __fadd2_rn(sum2, f);
              sq2 = __ffma2_rn(f, f, sq2);
            }
          }
          *slot = make_uint4(o[0], o[1], o[2], o[3]);
        }
      }
      if constexpr (LNM == kLnStats) sStat[hf * BM + row] = make_float2(sum2.x + sum2.y, sq2.x + sq2.y);
      if (gt < BN) {                                             // readers of that half finished a tile ago
        sBias[((k + 1) & 1) * BN + gt] = bias_next;
        if constexpr (LNM == kLnIn) sCol[((k + 1) & 1) * BN + gt] = col_next;
      }
      // tile is complete in shared memory -> one thread of the group TMA-stores it
      fence_proxy_async_smem();
      asm volatile("bar.sync %0, %1;" ::"r"(bar_b), "n"(C::kGroupThreads) : "memory");
      if constexpr (LNM == kLnIn) {
        if (tn < tiles) fetch_row_stats(tw_next.m);              // every thread of the group has read the slot (bar_b)
      }
      if constexpr (LNM == kLnStats) {
        if (gt < BM) {                                           // gt == row for the first four warps of the group
          const float2 v0 = sStat[gt], v1 = sStat[BM + gt];
          const float2 tot = make_float2(v0.x + v1.x, v0.y + v1.y);
          int64_t token = m0 + gt;
          if (p.conv.enabled) {                                  // tile row -> token of the 8(w) x 8(h) x 2(t) patch
            const TileCoord tc = conv_tile(p.conv, tw.m);
            token = (((int64_t)tc.b * p.conv.Tf + tc.t0 + (gt >> 6)) * p.conv.Hf + tc.h0 + ((gt >> 3) & 7)) * p.conv.Wf +
                    tc.w0 + (gt & 7);
          }
          if (token < p.M) p.stats_out[(int64_t)tw.n * p.M + token] = tot;
        }
      }
      if (gt == 0) {
        if (p.conv.enabled) {
          const TileCoord tc = conv_tile(p.conv, tw.m);
#pragma unroll
          for (int bx = 0; bx < C::kBoxes; ++bx)
            if (n0 + bx * kBoxCols < p.N)
              tma_store_5d(&tmap_y, obuf + bx * kBoxBytes, n0 + bx * kBoxCols, tc.w0, tc.h0, tc.t0, tc.b);
        } else {
#pragma unroll
          for (int bx = 0; bx < C::kBoxes; ++bx)
            if (n0 + bx * kBoxCols < p.N) tma_store_2d(&tmap_y, obuf + bx * kBoxBytes, n0 + bx * kBoxCols, (int)m0);
        }
        tma_store_commit();
        // the group keeps kBufsPerGroup - 1 stores in flight; the older one has drained its buffer -> hand it back
        if (k >= C::kBufsPerGroup - 1) {
          tma_store_wait_read<C::kBufsPerGroup - 1>();
          mbar_arrive(&buf_free[(it - 2 * (C::kBufsPerGroup - 1)) % C::kNB]);
        }
      }
    }
    if (gt == 0) tma_store_wait_all<0>();              // global writes complete before the CTA retires
  }
  tc_fence_before();
  if constexpr (PAIR) {
    cluster_sync_all();                                // both CTAs are done with TMEM, smem and remote barriers
    if (warp == 2) tmem_dealloc_pair(tmem_base, C::kTmemCols);
  } else {
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, C::kTmemCols);
  }
}

}  // namespace gemm

// ---------------------------------------------------------------- host side
namespace tc {
typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                             const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                             CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeFn get_encode() {
  static EncodeFn fn = nullptr;
  if (!fn) {
    void *sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeFn>(sym);
  }
  return fn;
}

int encode_tmap_bf16(CUtensorMap *out, const void *base, int rank, const uint64_t *dims,
                     const uint64_t *strides_bytes, const uint32_t *box, CUtensorMapSwizzle swizzle) {
  EncodeFn fn = get_encode();
  MVIT_REQUIRE(fn, "cuTensorMapEncodeTiled is not available from the installed driver");
  cuuint64_t d[5], s[5];
  cuuint32_t b[5], e[5];
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; e[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) s[i] = strides_bytes[i];
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void *>(base), d, s, b, e,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MVIT_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return 0;
}
}  // namespace tc

bool linear_tc_supported(const LinearArgs &a, const char **why) {
  if (a.K % 8 != 0) { *why = "K must be a multiple of 8 (16-byte TMA row pitch)"; return false; }
  if (a.N % 8 != 0) { *why = "N must be a multiple of 8 (16-byte output vectors)"; return false; }
  if (a.ldy % 8 != 0 || (a.residual && a.ldr % 8 != 0)) { *why = "leading dimensions must be multiples of 8"; return false; }
  if (a.M >= ((int64_t)1 << 31)) { *why = "M too large"; return false; }
  auto al = [](const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (!al(a.x) || !al(a.w) || !al(a.y) || (a.residual && !al(a.residual))) { *why = "pointers must be 16-byte aligned"; return false; }
  if (a.res_period % gemm::BM != 0) { *why = "residual_row_period must be a multiple of 128"; return false; }
  return true;
}

template <int BN, bool PAIR, int LNM>
static int launch_tc(const LinearArgs &a, cudaStream_t st) {
  using C = gemm::Cfg<BN, PAIR>;
  CUtensorMap tx, tw, ty, tr;
  auto enc2 = [&](CUtensorMap *m, const void *ptr, uint64_t cols, uint64_t rows, uint64_t pitch_elems, uint32_t bc,
                  uint32_t br, CUtensorMapSwizzle sw) {
    const uint64_t dims[2] = {cols, rows};
    const uint64_t strides[1] = {pitch_elems * 2};
    const uint32_t box[2] = {bc, br};
    return encode_tmap_bf16(m, ptr, 2, dims, strides, box, sw);
  };
  int r;
  if ((r = enc2(&tx, a.x, a.K, a.M, a.K, gemm::BK, gemm::BM, CU_TENSOR_MAP_SWIZZLE_128B))) return r;
  if ((r = enc2(&tw, a.w, a.K, a.N, a.K, gemm::BK, C::kBRows, CU_TENSOR_MAP_SWIZZLE_128B))) return r;
  if ((r = enc2(&ty, a.y, a.N, a.M, a.ldy, gemm::kBoxCols, gemm::BM, CU_TENSOR_MAP_SWIZZLE_64B))) return r;
  if (a.residual) {
    const uint64_t res_rows = a.res_period ? (uint64_t)a.res_period : (uint64_t)a.M;
    if ((r = enc2(&tr, a.residual, a.N, res_rows, a.ldr, gemm::kBoxCols, gemm::BM, CU_TENSOR_MAP_SWIZZLE_64B))) return r;
  } else {
    tr = ty;
  }
  MVIT_SMEM_OPT_IN((gemm::linear_tc_kernel<BN, PAIR, LNM>), C::kSmemBytes);
  gemm::Params p{a.bias, a.row_scale, a.M, a.rows_per_sample, a.res_period, a.N, a.K, a.epilogue, a.residual ? 1 : 0, {},
                 a.colsum, reinterpret_cast<const float2 *>(a.ln_stats), reinterpret_cast<float2 *>(a.stats_out),
                 a.ln_parts, 1.0f / (float)a.K, a.ln_eps};
  constexpr int TM = PAIR ? 2 * gemm::BM : gemm::BM;
  const int64_t tiles = ((a.M + TM - 1) / TM) * ((a.N + BN - 1) / BN);
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute attr[2];
  int n_attr = 0;
  if (PAIR) {
    cfg.gridDim = dim3(2 * (unsigned)std::min<int64_t>(tiles, num_sms() / 2));
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    n_attr = 1;
  } else {
    cfg.gridDim = dim3((unsigned)std::min<int64_t>(tiles, num_sms()));
  }
  n_attr = pdl_attr(attr, n_attr);
  cfg.attrs = attr;
  cfg.numAttrs = (unsigned)n_attr;
  cfg.blockDim = dim3(C::kThreads);
  cfg.dynamicSmemBytes = C::kSmemBytes;
  cfg.stream = st;
  MVIT_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm::linear_tc_kernel<BN, PAIR, LNM>, tx, tw, ty, tr, p));
  return 0;
}

// Patch-embedding convolution as an implicit GEMM over the folded clip (see ConvGeom).  M = B*Tf*Hf*Wf tokens,
// K = taps * Cf, N = Cout; tile = <96, false>.
int patch_conv_tc(const void *folded, const void *wf, const float *bias, const void *pos, void *out, float *stats_out, int B,
                  int Tf, int Hf, int Wf, int Cf, int nt, int nh, int nw, int lo_t, int lo_h, int lo_w, int N, cudaStream_t st) {
  using C = gemm::Cfg<96, false>;
  MVIT_REQUIRE(Tf % 2 == 0 && Hf % 8 == 0 && Wf % 8 == 0, "patch_conv: token grid must be a multiple of 2x8x8");
  MVIT_REQUIRE(Cf % gemm::BK == 0 && N % 8 == 0, "patch_conv: folded channels must be a multiple of 64, Cout of 8");
  CUtensorMap tx, tw, ty, tr;
  int r;
  {
    const uint64_t dims[5] = {(uint64_t)Cf, (uint64_t)Wf, (uint64_t)Hf, (uint64_t)Tf, (uint64_t)B};
    const uint64_t str[4] = {(uint64_t)Cf * 2, (uint64_t)Wf * Cf * 2, (uint64_t)Hf * Wf * Cf * 2, (uint64_t)Tf * Hf * Wf * Cf * 2};
    const uint32_t box[5] = {gemm::BK, 8, 8, 2, 1};
    if ((r = encode_tmap_bf16(&tx, folded, 5, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B))) return r;
  }
  const int K = nt * nh * nw * Cf;
  {
    const uint64_t dims[2] = {(uint64_t)K, (uint64_t)N};
    const uint64_t str[1] = {(uint64_t)K * 2};
    const uint32_t box[2] = {gemm::BK, 96};
    if ((r = encode_tmap_bf16(&tw, wf, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B))) return r;
  }
  {
    const uint64_t dims[5] = {(uint64_t)N, (uint64_t)Wf, (uint64_t)Hf, (uint64_t)Tf, (uint64_t)B};
    const uint64_t str[4] = {(uint64_t)N * 2, (uint64_t)Wf * N * 2, (uint64_t)Hf * Wf * N * 2, (uint64_t)Tf * Hf * Wf * N * 2};
    const uint32_t box[5] = {gemm::kBoxCols, 8, 8, 2, 1};
    if ((r = encode_tmap_bf16(&ty, out, 5, dims, str, box, CU_TENSOR_MAP_SWIZZLE_64B))) return r;
    if (pos) {
      const uint32_t box4[4] = {gemm::kBoxCols, 8, 8, 2};
      if ((r = encode_tmap_bf16(&tr, pos, 4, dims, str, box4, CU_TENSOR_MAP_SWIZZLE_64B))) return r;
    } else {
      tr = ty;
    }
  }
  const int64_t M = (int64_t)B * Tf * Hf * Wf;
  gemm::Params p{bias, nullptr, M, 0, 0, N, K, MVIT_EPI_NONE, pos ? 1 : 0,
                 {1, Tf, Hf, Wf, Cf / gemm::BK, nt, nh, nw, lo_t, lo_h, lo_w},
                 nullptr, nullptr, reinterpret_cast<float2 *>(stats_out), 0, 0.f, 0.f};
  const int64_t tiles = (M / gemm::BM) * ((N + 95) / 96);
  const unsigned grid = (unsigned)std::min<int64_t>(tiles, num_sms());
  if (stats_out) {
    MVIT_SMEM_OPT_IN((gemm::linear_tc_kernel<96, false, gemm::kLnStats>), C::kSmemBytes);
    MVIT_CUDA_OK(launch_pdl(gemm::linear_tc_kernel<96, false, gemm::kLnStats>, dim3(grid), dim3(C::kThreads), C::kSmemBytes, st,
                            tx, tw, ty, tr, p));
  } else {
    MVIT_SMEM_OPT_IN((gemm::linear_tc_kernel<96, false, gemm::kLnNone>), C::kSmemBytes);
    MVIT_CUDA_OK(launch_pdl(gemm::linear_tc_kernel<96, false, gemm::kLnNone>, dim3(grid), dim3(C::kThreads), C::kSmemBytes, st,
                            tx, tw, ty, tr, p));
  }
  MVIT_LAUNCH_OK("patch_conv(tcgen05)");
  return 0;
}

// tile configuration: compute-bound shapes (long K): CTA pairs on 256x192 tiles, else 128x128; memory-bound shapes:
// 128x96 tiles with a deeper output ring
static int pick_cfg(int64_t M, int N, int K) {
  static const int forced = [] {
    const char *e = getenv("MVIT_GEMM_CFG");      // experiments only: force a tile configuration where it is legal
    return e ? atoi(e) : -1;
  }();
  if (forced == 2 && N % 192 == 0) return 2;
  if (forced == 1 && N % 128 == 0) return 1;
  if (forced == 0) return 0;
  if (K >= 384 && N % 192 == 0 && ((M + 255) / 256) * (N / 192) >= num_sms() / 2) return 2;
  if (K >= 384 && N % 128 == 0 && ((M + 127) / 128) * (N / 128) >= num_sms()) return 1;
  return 0;
}

int linear_tc_stat_parts(int64_t M, int N, int K) {
  const int bn = pick_cfg(M, N, K) == 2 ? 192 : (pick_cfg(M, N, K) == 1 ? 128 : 96);
  return (N + bn - 1) / bn;
}

template <int LNM>
static int linear_tc_mode(const LinearArgs &a, cudaStream_t st) {
  switch (pick_cfg(a.M, a.N, a.K)) {
    case 2: return launch_tc<192, true, LNM>(a, st);
    case 1: return launch_tc<128, false, LNM>(a, st);
    default: return launch_tc<96, false, LNM>(a, st);
  }
}

int linear_tc(const LinearArgs &a, cudaStream_t st) {
  MVIT_REQUIRE(!(a.ln_stats && a.stats_out), "linear: LayerNorm-folded input and row statistics output cannot be combined");
  if (a.ln_stats) {
    MVIT_REQUIRE(a.colsum && a.ln_parts > 0 && a.ln_parts <= 4, "linear: ln_stats needs colsum and 1..4 parts");
    MVIT_REQUIRE(a.epilogue != MVIT_EPI_GELU_GRAD, "linear: the LayerNorm-folded input is a forward-only form");
    return linear_tc_mode<gemm::kLnIn>(a, st);
  }
  if (a.stats_out) {
    MVIT_REQUIRE(a.epilogue == MVIT_EPI_NONE, "linear: row statistics are emitted by plain (+residual) epilogues only");
    MVIT_REQUIRE(a.ldy == a.N, "linear: row statistics need a dense output (ldy == N)");
    return linear_tc_mode<gemm::kLnStats>(a, st);
  }
  return linear_tc_mode<gemm::kLnNone>(a, st);
}

int gemm_tc_fault_take() { return tc_fault_take(); }

}  // namespace mvit

/* See include/mvit_b200.h. */
extern "C" int mvit_linear_stat_parts(int64_t M, int N, int K) {
  if (M < 0 || N <= 0 || K <= 0) return -1;
  return mvit::linear_tc_stat_parts(M, N, K);
}

extern "C" int mvit_linear_ln_fwd(const void *x, const void *w, const float *bias, const float *colsum,
                                  const float *ln_stats, int ln_parts, float ln_eps, const void *residual,
                                  const float *row_scale, int64_t rows_per_sample, void *y, float *stats_out, int64_t M,
                                  int N, int K, int64_t ldy, int64_t ldr, int epilogue, void *stream) {
  using namespace mvit;
  MVIT_REQUIRE(x && w && y, "linear_ln: null pointer");
  MVIT_REQUIRE(M >= 0 && N > 0 && K > 0, "linear_ln: bad shape M=%lld N=%d K=%d", (long long)M, N, K);
  MVIT_REQUIRE(epilogue == MVIT_EPI_NONE || epilogue == MVIT_EPI_GELU, "linear_ln: epilogue must be NONE or GELU");
  MVIT_REQUIRE((ln_stats != nullptr) != (stats_out != nullptr), "linear_ln: exactly one of ln_stats / stats_out must be set");
  MVIT_REQUIRE(ldy >= N && (!residual || ldr >= N), "linear_ln: leading dimension smaller than N");
  MVIT_REQUIRE(!row_scale || rows_per_sample > 0, "linear_ln: row_scale needs rows_per_sample");
  auto al8 = [](const void *p) { return (reinterpret_cast<uintptr_t>(p) & 7) == 0; };
  MVIT_REQUIRE(al8(ln_stats) && al8(stats_out), "linear_ln: statistics buffers must be 8-byte aligned");
  if (M == 0) return 0;
  LinearArgs a{x, w, residual, bias, row_scale, y, M, rows_per_sample, ldy, ldr, 0, N, K, epilogue};
  a.colsum = colsum;
  a.ln_stats = ln_stats;
  a.ln_parts = ln_parts;
  a.ln_eps = ln_eps;
  a.stats_out = stats_out;
  const char *why = "";
  MVIT_REQUIRE(linear_tc_supported(a, &why), "linear_ln: tcgen05 path rejected: %s (there is no other implementation)", why);
  return linear_tc(a, static_cast<cudaStream_t>(stream));
}
