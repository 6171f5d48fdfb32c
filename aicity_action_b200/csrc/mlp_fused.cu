// Fused MLP of a MultiScaleBlock for the memory-bound front stages (C = 96, 192), eval path, bf16:
//
//     y = x + fc2( GELU( fc1( LayerNorm(x) ) ) )                       (attention.py:436-445, common.py:26-33)
//
// in ONE kernel.  The 4C-wide hidden activation never leaves the SM: as two GEMM launches it was written by fc1 and read
// back by fc2 (616 MB each way per forward at block 0, 308 MB at blocks 1-2) and those layers sat on the HBM write
// ceiling.  The structure is the attention kernel's (S -> P -> PV) with GELU in place of the softmax:
//
//   hidden is cut into chunks of HC units.  Per chunk g:  S_g = x . W1'[g]^T   (tcgen05.mma SS, M128 x N=HC x K=C) into one of
//   two TMEM score buffers;  eight epilogue warps (thread = row, two column halves per TMEM lane quarter) load S_g, finish
//   the folded LayerNorm (rstd*(acc - mean*colsum) + b1', see gemm_tc.cu kLnIn), apply GELU and write P_g (bf16) to a TMEM
//   operand buffer;  O += P_g . W2[:, g]^T   (tcgen05.mma TS, A from TMEM, N = C, K = HC).  fc1 of chunk g+2 is issued right
//   after fc2 of chunk g, so the tensor pipe, the GELU warps and the weight loads of three chunks overlap; chunks run on
//   seamlessly across the 128-row tiles a persistent CTA walks.
//   After the last chunk of a tile the same warps add b2 and the residual — the x tile still in shared memory — IN PLACE,
//   emit the row statistics the next block's folded norm1 needs, and one thread TMA-stores the tile.
//
//   warp 0: TMA producer for x tiles (double-buffered) and W1' chunks;  warp 3: TMA producer for W2 chunks (a W1 stage is
//   released as soon as fc1 has read it, a W2 stage only after fc2: separate rings, separate producers);  warp 1: MMA
//   issuer;  warp 2: TMEM allocator;  warps 4-11: GELU / output epilogue.
//
// TMEM (512 columns allocated):  C = 96 (HC = 128): S0 S1 | P0 P1 | O = 128+128+64+64+96 = 480;  C = 192 (HC = 64): 64+64+32+32+192.
// Shared memory: 2 x-tile buffers (C*256 B each) + 2 stages x (W1' chunk 24 KB + W2 chunk 24 KB) + 3 H-vectors.
#include "linear.cuh"
#include "tc_common.cuh"

namespace mvit {
using namespace tc;

namespace mlpf {
constexpr int BM = 128;
constexpr int kChunkCols = 32;                      // every operand is staged as 32-column (64-byte) chunks, SWIZZLE_64B
constexpr int kThreads = 384;
constexpr int kEpiWarp0 = 4, kEpiThreads = 256;
constexpr int kStages = 2;

template <int C> struct Cfg {
  static constexpr int H = 4 * C;
  static constexpr int HC = C == 96 ? 128 : 64;     // hidden units per chunk
  static constexpr int NC = H / HC;                 // chunks per tile: 3 / 12
  static constexpr int kXChunks = C / kChunkCols;   // 3 / 6
  static constexpr int kXChunkBytes = BM * 64;      // 8 KB
  static constexpr int kXTileBytes = kXChunks * kXChunkBytes;        // 24 / 48 KB
  static constexpr int kW1ChunkBytes = HC * 64;                       // [HC rows x 32 cols]
  static constexpr int kW1Bytes = kXChunks * kW1ChunkBytes;           // 24 KB
  static constexpr int kW2Chunks = HC / kChunkCols;                   // 4 / 2
  static constexpr int kW2ChunkBytes = C * 64;                        // [C rows x 32 cols]
  static constexpr int kW2Bytes = kW2Chunks * kW2ChunkBytes;          // 24 KB
  static constexpr uint32_t kColS = 0, kColP = 2 * HC, kColO = 2 * HC + HC;   // S0 S1 | P0 P1 (HC/2 each) | O
  static constexpr int kHalf = HC / 2;              // columns of one epilogue warp per chunk: 64 / 32
  static constexpr int kOB = C / 6;                 // output columns per epilogue batch: 16 / 32
  static constexpr int kSmemBytes = 2 * kXTileBytes + kStages * (kW1Bytes + kW2Bytes) + (2 * H + C) * 4 + 2 * BM * 8 + 256 + 1024;
  static_assert(kColO + C <= 512, "TMEM columns");
  static_assert(kSmemBytes <= 232448, "shared memory");
};

struct Params {
  const float *b1, *colsum, *b2;   // [H], [H], [C]
  const float2 *ln_stats;          // [ln_parts][M]
  float2 *stats_out;               // [M] or NULL
  int64_t M;
  int ln_parts;
  float ln_inv_c, ln_eps;
};

__device__ __forceinline__ uint32_t pack_bf16x2_alu(float lo, float hi) {   // round half up on the integer pipe
  return __byte_perm(__float_as_uint(lo) + 0x8000u, __float_as_uint(hi) + 0x8000u, 0x7632);
}

template <int C>
__global__ void __launch_bounds__(kThreads, 1)
mlp_fused_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w1,
                 const __grid_constant__ CUtensorMap tmap_w2, const __grid_constant__ CUtensorMap tmap_y, Params p) {
  using G = Cfg<C>;
  constexpr int HC = G::HC, NC = G::NC;
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t *sX = smem;                                                   // [2][x tile]
  uint8_t *sW1 = sX + 2 * G::kXTileBytes;                               // [stage][24 KB]
  uint8_t *sW2 = sW1 + kStages * G::kW1Bytes;                           // [stage][24 KB]
  float *sB1 = reinterpret_cast<float *>(sW2 + kStages * G::kW2Bytes);  // [H]
  float *sCs = sB1 + G::H;                                              // [H]
  float *sB2 = sCs + G::H;                                              // [C]
  float2 *sStat = reinterpret_cast<float2 *>(sB2 + C);                  // [2 halves][BM]
  uint64_t *bars = reinterpret_cast<uint64_t *>(sStat + 2 * BM);
  uint64_t *x_full = bars, *x_empty = x_full + 2;
  uint64_t *w1_full = x_empty + 2, *w1_empty = w1_full + kStages;
  uint64_t *w2_full = w1_empty + kStages, *w2_empty = w2_full + kStages;
  uint64_t *s_full = w2_empty + kStages, *p_ready = s_full + 2;
  uint64_t *o_full = p_ready + 2, *o_empty = o_full + 1;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(o_empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t m_tiles = (p.M + BM - 1) / BM;
  const int64_t my_tiles = blockIdx.x < m_tiles ? (m_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int64_t my_chunks = my_tiles * NC;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_w1);
    tma_prefetch_desc(&tmap_w2);
    tma_prefetch_desc(&tmap_y);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&x_full[i], 1);
      mbar_init(&x_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_ready[i], kEpiThreads);
    }
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&w1_full[i], 1);
      mbar_init(&w1_empty[i], 1);
      mbar_init(&w2_full[i], 1);
      mbar_init(&w2_empty[i], 1);
    }
    mbar_init(o_full, 1);
    mbar_init(o_empty, kEpiThreads / 32);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  pdl_wait();                 // the producer of x / its row statistics has completed (see common.cuh)
  pdl_launch_dependents();
  for (int i = threadIdx.x; i < G::H; i += kThreads) {
    sB1[i] = p.b1[i];
    sCs[i] = p.colsum[i];
  }
  for (int i = threadIdx.x; i < C; i += kThreads) sB2[i] = p.b2 ? p.b2[i] : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer: x tiles and W1' chunks
    if (lane == 0) {
      int64_t g = 0;
      for (int64_t it = 0; it < my_tiles; ++it) {
        const int64_t tile = blockIdx.x + it * gridDim.x;
        const int xb = (int)(it & 1);
        mbar_wait(&x_empty[xb], ((uint32_t)(it >> 1) & 1) ^ 1);            // the store of tile it-2 has read this buffer
        mbar_arrive_expect_tx(&x_full[xb], G::kXTileBytes);
        for (int c = 0; c < G::kXChunks; ++c)
          tma_load_2d(sX + xb * G::kXTileBytes + c * G::kXChunkBytes, &tmap_x, &x_full[xb], c * kChunkCols, (int)(tile * BM));
        for (int c = 0; c < NC; ++c, ++g) {
          const int s = (int)(g % kStages);
          mbar_wait(&w1_empty[s], ((uint32_t)(g / kStages) & 1) ^ 1);
          mbar_arrive_expect_tx(&w1_full[s], G::kW1Bytes);
          for (int k = 0; k < G::kXChunks; ++k)
            tma_load_2d(sW1 + s * G::kW1Bytes + k * G::kW1ChunkBytes, &tmap_w1, &w1_full[s], k * kChunkCols, c * HC);
        }
      }
    }
  } else if (warp == 3) {
    // ------------------------------------------------------------ TMA producer: W2 chunks
    if (lane == 0) {
      for (int64_t g = 0; g < my_chunks; ++g) {
        const int s = (int)(g % kStages), c = (int)(g % NC);
        mbar_wait(&w2_empty[s], ((uint32_t)(g / kStages) & 1) ^ 1);
        mbar_arrive_expect_tx(&w2_full[s], G::kW2Bytes);
        for (int k = 0; k < G::kW2Chunks; ++k)
          tma_load_2d(sW2 + s * G::kW2Bytes + k * G::kW2ChunkBytes, &tmap_w2, &w2_full[s], c * HC + k * kChunkCols, 0);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0 && my_chunks > 0) {
      constexpr uint32_t idesc_fc1 = make_idesc_bf16(BM, HC, 0, 0);   // A = x (K-major), B = W1' chunk (K-major)
      constexpr uint32_t idesc_fc2 = make_idesc_bf16(BM, C, 0, 0);    // A = P (TMEM),    B = W2 chunk (K-major)
      const uint64_t dsc_x = make_smem_desc(smem_u32(sX), 16, 512, SWZ_64B);
      const uint64_t dsc_w1 = make_smem_desc(smem_u32(sW1), 16, 512, SWZ_64B);
      const uint64_t dsc_w2 = make_smem_desc(smem_u32(sW2), 16, 512, SWZ_64B);
      const uint32_t tS = tmem_base + G::kColS, tP = tmem_base + G::kColP, tO = tmem_base + G::kColO;
      auto issue_fc1 = [&](int64_t g) {
        const int64_t it = g / NC;
        const int xb = (int)(it & 1), s = (int)(g % kStages), b = (int)(g & 1);
        if (g % NC == 0) mbar_wait(&x_full[xb], (uint32_t)(it >> 1) & 1);
        mbar_wait(&w1_full[s], (uint32_t)(g / kStages) & 1);
        tc_fence_after();
        const uint64_t dx = desc_advance(dsc_x, xb * G::kXTileBytes), dw = desc_advance(dsc_w1, s * G::kW1Bytes);
#pragma unroll
        for (int k = 0; k < C / 16; ++k)
          umma_ss(tS + b * HC, desc_advance(dx, (k >> 1) * G::kXChunkBytes + (k & 1) * 32),
                  desc_advance(dw, (k >> 1) * G::kW1ChunkBytes + (k & 1) * 32), idesc_fc1, k != 0);
        umma_commit(&s_full[b]);
        umma_commit(&w1_empty[s]);
      };
      issue_fc1(0);
      if (my_chunks > 1) issue_fc1(1);
      for (int64_t g = 0; g < my_chunks; ++g) {
        const int s = (int)(g % kStages), b = (int)(g & 1), c = (int)(g % NC);
        const int64_t it = g / NC;
        mbar_wait(&w2_full[s], (uint32_t)(g / kStages) & 1);
        mbar_wait(&p_ready[b], (uint32_t)(g >> 1) & 1);              // the epilogue warps wrote P(g)
        if (c == 0 && it > 0) mbar_wait(o_empty, (uint32_t)(it - 1) & 1);   // O of the previous tile has been read
        tc_fence_after();
        const uint64_t dw = desc_advance(dsc_w2, s * G::kW2Bytes);
#pragma unroll
        for (int k = 0; k < HC / 16; ++k)
          umma_ts(tO, tP + b * (HC / 2) + k * 8, desc_advance(dw, (k >> 1) * G::kW2ChunkBytes + (k & 1) * 32), idesc_fc2,
                  (c > 0 || k != 0));
        umma_commit(&w2_empty[s]);
        if (c == NC - 1) umma_commit(o_full);
        if (g + 2 < my_chunks) issue_fc1(g + 2);     // in order after fc2(g): S / P of buffer b are free by then
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ------------------------------------------------------------ GELU / output epilogue warps
    const int e = warp - kEpiWarp0;
    const int q = e & 3, hf = e >> 2;                 // TMEM lane quarter (== warp % 4), column half
    const int et = threadIdx.x - kEpiWarp0 * 32;
    const int row = q * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const uint32_t swz = (uint32_t)((row >> 1) & 3);
    auto row_stats = [&](int64_t tile) {
      const int64_t m = min(tile * BM + row, p.M - 1);
      float2 acc = __ldg(p.ln_stats + m);
      for (int i = 1; i < p.ln_parts; ++i) {
        const float2 v = __ldg(p.ln_stats + (int64_t)i * p.M + m);
        acc.x += v.x;
        acc.y += v.y;
      }
      return acc;
    };
    float2 st_next = my_tiles > 0 ? row_stats(blockIdx.x) : make_float2(0.f, 0.f);
    int64_t g = 0;
    for (int64_t it = 0; it < my_tiles; ++it) {
      const int64_t tile = blockIdx.x + it * gridDim.x;
      const int xb = (int)(it & 1);
      const float2 st = st_next;
      if (it + 1 < my_tiles) st_next = row_stats(tile + gridDim.x);     // next tile's statistics: a tile of slack
      const float mean = st.x * p.ln_inv_c;
      const float ln_a = rsqrtf(fmaxf(st.y * p.ln_inv_c - mean * mean, 0.f) + p.ln_eps);
      const float ln_b = -mean * ln_a;
      const float2 m2 = make_float2(ln_b, ln_b);
      for (int c = 0; c < NC; ++c, ++g) {
        const int b = (int)(g & 1);
        mbar_wait(&s_full[b], (uint32_t)(g >> 1) & 1);
        tc_fence_after();
        const uint32_t tS = tmem_base + lane_base + G::kColS + b * HC + hf * G::kHalf;
        const float *b1 = sB1 + c * HC + hf * G::kHalf, *cs = sCs + c * HC + hf * G::kHalf;
        uint32_t pk[G::kHalf / 2];
#pragma unroll
        for (int part = 0; part < G::kHalf / 32; ++part) {
          uint32_t r[32];
          tmem_ld32(tS + part * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            const int col = part * 32 + v * 8;
            const float4 c0 = *reinterpret_cast<const float4 *>(cs + col), c1 = *reinterpret_cast<const float4 *>(cs + col + 4);
            const float4 d0 = *reinterpret_cast<const float4 *>(b1 + col), d1 = *reinterpret_cast<const float4 *>(b1 + col + 4);
            const float2 sh[4] = {__ffma2_rn(m2, make_float2(c0.x, c0.y), make_float2(d0.x, d0.y)),
                                  __ffma2_rn(m2, make_float2(c0.z, c0.w), make_float2(d0.z, d0.w)),
                                  __ffma2_rn(m2, make_float2(c1.x, c1.y), make_float2(d1.x, d1.y)),
                                  __ffma2_rn(m2, make_float2(c1.z, c1.w), make_float2(d1.z, d1.w))};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 x = make_float2(fmaf(__uint_as_float(r[v * 8 + 2 * j]), ln_a, sh[j].x),
                                           fmaf(__uint_as_float(r[v * 8 + 2 * j + 1]), ln_a, sh[j].y));
              const float2 y = gelu_tanh2(x);
              pk[part * 16 + v * 4 + j] = pack_bf16x2_alu(y.x, y.y);
            }
          }
        }
        const uint32_t tP = tmem_base + lane_base + G::kColP + b * (HC / 2) + hf * (G::kHalf / 2);
        if constexpr (G::kHalf / 2 == 32) tmem_st32(tP, pk);
        else tmem_st16(tP, pk);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&p_ready[b]);
      }
      // ---- output: O + b2 + x (residual, in place over the x tile) -> bf16, row statistics, TMA store
      mbar_wait(o_full, (uint32_t)it & 1);
      tc_fence_after();
      uint8_t *xt = sX + xb * G::kXTileBytes;
      float2 sum2 = make_float2(0.f, 0.f), sq2 = make_float2(0.f, 0.f);
      const uint32_t tO = tmem_base + lane_base + G::kColO + hf * (C / 2);
#pragma unroll
      for (int bt = 0; bt < 3; ++bt) {
        uint32_t r[G::kOB];
        if constexpr (G::kOB == 32) tmem_ld32(tO + bt * 32, r);
        else tmem_ld16(tO + bt * 16, r);
        tmem_ld_wait();
        if (bt == 2) {                                   // O fully read by this warp -> the issuer may start the next tile
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(o_empty);
        }
#pragma unroll
        for (int v = 0; v < G::kOB / 8; ++v) {
          const int col = hf * (C / 2) + bt * G::kOB + v * 8;
          const float4 d0 = *reinterpret_cast<const float4 *>(sB2 + col), d1 = *reinterpret_cast<const float4 *>(sB2 + col + 4);
          const float bb[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
          uint4 *slot = reinterpret_cast<uint4 *>(xt + (col >> 5) * G::kXChunkBytes + row * 64 +
                                                  ((((uint32_t)(col & 31) >> 3) ^ swz) << 4));
          const uint4 rv = *slot;
          const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
          uint32_t o[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float lo = __uint_as_float(r[v * 8 + 2 * j]) + bb[2 * j] + __uint_as_float(rw[j] << 16);
            const float hi = __uint_as_float(r[v * 8 + 2 * j + 1]) + bb[2 * j + 1] + __uint_as_float(rw[j] & 0xffff0000u);
            const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
            o[j] = *reinterpret_cast<const uint32_t *>(&h);
            const float2 f = make_float2(__uint_as_float(o[j] << 16), __uint_as_float(o[j] & 0xffff0000u));
            sum2 = __fadd2_rn(sum2, f);
            sq2 = __ffma2_rn(f, f, sq2);
          }
          *slot = make_uint4(o[0], o[1], o[2], o[3]);
        }
      }
      sStat[hf * BM + row] = make_float2(sum2.x + sum2.y, sq2.x + sq2.y);
      fence_proxy_async_smem();
      asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
      if (et < BM && p.stats_out != nullptr) {
        const float2 v0 = sStat[et], v1 = sStat[BM + et];
        const int64_t m = tile * BM + et;
        if (m < p.M) p.stats_out[m] = make_float2(v0.x + v1.x, v0.y + v1.y);
      }
      if (et == 0) {
#pragma unroll
        for (int c = 0; c < G::kXChunks; ++c)
          tma_store_2d(&tmap_y, xt + c * G::kXChunkBytes, c * kChunkCols, (int)(tile * BM));
        tma_store_commit();
        tma_store_wait_read<0>();
        mbar_arrive(&x_empty[xb]);                       // the x buffer may be refilled (tile it + 2)
      }
      // (no second barrier: sStat is rewritten a whole tile later, and no warp can run more than two chunks ahead of the
      // slowest one — every chunk's fc2 waits for p_ready arrivals from all 256 threads)
    }
    if (et == 0) tma_store_wait_all<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

template <int C>
static int launch(const void *x, const void *w1f, const void *w2, void *y, const Params &p, cudaStream_t st) {
  using G = Cfg<C>;
  CUtensorMap tx, tw1, tw2, ty;
  auto enc2 = [&](CUtensorMap *m, const void *ptr, uint64_t cols, uint64_t rows, uint32_t box_rows) {
    const uint64_t dims[2] = {cols, rows};
    const uint64_t strides[1] = {cols * 2};
    const uint32_t box[2] = {kChunkCols, box_rows};
    return encode_tmap_bf16(m, ptr, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B);
  };
  int r;
  if ((r = enc2(&tx, x, C, (uint64_t)p.M, BM))) return r;
  if ((r = enc2(&tw1, w1f, C, G::H, G::HC))) return r;
  if ((r = enc2(&tw2, w2, G::H, C, C))) return r;
  if ((r = enc2(&ty, y, C, (uint64_t)p.M, BM))) return r;
  MVIT_SMEM_OPT_IN(mlp_fused_kernel<C>, G::kSmemBytes);
  const int64_t m_tiles = (p.M + BM - 1) / BM;
  const unsigned grid = (unsigned)std::min<int64_t>(m_tiles, num_sms());
  MVIT_CUDA_OK(launch_pdl(mlp_fused_kernel<C>, dim3(grid), dim3(kThreads), G::kSmemBytes, st, tx, tw1, tw2, ty, p));
  return 0;
}

int fault_take() { return tc_fault_take(); }

}  // namespace mlpf

int mlp_fused_fault_take() { return mlpf::fault_take(); }

}  // namespace mvit

/* See include/mvit_b200.h. */
extern "C" int mvit_mlp_fused_supported(int C, int H, int C_out) {
  return (C == 96 || C == 192) && H == 4 * C && C_out == C ? 1 : 0;
}

extern "C" int mvit_mlp_fused_fwd(const void *x, const float *ln_stats, int ln_parts, float ln_eps, const void *w1f,
                                  const float *b1f, const float *colsum1, const void *w2, const float *b2, void *y,
                                  float *stats_out, int64_t M, int C, int H, void *stream) {
  using namespace mvit;
  MVIT_REQUIRE(x && ln_stats && w1f && b1f && colsum1 && w2 && y, "mlp_fused: null pointer");
  MVIT_REQUIRE(mvit_mlp_fused_supported(C, H, C), "mlp_fused: C = %d, H = %d unsupported (C in {96, 192}, H = 4C)", C, H);
  MVIT_REQUIRE(M >= 0 && M < ((int64_t)1 << 31), "mlp_fused: bad M");
  MVIT_REQUIRE(ln_parts >= 1 && ln_parts <= 16, "mlp_fused: ln_parts out of range");
  auto al = [](const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  MVIT_REQUIRE(al(x) && al(w1f) && al(w2) && al(y), "mlp_fused: tensors must be 16-byte aligned");
  MVIT_REQUIRE((reinterpret_cast<uintptr_t>(ln_stats) & 7) == 0 && (reinterpret_cast<uintptr_t>(stats_out) & 7) == 0,
               "mlp_fused: statistics buffers must be 8-byte aligned");
  if (M == 0) return 0;
  mlpf::Params p{b1f, colsum1, b2, reinterpret_cast<const float2 *>(ln_stats), reinterpret_cast<float2 *>(stats_out), M,
                 ln_parts, 1.0f / (float)C, ln_eps};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return C == 96 ? mlpf::launch<96>(x, w1f, w2, y, p, st) : mlpf::launch<192>(x, w1f, w2, y, p, st);
}
