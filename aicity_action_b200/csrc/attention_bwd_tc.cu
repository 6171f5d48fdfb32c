// Backward of the fused pooling attention on sm_100a tensor cores (bf16, head_dim 96).
//
// Reference math (attention.py:267-279, differentiated by autograd in the reference):
//   P = softmax(q kᵀ scale),  O = P v (+ q)      given dO and the forward's log-sum-exp L:
//   P_ij = exp(q_i·k_j scale − L_i),  dP_ij = dO_i·v_j,  Δ_i = dO_i·(O_i − q_i),  dS_ij = P_ij (dP_ij − Δ_i)
//   dQ = scale · dS K (+ dO),   dK = scale · dSᵀ Q,   dV = Pᵀ dO.
// Three launches, none of which writes a [Lq, Lk] matrix to memory:
//   prep        Δ_i and L_i·log2(e) per query row (one warp per row);
//   dQ kernel   CTA = 256 query rows (two 128-row streams sharing every 64-key K/V tile); per tile S = Q Kᵀ and
//               dP = dO Vᵀ (tcgen05.mma SS) land in TMEM, the softmax warps (thread = query row) turn them into
//               dS (bf16, written back over S), and dQ += dS K is a TS MMA with K read MN-major from the SAME smem tile;
//   dK/dV kernel CTA = 128 keys x a range of query rows, transposed problem: Sᵀ = K Qᵀ and dPᵀ = V dOᵀ (thread = key),
//               Pᵀ / dSᵀ (bf16) over Sᵀ / dPᵀ, then dV += Pᵀ dO and dK += dSᵀ Q as TS MMAs (dO / Q MN-major from the
//               tiles already in smem).  Score buffers are double-buffered so the tensor pipe never waits for the
//               softmax warps; the query range is split across CTAs (fp32 vector reductions into dK / dV at the end).
// Both kernels reuse exactly the operand forms of the forward kernel (attention_tc.cu): 64B-swizzled 32-column TMA
// boxes, K-major for the score products and MN-major for the accumulating products.
#include "attention.cuh"
#include "tc_common.cuh"

namespace mvit {
using namespace tc;

namespace attn_bwd {
constexpr int D = 96;
constexpr int kChunkCols = 32, kChunks = 3;
constexpr int kChunk128 = 128 * kChunkCols * 2;   // 8 KB: 128 rows x 64 B (SWIZZLE_64B)
constexpr int kTile128 = kChunks * kChunk128;     // 24 KB
constexpr int kChunk64 = 64 * kChunkCols * 2;     // 4 KB
constexpr int kTile64 = kChunks * kChunk64;       // 12 KB
constexpr float kLog2e = 1.44269504088896340736f;

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// fp32 pair -> bf16x2, round-half-away on the integer pipe (F2FP shares the MUFU pipe on sm_100); finite inputs
__device__ __forceinline__ uint32_t pack_bf16x2_alu(float lo, float hi) {
  return __byte_perm(__float_as_uint(lo) + 0x8000u, __float_as_uint(hi) + 0x8000u, 0x7632);
}
__device__ __forceinline__ void bulk_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ------------------------------------------------------------------------------------------------ prep
// delta[bh, row] = dO·(O − add_q·q), lse2[bh, row] = lse·log2(e); rows in [Lq, Lq_pad) are written as zeros.
__global__ void __launch_bounds__(256) prep_kernel(const bf16 *__restrict__ q, const bf16 *__restrict__ out,
                                                   const bf16 *__restrict__ dout, const float *__restrict__ lse,
                                                   float *__restrict__ delta, float *__restrict__ lse2, int heads, int Lq,
                                                   int Lq_pad, int add_q, int64_t total) {
  // 16 lanes per row (12 of them active: 8 channels = one 16-byte load each), two rows per warp
  const int lane = threadIdx.x & 31, sub = lane & 15;
  const int64_t idx = ((int64_t)blockIdx.x * 8 + (threadIdx.x >> 5)) * 2 + (lane >> 4);
  const bool in_range = idx < total;
  const int bh = in_range ? (int)(idx / Lq_pad) : 0, row = in_range ? (int)(idx % Lq_pad) : 0;
  const bool live = in_range && row < Lq;
  float s = 0.f;
  if (live && sub < 12) {
    const int b = bh / heads, head = bh % heads;
    const int64_t o = (((int64_t)b * Lq + row) * heads + head) * D + sub * 8;
    float ov[8], gv[8];
    Vec16<bf16>::load(out + o, ov);
    Vec16<bf16>::load(dout + o, gv);
    if (add_q) {
      float qv[8];
      Vec16<bf16>::load(q + ((int64_t)bh * Lq + row) * D + sub * 8, qv);
#pragma unroll
      for (int e = 0; e < 8; ++e) ov[e] -= qv[e];
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) s = fmaf(gv[e], ov[e], s);
  }
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (in_range && sub == 0) {
    delta[idx] = live ? s : 0.f;
    lse2[idx] = live ? lse[(int64_t)bh * Lq + row] * kLog2e : 0.f;
  }
}

// ------------------------------------------------------------------------------------------------ dQ kernel
namespace dq {
constexpr int BQ = 128, BKV = 64;
constexpr int kStages = 4;
constexpr int kStageBytes = 2 * kTile64;            // K | V
constexpr int kThreads = 640;                       // 4 control warps + 2 streams x 8 dS warps (two per lane quarter)
constexpr int kSmemBytes = 4 * kTile128 + kStages * kStageBytes + 512 + 1024;
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kColS = 0, kColDQ = 256;         // stream i: S at 128 i, dP at 128 i + 64, dQ at 256 + 96 i
}  // namespace dq

struct DqParams {
  const bf16 *dout;
  bf16 *dq;
  const float *delta, *lse2;
  int heads, Lq, Lq_pad, Lk, add_q;
  float scale, scale_log2;
};

__global__ void __launch_bounds__(dq::kThreads, 1)
attention_bwd_dq_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_do,
                        const __grid_constant__ CUtensorMap tmap_k, const __grid_constant__ CUtensorMap tmap_v,
                        const __grid_constant__ CUtensorMap tmap_dq, DqParams p) {
  using namespace dq;
  extern __shared__ uint8_t smem_raw[];
  // 1 KB alignment by an OFFSET in the shared window: the pointer keeps its address space, so every access below compiles
  // to LDS / STS instead of generic LD / ST
  uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t *sQ = smem;                                  // [2][24 KB]
  uint8_t *sdO = smem + 2 * kTile128;                  // [2][24 KB]
  uint8_t *sKV = smem + 4 * kTile128;                  // [stage][K 12 KB | V 12 KB]
  uint64_t *bars = reinterpret_cast<uint64_t *>(sKV + kStages * kStageBytes);
  uint64_t *q_full = bars;                  // 1
  uint64_t *kv_full = bars + 1;             // kStages
  uint64_t *kv_empty = kv_full + kStages;   // kStages
  uint64_t *sdp_full = kv_empty + kStages;  // [stream]
  uint64_t *ds_ready = sdp_full + 2;        // [stream]
  uint64_t *dq_done = ds_ready + 2;         // [stream]
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(dq_done + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.y;
  const int b = bh / p.heads, head = bh % p.heads;
  const int q0 = blockIdx.x * (2 * BQ);
  const bool two = q0 + BQ < p.Lq;
  const int nkv = (p.Lk + BKV - 1) / BKV;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_do);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], two ? 2 : 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&sdp_full[i], 1);
      mbar_init(&ds_ready[i], 256);
      mbar_init(&dq_done[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    if (warp == 0 && lane == 0) {
      // -------------------------------------------------------------- TMA producer
      const int ntile = two ? 2 : 1;
      mbar_arrive_expect_tx(q_full, ntile * 2 * kTile128);
      for (int i = 0; i < ntile; ++i)
        for (int c = 0; c < kChunks; ++c) {
          tma_load_3d(sQ + i * kTile128 + c * kChunk128, &tmap_q, q_full, c * kChunkCols, q0 + i * BQ, bh);
          tma_load_4d(sdO + i * kTile128 + c * kChunk128, &tmap_do, q_full, c * kChunkCols, head, q0 + i * BQ, b);
        }
      for (int j = 0; j < nkv; ++j) {
        const int s = j % kStages;
        mbar_wait(&kv_empty[s], ((j / kStages) & 1) ^ 1);
        uint8_t *kdst = sKV + s * kStageBytes, *vdst = kdst + kTile64;
        mbar_arrive_expect_tx(&kv_full[s], kStageBytes);
        for (int c = 0; c < kChunks; ++c) {
          tma_load_3d(kdst + c * kChunk64, &tmap_k, &kv_full[s], c * kChunkCols, j * BKV, bh);
          tma_load_3d(vdst + c * kChunk64, &tmap_v, &kv_full[s], c * kChunkCols, j * BKV, bh);
        }
      }
    } else if ((warp == 1 || warp == 3) && lane == 0) {
      // -------------------------------------------------------------- MMA issuer of stream i
      const int i = warp == 1 ? 0 : 1;
      if (i == 0 || two) {
        constexpr uint32_t idesc_s = make_idesc_bf16(BQ, BKV, 0, 0);   // A (K-major) x B (K-major)
        constexpr uint32_t idesc_dq = make_idesc_bf16(BQ, D, 0, 1);    // A = dS (TMEM) x B = K (MN-major)
        const uint32_t sq = smem_u32(sQ) + i * kTile128, sdo = smem_u32(sdO) + i * kTile128, skv = smem_u32(sKV);
        const uint32_t tS = tmem_base + kColS + i * 128, tdP = tS + 64, tdQ = tmem_base + kColDQ + i * D;
        // descriptors are built once; every MMA only advances the start-address field (one add on the issue path)
        const uint64_t dsc_q = make_smem_desc(sq, 16, 512, SWZ_64B), dsc_do = make_smem_desc(sdo, 16, 512, SWZ_64B);
        const uint64_t dsc_k0 = make_smem_desc(skv, 16, 512, SWZ_64B);                 // K-major view of stage 0
        const uint64_t dsc_kmn0 = make_smem_desc(skv, kChunk64, 512, SWZ_64B);         // MN-major view of stage 0
        auto issue_sdp = [&](int s) {
          const uint64_t dk = desc_advance(dsc_k0, s * kStageBytes), dv = desc_advance(dk, kTile64);
#pragma unroll
          for (int k = 0; k < D / 16; ++k)
            umma_ss(tS, desc_advance(dsc_q, (k >> 1) * kChunk128 + (k & 1) * 32),
                    desc_advance(dk, (k >> 1) * kChunk64 + (k & 1) * 32), idesc_s, k != 0);
#pragma unroll
          for (int k = 0; k < D / 16; ++k)
            umma_ss(tdP, desc_advance(dsc_do, (k >> 1) * kChunk128 + (k & 1) * 32),
                    desc_advance(dv, (k >> 1) * kChunk64 + (k & 1) * 32), idesc_s, k != 0);
          umma_commit(&sdp_full[i]);
        };
        mbar_wait(q_full, 0);
        mbar_wait(&kv_full[0], 0);
        tc_fence_after();
        issue_sdp(0);
        for (int j = 0; j < nkv; ++j) {
          const int s = j % kStages;
          mbar_wait(&ds_ready[i], j & 1);              // dS(j) is in TMEM (over S)
          tc_fence_after();
          const uint64_t dkmn = desc_advance(dsc_kmn0, s * kStageBytes);
#pragma unroll
          for (int k = 0; k < BKV / 16; ++k)           // K tile read MN-major: LBO = chunk stride, SBO = 8 rows x 64 B
            umma_ts(tdQ, tS + k * 8, desc_advance(dkmn, k * 16 * 64), idesc_dq, (j > 0 || k != 0));
          umma_commit(&kv_empty[s]);
          if (j + 1 < nkv) {
            const int s2 = (j + 1) % kStages;
            mbar_wait(&kv_full[s2], ((j + 1) / kStages) & 1);
            tc_fence_after();
            issue_sdp(s2);                             // in order after dQ(j): may overwrite S / dS
          }
        }
        umma_commit(&dq_done[i]);
      }
    }
  } else {
    // ---------------------------------------------------------------- dS warps: thread = query row; the two warps of
    // a lane quarter split the 64 key columns of a tile (backward needs no cross-column reduction)
    const int i = (warp - 4) >> 3;
    const int quarter = warp & 3, colhalf = ((warp - 4) >> 2) & 1;
    if (i == 0 || two) {
      const int row = q0 + i * BQ + quarter * 32 + lane;
      const bool live = row < p.Lq;
      const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
      const uint32_t tS = tmem_base + lane_base + kColS + i * 128, tdP = tS + 64;
      const uint32_t tdQ = tmem_base + lane_base + kColDQ + i * D;
      const float lse2 = live ? p.lse2[(int64_t)bh * p.Lq_pad + row] : 0.f;
      const float delta = live ? p.delta[(int64_t)bh * p.Lq_pad + row] : 0.f;
      const float2 c2 = make_float2(p.scale_log2, p.scale_log2);
      const float2 nl2 = make_float2(-lse2, -lse2), nd2 = make_float2(-delta, -delta);
      for (int j = 0; j < nkv; ++j) {
        mbar_wait(&sdp_full[i], j & 1);
        tc_fence_after();
        uint32_t pk[16];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t s[16], dp[16];
          tmem_ld16(tS + colhalf * 32 + half * 16, s);
          tmem_ld16(tdP + colhalf * 32 + half * 16, dp);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float2 x = __ffma2_rn(make_float2(__uint_as_float(s[2 * e]), __uint_as_float(s[2 * e + 1])), c2, nl2);
            const float2 pe = make_float2(ex2_approx(x.x), ex2_approx(x.y));
            const float2 t = __fadd2_rn(make_float2(__uint_as_float(dp[2 * e]), __uint_as_float(dp[2 * e + 1])), nd2);
            const float2 ds = __fmul2_rn(pe, t);
            pk[half * 8 + e] = pack_bf16x2_alu(ds.x, ds.y);
          }
        }
        // dS (bf16) goes over the first 32 columns of S: the partner warp must have read them first
        asm volatile("bar.sync %0, 64;" ::"r"(1 + i * 4 + quarter) : "memory");
        tmem_st16(tS + colhalf * 16, pk);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&ds_ready[i]);
      }
      // ---- epilogue: dQ = scale * acc (+ dO); colhalf 0 handles channels 0..47, colhalf 1 channels 48..95.  The tile is
      // assembled over the (no longer needed) Q tile in shared memory, in the same 64B-swizzled chunk layout, and leaves
      // as three TMA stores (coalesced; rows past Lq are clipped by the tensor map).
      mbar_wait(&dq_done[i], 0);
      tc_fence_after();
      const int rl = quarter * 32 + lane;               // row inside the 128-row tile
      if (p.add_q) mbar_wait(q_full, 0);                // acquire the TMA-written dO tile for the generic-proxy reads
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        uint32_t o[16];
        tmem_ld16(tdQ + colhalf * 48 + c * 16, o);
        tmem_ld_wait();
#pragma unroll
        for (int v4 = 0; v4 < 2; ++v4) {
          uint32_t w[4];
          const int ch = colhalf * 48 + c * 16 + v4 * 8;
          const uint32_t off = (ch >> 5) * kChunk128 + rl * 64 + ((((ch & 31) >> 3) ^ ((rl >> 1) & 3)) << 4);
          uint4 gv = make_uint4(0, 0, 0, 0);
          if (p.add_q) gv = *reinterpret_cast<const uint4 *>(sdO + i * kTile128 + off);   // dO, still resident
          const uint32_t gw[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float lo = fmaf(__uint_as_float(o[v4 * 8 + 2 * e]), p.scale, __uint_as_float(gw[e] << 16));
            const float hi = fmaf(__uint_as_float(o[v4 * 8 + 2 * e + 1]), p.scale, __uint_as_float(gw[e] & 0xffff0000u));
            __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
            w[e] = *reinterpret_cast<uint32_t *>(&h);
          }
          *reinterpret_cast<uint4 *>(sQ + i * kTile128 + off) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
      fence_proxy_async_smem();
      asm volatile("bar.sync %0, 256;" ::"r"(9 + i) : "memory");    // the 256 threads of this stream
      if (quarter == 0 && colhalf == 0 && lane == 0) {
#pragma unroll
        for (int c = 0; c < kChunks; ++c)
          tma_store_3d(&tmap_dq, sQ + i * kTile128 + c * kChunk128, c * kChunkCols, q0 + i * BQ, bh);
        tma_store_commit();
        tma_store_wait_all<0>();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, kTmemCols);
}

// ------------------------------------------------------------------------------------------------ dK / dV kernel
namespace dkv {
constexpr int BK = 128, BQ = 64;
constexpr int kStages = 5;
constexpr int kStageBytes = 2 * kTile64;            // Q | dO
constexpr int kVecBytes = 512;                      // lse2[64] | delta[64] fp32
constexpr int kThreads = 384;                       // 4 control warps + 8 P/dS warps (two per TMEM lane quarter)
constexpr int kSmemBytes = 2 * kTile128 + kStages * (kStageBytes + kVecBytes) + 512 + 1024;
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kColDK = 256, kColDV = 352;      // buffer b: S^T at 128 b, dP^T at 128 b + 64
}  // namespace dkv

struct DkvParams {
  float *dk, *dv;
  const float *delta, *lse2;
  int heads, Lq, Lq_pad, Lk, tiles_per_split;
  float scale, scale_log2;
};

__global__ void __launch_bounds__(dkv::kThreads, 1)
attention_bwd_dkv_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_do,
                         const __grid_constant__ CUtensorMap tmap_k, const __grid_constant__ CUtensorMap tmap_v,
                         DkvParams p) {
  using namespace dkv;
  extern __shared__ uint8_t smem_raw[];
  // 1 KB alignment by an OFFSET in the shared window: the pointer keeps its address space, so every access below compiles
  // to LDS / STS instead of generic LD / ST
  uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t *sK = smem, *sV = smem + kTile128;
  uint8_t *sRing = smem + 2 * kTile128;                // [stage][Q 12 KB | dO 12 KB]
  uint8_t *sVec = sRing + kStages * kStageBytes;       // [stage][lse2 64 | delta 64]
  uint64_t *bars = reinterpret_cast<uint64_t *>(sVec + kStages * kVecBytes);
  uint64_t *kv_full = bars;                  // 1
  uint64_t *qd_full = bars + 1;              // kStages
  uint64_t *qd_empty = qd_full + kStages;    // kStages
  uint64_t *sdp_full = qd_empty + kStages;   // [buffer]
  uint64_t *pds_ready = sdp_full + 2;        // [buffer]
  uint64_t *acc_done = pds_ready + 2;        // 1
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.y;
  const int b = bh / p.heads, head = bh % p.heads;
  const int key0 = blockIdx.x * BK;
  const int nq_total = (p.Lq + BQ - 1) / BQ;
  const int t0 = blockIdx.z * p.tiles_per_split;
  const int n = min(p.tiles_per_split, nq_total - t0);   // >= 1 by construction of the grid

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_do);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(kv_full, 1);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&qd_full[i], 1);
      mbar_init(&qd_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&sdp_full[i], 1);
      mbar_init(&pds_ready[i], 256);
    }
    mbar_init(acc_done, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 && lane == 0) {
    // ---------------------------------------------------------------- TMA producer
    mbar_arrive_expect_tx(kv_full, 2 * kTile128);
    for (int c = 0; c < kChunks; ++c) {
      tma_load_3d(sK + c * kChunk128, &tmap_k, kv_full, c * kChunkCols, key0, bh);
      tma_load_3d(sV + c * kChunk128, &tmap_v, kv_full, c * kChunkCols, key0, bh);
    }
    for (int j = 0; j < n; ++j) {
      const int s = j % kStages;
      mbar_wait(&qd_empty[s], ((j / kStages) & 1) ^ 1);
      const int row0 = (t0 + j) * BQ;
      uint8_t *qdst = sRing + s * kStageBytes, *ddst = qdst + kTile64;
      mbar_arrive_expect_tx(&qd_full[s], kStageBytes + kVecBytes);
      for (int c = 0; c < kChunks; ++c) {
        tma_load_3d(qdst + c * kChunk64, &tmap_q, &qd_full[s], c * kChunkCols, row0, bh);
        tma_load_4d(ddst + c * kChunk64, &tmap_do, &qd_full[s], c * kChunkCols, head, row0, b);
      }
      bulk_load_1d(sVec + s * kVecBytes, p.lse2 + (int64_t)bh * p.Lq_pad + row0, 256, &qd_full[s]);
      bulk_load_1d(sVec + s * kVecBytes + 256, p.delta + (int64_t)bh * p.Lq_pad + row0, 256, &qd_full[s]);
    }
  } else if (warp == 1 && lane == 0) {
    // ---------------------------------------------------------------- MMA issuer
    constexpr uint32_t idesc_s = make_idesc_bf16(BK, BQ, 0, 0);     // A = K or V (K-major) x B = Q or dO (K-major)
    constexpr uint32_t idesc_acc = make_idesc_bf16(BK, D, 0, 1);    // A = P^T / dS^T (TMEM) x B = dO / Q (MN-major)
    const uint32_t sk = smem_u32(sK), sv = smem_u32(sV), sring = smem_u32(sRing);
    const uint32_t tdK = tmem_base + kColDK, tdV = tmem_base + kColDV;
    // descriptors are built once; every MMA only advances the start-address field (one add on the issue path)
    const uint64_t dsc_k = make_smem_desc(sk, 16, 512, SWZ_64B), dsc_v = make_smem_desc(sv, 16, 512, SWZ_64B);
    const uint64_t dsc_r0 = make_smem_desc(sring, 16, 512, SWZ_64B);                   // K-major view of ring stage 0
    const uint64_t dsc_rmn0 = make_smem_desc(sring, kChunk64, 512, SWZ_64B);           // MN-major view of ring stage 0
    auto issue_sdp = [&](int s, int bf) {
      const uint64_t dq = desc_advance(dsc_r0, s * kStageBytes), dd = desc_advance(dq, kTile64);
      const uint32_t tS = tmem_base + bf * 128, tdP = tS + 64;
#pragma unroll
      for (int k = 0; k < D / 16; ++k)
        umma_ss(tS, desc_advance(dsc_k, (k >> 1) * kChunk128 + (k & 1) * 32),
                desc_advance(dq, (k >> 1) * kChunk64 + (k & 1) * 32), idesc_s, k != 0);
#pragma unroll
      for (int k = 0; k < D / 16; ++k)
        umma_ss(tdP, desc_advance(dsc_v, (k >> 1) * kChunk128 + (k & 1) * 32),
                desc_advance(dd, (k >> 1) * kChunk64 + (k & 1) * 32), idesc_s, k != 0);
      umma_commit(&sdp_full[bf]);
    };
    mbar_wait(kv_full, 0);
    for (int jj = 0; jj < 2 && jj < n; ++jj) {
      mbar_wait(&qd_full[jj % kStages], 0);
      tc_fence_after();
      issue_sdp(jj % kStages, jj);
    }
    for (int j = 0; j < n; ++j) {
      const int s = j % kStages, bf = j & 1;
      mbar_wait(&pds_ready[bf], (j >> 1) & 1);
      tc_fence_after();
      const uint64_t dqmn = desc_advance(dsc_rmn0, s * kStageBytes), ddmn = desc_advance(dqmn, kTile64);
      const uint32_t tS = tmem_base + bf * 128, tdP = tS + 64;
#pragma unroll
      for (int k = 0; k < BQ / 16; ++k)
        umma_ts(tdV, tS + k * 8, desc_advance(ddmn, k * 16 * 64), idesc_acc, (j > 0 || k != 0));
#pragma unroll
      for (int k = 0; k < BQ / 16; ++k)
        umma_ts(tdK, tdP + k * 8, desc_advance(dqmn, k * 16 * 64), idesc_acc, (j > 0 || k != 0));
      umma_commit(&qd_empty[s]);
      if (j + 2 < n) {
        const int s2 = (j + 2) % kStages;
        mbar_wait(&qd_full[s2], ((j + 2) / kStages) & 1);
        tc_fence_after();
        issue_sdp(s2, bf);
      }
    }
    umma_commit(acc_done);
  } else if (warp >= 4) {
    // ---------------------------------------------------------------- P^T / dS^T warps: thread = key, and the two
    // warps of a lane quarter split the 64 query columns of a tile (no cross-column reduction is needed in backward)
    const int quarter = warp & 3, colhalf = (warp - 4) >> 2;
    const int key = key0 + quarter * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    const float2 c2 = make_float2(p.scale_log2, p.scale_log2);
    for (int j = 0; j < n; ++j) {
      const int s = j % kStages, bf = j & 1;
      mbar_wait(&qd_full[s], (j / kStages) & 1);       // visibility of the bulk-copied lse2 / delta vectors
      mbar_wait(&sdp_full[bf], (j >> 1) & 1);
      tc_fence_after();
      const uint32_t tS = tmem_base + lane_base + bf * 128, tdP = tS + 64;
      const float4 *vl = reinterpret_cast<const float4 *>(sVec + s * kVecBytes) + colhalf * 8;
      const float4 *vd = vl + 16;
      uint32_t pkp[16], pkd[16];
      {
        uint32_t sc[32], dp[32];
        tmem_ld32(tS + colhalf * 32, sc);
        tmem_ld32(tdP + colhalf * 32, dp);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const float4 l4 = vl[g], d4 = vd[g];
          const float lv[4] = {l4.x, l4.y, l4.z, l4.w}, dv[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {
            const int e = g * 4 + h2 * 2;
            const float2 x = __ffma2_rn(make_float2(__uint_as_float(sc[e]), __uint_as_float(sc[e + 1])), c2,
                                        make_float2(-lv[h2 * 2], -lv[h2 * 2 + 1]));
            const float2 pe = make_float2(ex2_approx(x.x), ex2_approx(x.y));
            const float2 t = __fadd2_rn(make_float2(__uint_as_float(dp[e]), __uint_as_float(dp[e + 1])),
                                        make_float2(-dv[h2 * 2], -dv[h2 * 2 + 1]));
            const float2 ds = __fmul2_rn(pe, t);
            pkp[g * 2 + h2] = pack_bf16x2_alu(pe.x, pe.y);
            pkd[g * 2 + h2] = pack_bf16x2_alu(ds.x, ds.y);
          }
        }
      }
      // P^T / dS^T (bf16) go over the first 32 columns of S^T / dP^T: the partner warp must have read them first
      asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");
      tmem_st16(tS + colhalf * 16, pkp);
      tmem_st16(tdP + colhalf * 16, pkd);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&pds_ready[bf]);
    }
    // ---- epilogue: fp32 vector reductions; colhalf 0 -> dK (scaled), colhalf 1 -> dV
    mbar_wait(acc_done, 0);
    tc_fence_after();
    const bool live = key < p.Lk;
    const uint32_t tA = tmem_base + lane_base + (colhalf == 0 ? kColDK : kColDV);
    float *dst = (colhalf == 0 ? p.dk : p.dv) + ((int64_t)bh * p.Lk + key) * D;
    const float mul = colhalf == 0 ? p.scale : 1.0f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      uint32_t o[32];
      tmem_ld32(tA + c * 32, o);
      tmem_ld_wait();
      if (live) {
#pragma unroll
        for (int v4 = 0; v4 < 8; ++v4) {
          const float4 val = make_float4(__uint_as_float(o[v4 * 4]) * mul, __uint_as_float(o[v4 * 4 + 1]) * mul,
                                         __uint_as_float(o[v4 * 4 + 2]) * mul, __uint_as_float(o[v4 * 4 + 3]) * mul);
          atomicAdd(reinterpret_cast<float4 *>(dst + c * 32 + v4 * 4), val);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, kTmemCols);
}

}  // namespace attn_bwd

bool attention_bwd_tc_supported(const AttnBwdArgs &a, const char **why) {
  auto al = [](const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (!al(a.q) || !al(a.k) || !al(a.v) || !al(a.out) || !al(a.dout) || !al(a.dq) || !al(a.dk) || !al(a.dv) ||
      !al(a.workspace)) {
    *why = "pointers must be 16-byte aligned";
    return false;
  }
  if (!a.workspace) { *why = "needs the workspace (mvit_attention_bwd_workspace_floats)"; return false; }
  if ((int64_t)a.B * a.heads >= 65536) { *why = "B*heads too large"; return false; }
  return true;
}

size_t attention_bwd_workspace_floats(int B, int heads, int Lq) {
  const size_t lq_pad = ((size_t)Lq + 63) / 64 * 64;
  return 2 * (size_t)B * heads * lq_pad;
}

int attention_bwd_tc(const AttnBwdArgs &a, cudaStream_t st) {
  using namespace attn_bwd;
  const int BH = a.B * a.heads;
  const int Lq_pad = (a.Lq + 63) / 64 * 64;
  float *delta = a.workspace, *lse2 = a.workspace + (size_t)BH * Lq_pad;
  {
    const int64_t total = (int64_t)BH * Lq_pad;
    prep_kernel<<<(unsigned)((total + 15) / 16), 256, 0, st>>>(static_cast<const bf16 *>(a.q), static_cast<const bf16 *>(a.out),
                                                            static_cast<const bf16 *>(a.dout), a.lse, delta, lse2, a.heads,
                                                            a.Lq, Lq_pad, a.add_q, total);
    MVIT_LAUNCH_OK("attention_bwd(prep)");
  }
  auto enc3 = [&](CUtensorMap *m, const void *ptr, int L, int box_rows) {
    const uint64_t dims[3] = {(uint64_t)D, (uint64_t)L, (uint64_t)BH};
    const uint64_t strides[2] = {(uint64_t)D * 2, (uint64_t)L * D * 2};
    const uint32_t box[3] = {kChunkCols, (uint32_t)box_rows, 1};
    return encode_tmap_bf16(m, ptr, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B);
  };
  auto enc_do = [&](CUtensorMap *m, int box_rows) {   // dout [B, Lq, heads, 96]
    const uint64_t dims[4] = {(uint64_t)D, (uint64_t)a.heads, (uint64_t)a.Lq, (uint64_t)a.B};
    const uint64_t strides[3] = {(uint64_t)D * 2, (uint64_t)a.heads * D * 2, (uint64_t)a.Lq * a.heads * D * 2};
    const uint32_t box[4] = {kChunkCols, 1, (uint32_t)box_rows, 1};
    return encode_tmap_bf16(m, a.dout, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B);
  };
  MVIT_SMEM_OPT_IN(attention_bwd_dq_kernel, dq::kSmemBytes);
  MVIT_SMEM_OPT_IN(attention_bwd_dkv_kernel, dkv::kSmemBytes);
  const float scale_log2 = a.scale * kLog2e;
  int r;
  {
    CUtensorMap tq, tdo, tk, tv;
    if ((r = enc3(&tq, a.q, a.Lq, dq::BQ))) return r;
    if ((r = enc_do(&tdo, dq::BQ))) return r;
    if ((r = enc3(&tk, a.k, a.Lk, dq::BKV))) return r;
    if ((r = enc3(&tv, a.v, a.Lk, dq::BKV))) return r;
    CUtensorMap tdq;
    if ((r = enc3(&tdq, a.dq, a.Lq, dq::BQ))) return r;
    DqParams p{static_cast<const bf16 *>(a.dout), static_cast<bf16 *>(a.dq), delta, lse2, a.heads, a.Lq, Lq_pad, a.Lk,
               a.add_q, a.scale, scale_log2};
    dim3 grid((unsigned)((a.Lq + 2 * dq::BQ - 1) / (2 * dq::BQ)), (unsigned)BH);
    attention_bwd_dq_kernel<<<grid, dq::kThreads, dq::kSmemBytes, st>>>(tq, tdo, tk, tv, tdq, p);
    MVIT_LAUNCH_OK("attention_bwd(dq)");
  }
  {
    CUtensorMap tq, tdo, tk, tv;
    if ((r = enc3(&tq, a.q, a.Lq, dkv::BQ))) return r;
    if ((r = enc_do(&tdo, dkv::BQ))) return r;
    if ((r = enc3(&tk, a.k, a.Lk, dkv::BK))) return r;
    if ((r = enc3(&tv, a.v, a.Lk, dkv::BK))) return r;
    const int ktiles = (a.Lk + dkv::BK - 1) / dkv::BK, nq = (a.Lq + dkv::BQ - 1) / dkv::BQ;
    int splits = std::max(1, std::min(nq, (2 * num_sms() + ktiles * BH - 1) / (ktiles * BH)));
    const int tps = (nq + splits - 1) / splits;
    splits = (nq + tps - 1) / tps;
    DkvParams p{a.dk, a.dv, delta, lse2, a.heads, a.Lq, Lq_pad, a.Lk, tps, a.scale, scale_log2};
    dim3 grid((unsigned)ktiles, (unsigned)BH, (unsigned)splits);
    attention_bwd_dkv_kernel<<<grid, dkv::kThreads, dkv::kSmemBytes, st>>>(tq, tdo, tk, tv, p);
    MVIT_LAUNCH_OK("attention_bwd(dkv)");
  }
  return 0;
}

int attention_bwd_tc_fault_take() { return tc_fault_take(); }

}  // namespace mvit
