// nn.Linear weight gradient on sm_100a tensor cores:  dW[N, K] += dy[M, N]ᵀ · x[M, K]   (bf16 operands, fp32 result).
//
// The reduction runs over the token axis M, which is the ROW index of both row-major operands, so both are consumed
// MN-major: a TMA box [64 tokens x 64 features] (128-byte rows, SWIZZLE_128B) is exactly one UMMA "MN-major" atom
// column, and no transposed copy of an activation is ever made.  CTA = one 128 x BQ tile of dW and one slice of the token
// axis (split-M: tiles x splits ~ 2 waves of SMs); warp 0 streams [64-token] chunks of dy / x through a TMA ring, one
// lane of warp 1 issues tcgen05.mma (M128 x N{128,192} x K16, four per chunk) into a TMEM accumulator, warps 4-7 add the
// finished tile into dW with vector fp32 reductions.  CTAs of one token slice are adjacent in the grid so the
// activation chunks they share are served by L2.
// Bias gradient for free: every stage carries one extra, constant x-box whose feature 0 is 1.0; the CTAs of the first
// feature tile run their MMAs 16 columns wider, so accumulator column BQ is sum_m dy[m, n] = db[n] - dy is not re-read.
#include "linear.cuh"
#include "tc_common.cuh"

namespace mvit {
using namespace tc;

namespace wgrad {
constexpr int BP = 128, BMT = 64, kBox = 64;      // dW rows per CTA, tokens per chunk, features per TMA box
constexpr int kBoxBytes = BMT * kBox * 2;         // 8 KB
constexpr int kThreads = 256;

template <int BQ> struct Cfg {
  static constexpr int kABoxes = BP / kBox, kBBoxes = BQ / kBox;
  static constexpr int kStageBytes = (kABoxes + kBBoxes + 1) * kBoxBytes;   // + the constant ones box
  static constexpr int kStages = BQ == 192 ? 4 : 5;
  static constexpr int kSmemBytes = kStages * kStageBytes + 256 + 1024;
  static constexpr uint32_t kTmemCols = 256;
};

template <int BQ>
__global__ void __launch_bounds__(kThreads, 1)
linear_wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmap_dy, const __grid_constant__ CUtensorMap tmap_x,
                       float *__restrict__ dw, float *__restrict__ db, int P, int Q, int64_t M, int64_t m_per_split,
                       int q_tiles) {
  using C = Cfg<BQ>;
  extern __shared__ uint8_t smem_raw[];
  // 1 KB alignment by an OFFSET in the shared window: the pointer keeps its address space, so every access below compiles
  // to LDS / STS instead of generic LD / ST
  uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + C::kStages * C::kStageBytes);
  uint64_t *full = bars, *empty = bars + C::kStages, *done = empty + C::kStages;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int p0 = (blockIdx.x / q_tiles) * BP, q0 = (blockIdx.x % q_tiles) * BQ;
  const int64_t m_begin = (int64_t)blockIdx.y * m_per_split, m_end = min(M, m_begin + m_per_split);
  const int chunks = (int)((m_end - m_begin + BMT - 1) / BMT);      // >= 1 by construction of the grid

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_dy);
    tma_prefetch_desc(&tmap_x);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < C::kStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, C::kTmemCols);
    tmem_relinquish();
  }
  const bool with_bias = db != nullptr && q0 == 0;     // this CTA also produces db[p0 .. p0+128)
  if (warp == 3 && with_bias) {
    // ones box of every stage: [64 tokens][64 features] bf16, 128B-swizzled, feature 0 = 1.0, the rest 0
    for (int st = 0; st < C::kStages; ++st) {
      uint4 *box = reinterpret_cast<uint4 *>(smem + st * C::kStageBytes + (C::kABoxes + C::kBBoxes) * kBoxBytes);
      for (int i = lane; i < kBoxBytes / 16; i += 32) {
        const int r = i >> 3, c16 = i & 7;             // token row, physical 16-byte slot
        box[i] = make_uint4(c16 == (r & 7) ? 0x00003F80u : 0u, 0u, 0u, 0u);
      }
    }
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < chunks; ++i) {
      const int s = i % C::kStages;
      mbar_wait(&empty[s], ((i / C::kStages) & 1) ^ 1);
      uint8_t *a = smem + s * C::kStageBytes, *b = a + C::kABoxes * kBoxBytes;
      // rows past M and columns past P / Q are zero-filled by TMA and contribute nothing
      const int m0 = (int)(m_begin + (int64_t)i * BMT);
      mbar_arrive_expect_tx(&full[s], (C::kABoxes + C::kBBoxes) * kBoxBytes);
#pragma unroll
      for (int c = 0; c < C::kABoxes; ++c) tma_load_2d(a + c * kBoxBytes, &tmap_dy, &full[s], p0 + c * kBox, m0);
#pragma unroll
      for (int c = 0; c < C::kBBoxes; ++c) tma_load_2d(b + c * kBoxBytes, &tmap_x, &full[s], q0 + c * kBox, m0);
    }
  } else if (warp == 1 && lane == 0) {
    // A and B both MN-major; 16 more columns (the ones box) when this CTA also reduces the bias gradient
    const uint32_t idesc = with_bias ? make_idesc_bf16(BP, BQ + 16, 1, 1) : make_idesc_bf16(BP, BQ, 1, 1);
    const uint64_t dsc0 = make_smem_desc(smem_u32(smem), kBoxBytes, 1024, SWZ_128B);   // stage 0, advanced per MMA
    for (int i = 0; i < chunks; ++i) {
      const int s = i % C::kStages;
      mbar_wait(&full[s], (i / C::kStages) & 1);
      tc_fence_after();
      const uint64_t da = desc_advance(dsc0, s * C::kStageBytes), db = desc_advance(da, C::kABoxes * kBoxBytes);
#pragma unroll
      for (int k = 0; k < BMT / 16; ++k)   // 16 tokens = 16 rows x 128 B; LBO = next 64-feature box, SBO = 8 rows x 128 B
        umma_ss(tmem_base, desc_advance(da, k * 16 * 128), desc_advance(db, k * 16 * 128), idesc, (i > 0 || k != 0));
      umma_commit(&empty[s]);
    }
    umma_commit(done);
  } else if (warp >= 4) {
    const int quarter = warp & 3;
    mbar_wait(done, 0);
    tc_fence_after();
    const int prow = p0 + quarter * 32 + lane;
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    float *dst = dw + (int64_t)prow * Q + q0;
#pragma unroll
    for (int c = 0; c < BQ / 32; ++c) {
      uint32_t o[32];
      tmem_ld32(taddr + c * 32, o);
      tmem_ld_wait();
      if (prow < P) {
#pragma unroll
        for (int v4 = 0; v4 < 8; ++v4) {
          if (q0 + c * 32 + v4 * 4 < Q)      // Q % 4 == 0
            atomicAdd(reinterpret_cast<float4 *>(dst + c * 32 + v4 * 4),
                      make_float4(__uint_as_float(o[v4 * 4]), __uint_as_float(o[v4 * 4 + 1]),
                                  __uint_as_float(o[v4 * 4 + 2]), __uint_as_float(o[v4 * 4 + 3])));
        }
      }
    }
    if (with_bias) {
      uint32_t o[16];
      tmem_ld16(taddr + BQ, o);
      tmem_ld_wait();
      if (prow < P) atomicAdd(&db[prow], __uint_as_float(o[0]));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, C::kTmemCols);
}

}  // namespace wgrad

bool linear_wgrad_tc_supported(const void *dy, const void *x, const float *dw, int64_t M, int N, int K, const char **why) {
  if (N % 8 != 0 || K % 8 != 0) { *why = "N and K must be multiples of 8 (16-byte TMA row pitch)"; return false; }
  if (M >= ((int64_t)1 << 31)) { *why = "M too large"; return false; }
  auto al = [](const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (!al(dy) || !al(x) || !al(dw)) { *why = "pointers must be 16-byte aligned"; return false; }
  return true;
}

template <int BQ>
static int launch_wgrad(const void *dy, const void *x, float *dw, float *db, int64_t M, int N, int K, cudaStream_t st) {
  using C = wgrad::Cfg<BQ>;
  CUtensorMap tdy, tx;
  auto enc2 = [&](CUtensorMap *m, const void *ptr, uint64_t cols, uint64_t rows) {
    const uint64_t dims[2] = {cols, rows};
    const uint64_t strides[1] = {cols * 2};
    const uint32_t box[2] = {wgrad::kBox, wgrad::BMT};
    return encode_tmap_bf16(m, ptr, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
  };
  int r;
  if ((r = enc2(&tdy, dy, (uint64_t)N, (uint64_t)M))) return r;
  if ((r = enc2(&tx, x, (uint64_t)K, (uint64_t)M))) return r;
  MVIT_SMEM_OPT_IN(wgrad::linear_wgrad_tc_kernel<BQ>, C::kSmemBytes);
  const int p_tiles = (N + wgrad::BP - 1) / wgrad::BP, q_tiles = (K + BQ - 1) / BQ, tiles = p_tiles * q_tiles;
  const int64_t chunks = (M + wgrad::BMT - 1) / wgrad::BMT;
  int64_t splits = std::max<int64_t>(1, (2 * num_sms() + tiles - 1) / tiles);
  splits = std::min<int64_t>(splits, std::max<int64_t>(1, chunks / 8));       // at least 8 chunks per CTA
  const int64_t cps = (chunks + splits - 1) / splits;
  splits = (chunks + cps - 1) / cps;
  MVIT_REQUIRE(splits < 65536, "linear_wgrad: too many token splits");
  dim3 grid((unsigned)tiles, (unsigned)splits);
  wgrad::linear_wgrad_tc_kernel<BQ><<<grid, wgrad::kThreads, C::kSmemBytes, st>>>(tdy, tx, dw, db, N, K, M, cps * wgrad::BMT, q_tiles);
  MVIT_LAUNCH_OK("linear_wgrad(tcgen05)");
  return 0;
}

int linear_wgrad_tc(const void *dy, const void *x, float *dw, float *db, int64_t M, int N, int K, cudaStream_t st) {
  return K > 128 ? launch_wgrad<192>(dy, x, dw, db, M, N, K, st) : launch_wgrad<128>(dy, x, dw, db, M, N, K, st);
}

int gemm_wgrad_tc_fault_take() { return tc_fault_take(); }

}  // namespace mvit
