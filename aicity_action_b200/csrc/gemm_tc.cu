// tcgen05 Linear for sm_100a:  y = epi(x·wᵀ + bias)·row_scale + residual      (bf16 in/out, fp32 accumulate)
//
// Replaces the cuBLAS GEMMs + separate bias / GELU / residual-add / DropPath kernels behind
// attention.py:231,281,426,443 and common.py:27-31 with one persistent, warp-specialised kernel:
//   warp 0     TMA producer: x tile [128 x 64] and w tile [BN x 64] (128B-swizzled, K-major) into a
//              kStages-deep shared-memory ring, completion on `full` mbarriers;
//   warp 1     MMA issuer: one elected lane issues tcgen05.mma (M=128, N=BN, K=16) x 4 per stage into a
//              double-buffered fp32 accumulator in TMEM; tcgen05.commit releases the smem slot / signals
//              the epilogue;
//   warp 2     TMEM allocator (alloc at start, dealloc at exit);
//   warps 4-7  epilogue: tcgen05.ld (lane = row) -> +bias -> GELU -> *row_scale -> bf16 -> per-warp smem
//              staging -> coalesced 16-byte row stores with the residual added on the way out.
// The accumulator of tile i+1 is produced while tile i is drained, so the epilogue hides behind the MMAs.
#include "linear.cuh"
#include "tc_common.cuh"

namespace mvit {
using namespace tc;

namespace gemm {
constexpr int BM = 128;
constexpr int BK = 64;                     // 64 bf16 = one 128-byte swizzle atom row
constexpr int UMMA_K = 16;
constexpr int kThreads = 256;
constexpr int kEpiWarp0 = 4;

template <int BN> struct Cfg {
  static constexpr int kStages = BN == 192 ? 4 : 6;
  static constexpr int kABytes = BM * BK * 2;          // 16 KB
  static constexpr int kBBytes = BN * BK * 2;          // 24 KB / 12 KB
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kRowPitch = BN * 2 + 16;        // staging row pitch (bytes), 16B-aligned, bank-skewed
  static constexpr int kStagingBytes = 4 * 32 * kRowPitch;
  static constexpr int kTmemCols = BN == 192 ? 512 : 256;   // 2 accumulators of BN columns, power of two
  static constexpr int kSmemBytes = kStages * kStageBytes + kStagingBytes + BN * 4 /*bias*/ + 256 /*barriers*/ + 1024 /*align*/;
};

struct Params {
  const float *bias, *row_scale;
  const bf16 *residual;
  bf16 *y;
  int64_t M, rows_per_sample, ldy, ldr;
  int N, K, epilogue;
};

template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
linear_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w, Params p) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t *sA = smem;
  uint8_t *sB = smem + C::kStages * C::kABytes;
  uint8_t *sStage = smem + C::kStages * C::kStageBytes;
  float *sBias = reinterpret_cast<float *>(sStage + C::kStagingBytes);
  uint64_t *bars = reinterpret_cast<uint64_t *>(sBias + BN);
  uint64_t *full = bars, *empty = bars + C::kStages, *tfull = bars + 2 * C::kStages, *tempty = tfull + 2;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (p.N + BN - 1) / BN;
  const int64_t m_tiles = (p.M + BM - 1) / BM;
  const int64_t tiles = m_tiles * n_tiles;
  const int k_blocks = (p.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_w);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < C::kStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 4);   // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, C::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
        const int m0 = (int)(t / n_tiles) * BM, n0 = (int)(t % n_tiles) * BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full[stage], C::kStageBytes);
          tma_load_2d(sA + stage * C::kABytes, &tmap_x, &full[stage], kb * BK, m0);
          tma_load_2d(sB + stage * C::kBBytes, &tmap_w, &full[stage], kb * BK, n0);
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
        mbar_wait(&tempty[acc], acc_phase ^ 1);       // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a0 = smem_u32(sA + stage * C::kABytes), b0 = smem_u32(sB + stage * C::kBBytes);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t da = make_smem_desc(a0 + k * UMMA_K * 2, 16, 1024, SWZ_128B);
            const uint64_t db = make_smem_desc(b0 + k * UMMA_K * 2, 16, 1024, SWZ_128B);
            umma_ss(d_tmem, da, db, idesc, (kb | k) != 0);
          }
          umma_commit(&empty[stage]);                 // smem slot free once these MMAs have read it
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull[acc]);                     // accumulator complete
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ------------------------------------------------------------ epilogue (4 warps, TMEM lane quarter = warp % 4)
    const int q = warp & 3;
    uint8_t *stg = sStage + q * 32 * C::kRowPitch;
    constexpr int kLanesPerRow = BN / 8;              // 16-byte vectors per output row
    constexpr int kRowsPerIter = 32 / kLanesPerRow;
    const int et = threadIdx.x - kEpiWarp0 * 32;      // 0..127
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
      const int64_t m0 = (t / n_tiles) * BM;
      const int n0 = (int)(t % n_tiles) * BN;
      // stage this tile's bias slice (all 4 epilogue warps cooperate; named barrier 1)
      asm volatile("bar.sync 1, 128;" ::: "memory");   // previous tile's readers are done with sBias
      for (int i = et; i < BN; i += 128) sBias[i] = (p.bias && n0 + i < p.N) ? p.bias[n0 + i] : 0.f;
      asm volatile("bar.sync 1, 128;" ::: "memory");
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const int64_t row = m0 + q * 32 + lane;
      float rs = 1.f;
      if (p.row_scale && row < p.M) rs = p.row_scale[row / p.rows_per_sample];
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(taddr + c0, r);
        tmem_ld_wait();
        uint32_t packed[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float v0 = __uint_as_float(r[2 * j]) + sBias[c0 + 2 * j];
          float v1 = __uint_as_float(r[2 * j + 1]) + sBias[c0 + 2 * j + 1];
          if (p.epilogue == MVIT_EPI_GELU) { v0 = gelu_erf(v0); v1 = gelu_erf(v1); }
          v0 *= rs; v1 *= rs;
          __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);
          packed[j] = *reinterpret_cast<uint32_t *>(&h);
        }
        uint4 *dst = reinterpret_cast<uint4 *>(stg + lane * C::kRowPitch + c0 * 2);
#pragma unroll
        for (int j = 0; j < 4; ++j) dst[j] = make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
      }
      // accumulator fully read -> hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      // coalesced write-out of this warp's 32 rows (+ residual)
      const int rsub = lane / kLanesPerRow, cv = lane % kLanesPerRow;
      const int n = n0 + cv * 8;
      if (rsub < kRowsPerIter && n < p.N) {
#pragma unroll 4
        for (int r0 = 0; r0 < 32; r0 += kRowsPerIter) {
          const int rr = r0 + rsub;
          const int64_t m = m0 + q * 32 + rr;
          if (m >= p.M) break;
          uint4 v = *reinterpret_cast<const uint4 *>(stg + rr * C::kRowPitch + cv * 16);
          if (p.residual) {
            const uint4 rv = *reinterpret_cast<const uint4 *>(p.residual + m * p.ldr + n);
            const uint32_t a[4] = {v.x, v.y, v.z, v.w}, b[4] = {rv.x, rv.y, rv.z, rv.w};
            uint32_t o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float lo = __uint_as_float(a[j] << 16) + __uint_as_float(b[j] << 16);
              const float hi = __uint_as_float(a[j] & 0xffff0000u) + __uint_as_float(b[j] & 0xffff0000u);
              __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
              o[j] = *reinterpret_cast<uint32_t *>(&h);
            }
            v = make_uint4(o[0], o[1], o[2], o[3]);
          }
          *reinterpret_cast<uint4 *>(p.y + m * p.ldy + n) = v;
        }
      }
      __syncwarp();   // staging rows are rewritten by the next tile
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, C::kTmemCols);
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
  return true;
}

template <int BN>
static int launch_tc(const LinearArgs &a, cudaStream_t st) {
  using C = gemm::Cfg<BN>;
  CUtensorMap tx, tw;
  {
    const uint64_t dims[2] = {(uint64_t)a.K, (uint64_t)a.M};
    const uint64_t strides[1] = {(uint64_t)a.K * 2};
    const uint32_t box[2] = {gemm::BK, gemm::BM};
    int r = encode_tmap_bf16(&tx, a.x, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (r) return r;
  }
  {
    const uint64_t dims[2] = {(uint64_t)a.K, (uint64_t)a.N};
    const uint64_t strides[1] = {(uint64_t)a.K * 2};
    const uint32_t box[2] = {gemm::BK, BN};
    int r = encode_tmap_bf16(&tw, a.w, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (r) return r;
  }
  static bool attr_done = false;
  if (!attr_done) {
    MVIT_CUDA_OK(cudaFuncSetAttribute(gemm::linear_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
    attr_done = true;
  }
  gemm::Params p{a.bias, a.row_scale, static_cast<const bf16 *>(a.residual), static_cast<bf16 *>(a.y),
                 a.M, a.rows_per_sample, a.ldy, a.ldr, a.N, a.K, a.epilogue};
  const int64_t tiles = ((a.M + gemm::BM - 1) / gemm::BM) * ((a.N + BN - 1) / BN);
  const unsigned grid = (unsigned)std::min<int64_t>(tiles, num_sms());
  gemm::linear_tc_kernel<BN><<<grid, gemm::kThreads, C::kSmemBytes, st>>>(tx, tw, p);
  MVIT_LAUNCH_OK("linear(tcgen05)");
  return 0;
}

int linear_tc(const LinearArgs &a, cudaStream_t st) {
  const int64_t m_tiles = (a.M + gemm::BM - 1) / gemm::BM;
  const bool wide = (a.N % 192 == 0) && m_tiles * (a.N / 192) >= num_sms();
  return wide ? launch_tc<192>(a, st) : launch_tc<96>(a, st);
}

}  // namespace mvit
