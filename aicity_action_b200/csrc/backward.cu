// Backward kernels of the MViTv2 multiscale-attention path — first correct CUDA path (CUDA cores, fp32 accumulation,
// fp32 or bf16 activations; parameter gradients always fp32).  They make the drop-in modules trainable
// (tools/train_net.py:229-246 calls loss.backward() on this path; the reference relies on autograd for every op of
// attention.py / common.py).  Tensor-core versions of the two heavy ones (wgrad, attention backward) are the next step;
// dgrad GEMMs already run on the forward tcgen05 kernel (dx = dy · W is a Linear with the transposed weight).
#include "attention.cuh"
#include "common.cuh"
#include "linear.cuh"
#include "pool.cuh"

namespace mvit {

// ------------------------------------------------------------------------------------------------ LayerNorm backward
// dx = rstd * (g*dy - mean(g*dy) - xhat * mean(g*dy*xhat));  dgamma += sum_rows dy*xhat;  dbeta += sum_rows dy
// One warp per row (any C); per-CTA partial dgamma/dbeta in shared memory, one atomicAdd per channel per CTA.
template <typename T>
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const T *__restrict__ x, const float *__restrict__ gamma,
                                                            const T *__restrict__ dy, T *__restrict__ dx,
                                                            float *__restrict__ dgamma, float *__restrict__ dbeta,
                                                            int64_t rows, int C, float eps, int rows_per_cta) {
  extern __shared__ float sm[];          // dgamma[C] | dbeta[C]
  float *sg = sm, *sb = sm + C;
  for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) sm[c] = 0.f;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta;
  for (int64_t r = r0 + warp; r < min(rows, r0 + rows_per_cta); r += nwarps) {
    const T *px = x + r * C, *pdy = dy + r * C;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += to_f32(px[c]);
    const float mean = warp_sum(s) / C;
    float ss = 0.f;
    for (int c = lane; c < C; c += 32) { const float d = to_f32(px[c]) - mean; ss += d * d; }
    const float rstd = rsqrtf(warp_sum(ss) / C + eps);
    float a = 0.f, b = 0.f;                // sum g*dy, sum g*dy*xhat
    for (int c = lane; c < C; c += 32) {
      const float xh = (to_f32(px[c]) - mean) * rstd, gdy = gamma[c] * to_f32(pdy[c]);
      a += gdy;
      b += gdy * xh;
    }
    a = warp_sum(a) / C;
    b = warp_sum(b) / C;
    for (int c = lane; c < C; c += 32) {
      const float xh = (to_f32(px[c]) - mean) * rstd, d = to_f32(pdy[c]);
      dx[r * C + c] = from_f32<T>(rstd * (gamma[c] * d - a - xh * b));
      atomicAdd(&sg[c], d * xh);
      atomicAdd(&sb[c], d);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    atomicAdd(&dgamma[c], sg[c]);
    atomicAdd(&dbeta[c], sb[c]);
  }
}

// Tuned form for even C <= 64*NP: lane owns the channel pairs {lane + 32 i}; the row's x / dy stay in registers across
// the three passes, dgamma / dbeta partial sums stay in registers across ALL rows of the warp (grid-stride), and only
// one shared-memory + one global atomic per channel is issued per CTA.
template <typename T> struct Pair;
template <> struct Pair<float> {
  __device__ __forceinline__ static float2 load(const float *p) { return *reinterpret_cast<const float2 *>(p); }
  __device__ __forceinline__ static void store(float *p, float2 v) { *reinterpret_cast<float2 *>(p) = v; }
};
template <> struct Pair<bf16> {
  __device__ __forceinline__ static float2 load(const bf16 *p) {
    const uint32_t u = *reinterpret_cast<const uint32_t *>(p);
    return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
  }
  __device__ __forceinline__ static void store(bf16 *p, float2 v) {
    *reinterpret_cast<__nv_bfloat162 *>(p) = __floats2bfloat162_rn(v.x, v.y);
  }
};

template <int LPR> __device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// LPR lanes share one row (16: two rows per warp, for C = 96; 32 otherwise); lane `sub` owns channel pairs {sub + LPR i}.
template <typename T, int NP, int LPR>
__global__ void __launch_bounds__(256) layernorm_bwd_pairs_kernel(const T *__restrict__ x, const float *__restrict__ gamma,
                                                                  const T *__restrict__ dy, T *__restrict__ dx,
                                                                  float *__restrict__ dgamma, float *__restrict__ dbeta,
                                                                  int64_t rows, int C, float eps) {
  extern __shared__ float sm[];          // gamma[C] | dgamma[C] | dbeta[C]
  float *s_g = sm, *sg = sm + C, *sb = sm + 2 * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) { s_g[c] = gamma[c]; sg[c] = 0.f; sb[c] = 0.f; }
  __syncthreads();
  constexpr int RPW = 32 / LPR;          // rows per warp
  const int lane = threadIdx.x & 31, sub = lane % LPR, half = C >> 1;
  const float invC = 1.0f / C;
  float2 g[NP], adg[NP], adb[NP];
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    const int c2 = sub + LPR * i;
    g[i] = c2 < half ? make_float2(s_g[2 * c2], s_g[2 * c2 + 1]) : make_float2(0.f, 0.f);
    adg[i] = make_float2(0.f, 0.f);
    adb[i] = make_float2(0.f, 0.f);
  }
  const int64_t rstride = (int64_t)gridDim.x * (blockDim.x >> 5) * RPW;
  // every lane of a warp runs the same number of iterations (shuffles are warp-wide); rows past the end are masked
  for (int64_t r0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW; r0 < rows; r0 += rstride) {
    const int64_t r = r0 + lane / LPR;
    const bool live = r < rows;
    const T *px = x + r * C, *pdy = dy + r * C;
    float2 xv[NP], dv[NP];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      const int c2 = sub + LPR * i;
      if (live && c2 < half) {
        xv[i] = Pair<T>::load(px + 2 * c2);
        dv[i] = Pair<T>::load(pdy + 2 * c2);
      } else {
        xv[i] = make_float2(0.f, 0.f);
        dv[i] = make_float2(0.f, 0.f);
      }
      s += xv[i].x + xv[i].y;
    }
    const float mean = group_sum<LPR>(s) * invC;
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      const int c2 = sub + LPR * i;
      if (c2 < half) {
        xv[i].x -= mean; xv[i].y -= mean;
        ss = fmaf(xv[i].x, xv[i].x, fmaf(xv[i].y, xv[i].y, ss));
      }
    }
    const float rstd = rsqrtf(group_sum<LPR>(ss) * invC + eps);
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      xv[i].x *= rstd; xv[i].y *= rstd;                       // xhat (zero in unused slots)
      const float gx = g[i].x * dv[i].x, gy = g[i].y * dv[i].y;
      a += gx + gy;
      b = fmaf(gx, xv[i].x, fmaf(gy, xv[i].y, b));
      adg[i].x = fmaf(dv[i].x, xv[i].x, adg[i].x); adg[i].y = fmaf(dv[i].y, xv[i].y, adg[i].y);
      adb[i].x += dv[i].x; adb[i].y += dv[i].y;
    }
    a = group_sum<LPR>(a) * invC;
    b = group_sum<LPR>(b) * invC;
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      const int c2 = sub + LPR * i;
      if (live && c2 < half)
        Pair<T>::store(dx + r * C + 2 * c2, make_float2(rstd * (g[i].x * dv[i].x - a - xv[i].x * b),
                                                        rstd * (g[i].y * dv[i].y - a - xv[i].y * b)));
    }
  }
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    const int c2 = sub + LPR * i;
    if (c2 < half) {
      atomicAdd(&sg[2 * c2], adg[i].x); atomicAdd(&sg[2 * c2 + 1], adg[i].y);
      atomicAdd(&sb[2 * c2], adb[i].x); atomicAdd(&sb[2 * c2 + 1], adb[i].y);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    atomicAdd(&dgamma[c], sg[c]);
    atomicAdd(&dbeta[c], sb[c]);
  }
}

// ------------------------------------------------------------------------------------------------ GELU backward
// dpre = dy * d/dx[ x * Phi(x) ] = dy * (Phi(x) + x * phi(x)),  exact erf form (common.py:20 nn.GELU)
template <typename T>
__global__ void gelu_bwd_kernel(const T *__restrict__ pre, const T *__restrict__ dy, T *__restrict__ dpre, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    dpre[i] = from_f32<T>(to_f32(dy[i]) * gelu_grad(to_f32(pre[i])));
}
// 16-byte vectors (n a multiple of the vector width, aligned pointers)
template <typename T>
__global__ void __launch_bounds__(256) gelu_bwd_vec_kernel(const T *__restrict__ pre, const T *__restrict__ dy,
                                                           T *__restrict__ dpre, int64_t nvec) {
  constexpr int W = Vec16<T>::N;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    float a[W], g[W];
    Vec16<T>::load(pre + i * W, a);
    Vec16<T>::load(dy + i * W, g);
#pragma unroll
    for (int e = 0; e < W; ++e) a[e] = g[e] * gelu_grad(a[e]);
    Vec16<T>::store(dpre + i * W, a);
  }
}

// ------------------------------------------------------------------------------------------------ Linear weight / bias gradient
// dW[N, K] += dy[M, N]^T · x[M, K];  db[N] += colsum(dy).  64x64 output tile per CTA, the M axis split across gridDim.z
// CTAs that reduce with fp32 atomics.
constexpr int WT = 64, WK = 16;
template <typename T>
__global__ void __launch_bounds__(256) linear_wgrad_kernel(const T *__restrict__ dy, const T *__restrict__ x,
                                                           float *__restrict__ dw, float *__restrict__ db, int64_t M, int N,
                                                           int K, int64_t m_per_cta) {
  __shared__ float sA[WK][WT + 4];   // dy chunk: [m][n]
  __shared__ float sB[WK][WT + 4];   // x  chunk: [m][k]
  const int n0 = blockIdx.x * WT, k0 = blockIdx.y * WT;
  const int64_t m_begin = (int64_t)blockIdx.z * m_per_cta, m_end = min(M, m_begin + m_per_cta);
  const int tid = threadIdx.x, tn = (tid / 16) * 4, tk = (tid % 16) * 4;
  float acc[4][4] = {};
  float bsum = 0.f;                  // threads 0..63 of CTAs with blockIdx.y == 0 accumulate the bias gradient
  const int lr = tid / 16, lc = (tid % 16) * 4;
  for (int64_t m0 = m_begin; m0 < m_end; m0 += WK) {
    const int64_t m = m0 + lr;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int n = n0 + lc + e, k = k0 + lc + e;
      sA[lr][lc + e] = (m < m_end && n < N) ? to_f32(dy[m * N + n]) : 0.f;
      sB[lr][lc + e] = (m < m_end && k < K) ? to_f32(x[m * K + k]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int mm = 0; mm < WK; ++mm) {
      const float4 av = *reinterpret_cast<const float4 *>(&sA[mm][tn]);
      const float4 bv = *reinterpret_cast<const float4 *>(&sB[mm][tk]);
      const float ar[4] = {av.x, av.y, av.z, av.w}, br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
    if (db && blockIdx.y == 0 && tid < WT) {
#pragma unroll
      for (int mm = 0; mm < WK; ++mm) bsum += sA[mm][tid];
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tn + i, k = k0 + tk + j;
      if (n < N && k < K) atomicAdd(&dw[(int64_t)n * K + k], acc[i][j]);
    }
  if (db && blockIdx.y == 0 && tid < WT && n0 + tid < N) atomicAdd(&db[n0 + tid], bsum);
}

// ------------------------------------------------------------------------------------------------ attention backward
// Given dout [B, Lq, h*96] and the forward's lse [B, h, Lq]:   P = exp(q k^T scale - lse),  D_i = dO_i · O_i,
//   dV_j += sum_i P_ij dO_i,   dS_ij = P_ij (dO_i·V_j - D_i) scale,   dQ_i = sum_j dS_ij K_j (+ dO_i if the pooled-q
//   residual was added),   dK_j += sum_i dS_ij Q_i.      (attention.py:267-279)
// One CTA per (b, head, 32 query rows); dK / dV are accumulated with fp32 atomics into zero-initialised buffers.
constexpr int AD = 96, ABQ = 32, ABK = 32;
struct AttnBwdSmem {
  float q[ABQ][AD + 1], dO[ABQ][AD + 1], k[ABK][AD + 1], v[ABK][AD + 1];
  float p[ABQ][ABK + 1], dS[ABQ][ABK + 1], D[ABQ], L[ABQ];
};
template <typename T>
__global__ void __launch_bounds__(128) attention_bwd_kernel(const T *__restrict__ q, const T *__restrict__ k,
                                                            const T *__restrict__ v, const T *__restrict__ out,
                                                            const T *__restrict__ dout, const float *__restrict__ lse,
                                                            T *__restrict__ dq, float *__restrict__ dk,
                                                            float *__restrict__ dv, int heads, int Lq, int Lk, float scale,
                                                            int add_q) {
  extern __shared__ float attn_bwd_smem[];
  AttnBwdSmem &S = *reinterpret_cast<AttnBwdSmem *>(attn_bwd_smem);
  auto &sQ = S.q; auto &sdO = S.dO; auto &sK = S.k; auto &sV = S.v; auto &sP = S.p; auto &sdS = S.dS; auto &sD = S.D; auto &sL = S.L;
  const int tid = threadIdx.x, bh = blockIdx.y, b = bh / heads, head = bh % heads, q0 = blockIdx.x * ABQ;
  const T *qp = q + (int64_t)bh * Lq * AD, *kp = k + (int64_t)bh * Lk * AD, *vp = v + (int64_t)bh * Lk * AD;
  for (int i = tid; i < ABQ * AD; i += 128) {
    const int r = i / AD, c = i % AD;
    const bool ok = q0 + r < Lq;
    const int64_t o = (((int64_t)b * Lq + q0 + r) * heads + head) * AD + c;
    sQ[r][c] = ok ? to_f32(qp[(int64_t)(q0 + r) * AD + c]) : 0.f;
    sdO[r][c] = ok ? to_f32(dout[o]) : 0.f;
  }
  __syncthreads();
  if (tid < ABQ) {
    float d = 0.f;
    if (q0 + tid < Lq) {
      for (int c = 0; c < AD; ++c) {
        const int64_t o = (((int64_t)b * Lq + q0 + tid) * heads + head) * AD + c;
        const float oa = to_f32(out[o]) - (add_q ? sQ[tid][c] : 0.f);   // attention output before the residual add
        d += sdO[tid][c] * oa;
      }
      sL[tid] = lse[(int64_t)bh * Lq + q0 + tid];
    } else {
      sL[tid] = 0.f;
    }
    sD[tid] = d;
  }
  const int r = tid / 4, sub = tid % 4;          // S/P mapping: row r, cols sub + 4j;  dQ mapping: row r, channels sub + 4i
  float dqa[24];
#pragma unroll
  for (int i = 0; i < 24; ++i) dqa[i] = 0.f;
  for (int k0 = 0; k0 < Lk; k0 += ABK) {
    __syncthreads();
    for (int i = tid; i < ABK * AD; i += 128) {
      const int rr = i / AD, c = i % AD;
      const bool ok = k0 + rr < Lk;
      sK[rr][c] = ok ? to_f32(kp[(int64_t)(k0 + rr) * AD + c]) : 0.f;
      sV[rr][c] = ok ? to_f32(vp[(int64_t)(k0 + rr) * AD + c]) : 0.f;
    }
    __syncthreads();
    float s[8], dp[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { s[j] = 0.f; dp[j] = 0.f; }
    for (int c = 0; c < AD; ++c) {
      const float qv = sQ[r][c], dov = sdO[r][c];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s[j] = fmaf(qv, sK[sub + 4 * j][c], s[j]);
        dp[j] = fmaf(dov, sV[sub + 4 * j][c], dp[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const bool ok = (k0 + sub + 4 * j < Lk) && (q0 + r < Lq);
      const float pj = ok ? expf(s[j] * scale - sL[r]) : 0.f;
      sP[r][sub + 4 * j] = pj;
      sdS[r][sub + 4 * j] = pj * (dp[j] - sD[r]) * scale;
    }
    __syncthreads();
    // dQ_r += dS_r · K
    for (int c = 0; c < ABK; ++c) {
      const float ds = sdS[r][c];
#pragma unroll
      for (int i = 0; i < 24; ++i) dqa[i] = fmaf(ds, sK[c][sub + 4 * i], dqa[i]);
    }
    // dK_j += dS^T Q, dV_j += P^T dO : thread -> key row kr = tid / 4, channels sub + 4i
    {
      const int kr = tid / 4;
      if (k0 + kr < Lk) {
        float dka[24], dva[24];
#pragma unroll
        for (int i = 0; i < 24; ++i) { dka[i] = 0.f; dva[i] = 0.f; }
        for (int rr = 0; rr < ABQ; ++rr) {
          const float ds = sdS[rr][kr], pp = sP[rr][kr];
#pragma unroll
          for (int i = 0; i < 24; ++i) {
            dka[i] = fmaf(ds, sQ[rr][sub + 4 * i], dka[i]);
            dva[i] = fmaf(pp, sdO[rr][sub + 4 * i], dva[i]);
          }
        }
        float *dkp = dk + ((int64_t)bh * Lk + k0 + kr) * AD, *dvp = dv + ((int64_t)bh * Lk + k0 + kr) * AD;
#pragma unroll
        for (int i = 0; i < 24; ++i) {
          atomicAdd(&dkp[sub + 4 * i], dka[i]);
          atomicAdd(&dvp[sub + 4 * i], dva[i]);
        }
      }
    }
  }
  if (q0 + r < Lq) {
    T *dqp = dq + ((int64_t)bh * Lq + q0 + r) * AD;
#pragma unroll
    for (int i = 0; i < 24; ++i) dqp[sub + 4 * i] = from_f32<T>(dqa[i] + (add_q ? sdO[r][sub + 4 * i] : 0.f));
  }
}

// ------------------------------------------------------------------------------------------------ attention_pool backward
struct PoolBwdParams {
  int64_t x_bs, x_ls, x_hs;            // strides of the pooling input / its gradient (elements)
  int B, heads, d, T, H, W, kt, kh, kw, st, sh, sw, pt, ph, pw, To, Ho, Wo;
  int dy_head_major;                   // max-pool gradient: dy is [B, heads, L', d] (1) or tokens [B, L', heads*d] (0)
};

// conv input gradient, gather form (deterministic): dx[pos] = sum over taps with (pos + pad - tap) % stride == 0 of
// w[tap] * dy[(pos + pad - tap) / stride].  One warp per input token*head; lane owns channels lane + 32j.
template <typename T, int NC>
__global__ void __launch_bounds__(256) pool_conv_dgrad_kernel(const T *__restrict__ dy, const float *__restrict__ weight,
                                                              T *__restrict__ dx, PoolBwdParams p) {
  extern __shared__ float w_s[];       // [taps][d]
  const int taps = p.kt * p.kh * p.kw;
  for (int i = threadIdx.x; i < taps * p.d; i += blockDim.x) {
    const int tap = i / p.d, c = i - tap * p.d;
    w_s[i] = weight[c * taps + tap];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int L = p.T * p.H * p.W, Lo = p.To * p.Ho * p.Wo;
  const int64_t total = (int64_t)p.B * L * p.heads;
  const int64_t wstride = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t o = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); o < total; o += wstride) {
  const int head = (int)(o % p.heads);
  const int64_t bl = o / p.heads;
  const int l = (int)(bl % L), b = (int)(bl / L);
  const int w = l % p.W, h = (l / p.W) % p.H, t = l / (p.W * p.H);
  float acc[NC];
#pragma unroll
  for (int j = 0; j < NC; ++j) acc[j] = 0.f;
  const T *dyb = dy + ((int64_t)(b * p.heads + head) * Lo) * p.d;
  // output index feeding this input position through tap a / bq / c of each axis (or -1): 9 divisions per position
  int tv[3], hv[3], wv[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const int tn = t + p.pt - a, hn = h + p.ph - a, wn = w + p.pw - a;
    const int tq = tn / p.st, hq = hn / p.sh, wq = wn / p.sw;
    tv[a] = (a < p.kt && tn >= 0 && tq * p.st == tn && tq < p.To) ? tq : -1;
    hv[a] = (a < p.kh && hn >= 0 && hq * p.sh == hn && hq < p.Ho) ? hq : -1;
    wv[a] = (a < p.kw && wn >= 0 && wq * p.sw == wn && wq < p.Wo) ? wq : -1;
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    if (tv[a] < 0) continue;
#pragma unroll
    for (int bq = 0; bq < 3; ++bq) {
      if (hv[bq] < 0) continue;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        if (wv[c] < 0) continue;
        const T *pdy = dyb + (int64_t)((tv[a] * p.Ho + hv[bq]) * p.Wo + wv[c]) * p.d;
        const float *pw = w_s + ((a * p.kh + bq) * p.kw + c) * p.d;
#pragma unroll
        for (int j = 0; j < NC; ++j) acc[j] = fmaf(to_f32(pdy[lane + 32 * j]), pw[lane + 32 * j], acc[j]);
      }
    }
  }
  T *dst = dx + b * p.x_bs + (int64_t)l * p.x_ls + head * p.x_hs;
#pragma unroll
  for (int j = 0; j < NC; ++j) dst[lane + 32 * j] = from_f32<T>(acc[j]);
  }
}

// Vectorised, divergence-free form of the same gather.  thread = (input token, 16-byte channel group); at any time a CTA
// only holds tokens of ONE frame t and ONE residue class (h mod sh, w mod sw), so which of the <= 27 taps reach an output
// is decided per CTA: the contributing taps (1.5 per strided axis on average) are compacted into a short list in shared
// memory and every thread runs a branch-free loop over it (border taps are masked, not skipped), which lets the dy loads
// of several taps be in flight together.  blockIdx = (class-and-chunk id of a grid-stride loop, batch*head).
template <typename T>
__global__ void __launch_bounds__(256) pool_conv_dgrad_vec_kernel(const T *__restrict__ dy, const float *__restrict__ weight,
                                                                  T *__restrict__ dx, PoolBwdParams p, int nsplit) {
  constexpr int N = Vec16<T>::N;
  extern __shared__ __align__(16) float w_s[];       // [taps][d]
  __shared__ int4 s_list[27];                        // (weight row offset, output frame, dho, dwo) of the live taps
  __shared__ int s_n;
  const int taps = p.kt * p.kh * p.kw;
  for (int i = threadIdx.x; i < taps * p.d; i += blockDim.x) {
    const int tap = i / p.d, c = i - tap * p.d;
    w_s[i] = weight[c * taps + tap];
  }
  const int groups = p.d / N;
  const int Lo = p.To * p.Ho * p.Wo;
  const int Hc = (p.H + p.sh - 1) / p.sh, Wc = (p.W + p.sw - 1) / p.sw;
  const int b = blockIdx.y / p.heads, head = blockIdx.y % p.heads;
  const T *dy_bh = dy + (int64_t)blockIdx.y * Lo * p.d;
  T *dx_bh = dx + b * p.x_bs + head * p.x_hs;
  const int total = Hc * Wc * groups;
  const int classes = p.sh * p.sw * p.T;
  for (int work = blockIdx.x; work < classes * nsplit; work += gridDim.x) {
    const int cls = work / nsplit, split = work - cls * nsplit;
    const int t = cls % p.T, res = cls / p.T, rh = res / p.sw, rw = res % p.sw;
    __syncthreads();                                 // previous list no longer in use (and w_s staged, first pass)
    if (threadIdx.x == 0) {
      int n = 0;
      for (int a = 0; a < p.kt; ++a) {
        const int tn = t + p.pt - a;
        if (tn < 0 || tn % p.st != 0 || tn / p.st >= p.To) continue;
        for (int bq = 0; bq < p.kh; ++bq) {
          const int eh = rh + p.ph - bq;
          if (eh % p.sh != 0) continue;
          for (int c = 0; c < p.kw; ++c) {
            const int ew = rw + p.pw - c;
            if (ew % p.sw != 0) continue;
            s_list[n++] = make_int4(((a * p.kh + bq) * p.kw + c) * p.d, tn / p.st, eh / p.sh, ew / p.sw);
          }
        }
      }
      s_n = n;
    }
    __syncthreads();
    const int nv = s_n;
    for (int i = split * blockDim.x + threadIdx.x; i < total; i += nsplit * blockDim.x) {
      const int r = i / groups, g = i - r * groups;
      const int hq = r / Wc, wq = r - hq * Wc;
      const int h = hq * p.sh + rh, w = wq * p.sw + rw;
      if (h >= p.H || w >= p.W) continue;
      float acc[N];
#pragma unroll
      for (int e = 0; e < N; ++e) acc[e] = 0.f;
      const T *dyb = dy_bh + g * N;
#pragma unroll 4
      for (int n = 0; n < nv; ++n) {
        const int4 e4 = s_list[n];
        const int ho = hq + e4.z, wo = wq + e4.w;
        const bool ok = (unsigned)ho < (unsigned)p.Ho && (unsigned)wo < (unsigned)p.Wo;
        float v[N];
        Vec16<T>::load(dyb + (ok ? ((e4.y * p.Ho + ho) * p.Wo + wo) * p.d : 0), v);
        const float m = ok ? 1.f : 0.f;
        const float *pw = w_s + e4.x + g * N;
#pragma unroll
        for (int e = 0; e < N; e += 4) {
          const float4 w4 = *reinterpret_cast<const float4 *>(pw + e);
          acc[e] = fmaf(v[e] * m, w4.x, acc[e]);
          acc[e + 1] = fmaf(v[e + 1] * m, w4.y, acc[e + 1]);
          acc[e + 2] = fmaf(v[e + 2] * m, w4.z, acc[e + 2]);
          acc[e + 3] = fmaf(v[e + 3] * m, w4.w, acc[e + 3]);
        }
      }
      const int l = (t * p.H + h) * p.W + w;
      Vec16<T>::store(dx_bh + (int64_t)l * p.x_ls + g * N, acc);
    }
  }
}

// conv weight gradient: dW[c][tap] += sum over (b, head, output position) x[input(tap)] * dy.  One warp per output
// token*head accumulates 27 x NC products per lane in registers over a strip of outputs, then atomics.
template <typename T, int NC>
__global__ void __launch_bounds__(256) pool_conv_wgrad_kernel(const T *__restrict__ x, const T *__restrict__ dy,
                                                              float *__restrict__ dw, PoolBwdParams p, int outs_per_warp) {
  const int lane = threadIdx.x & 31;
  const int taps = p.kt * p.kh * p.kw;   // <= 27 supported by the register accumulator below
  const int Lo = p.To * p.Ho * p.Wo;
  const int64_t total = (int64_t)p.B * Lo * p.heads;
  const int64_t o0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * outs_per_warp;
  float acc[27][NC];
#pragma unroll
  for (int i = 0; i < 27; ++i)
#pragma unroll
    for (int j = 0; j < NC; ++j) acc[i][j] = 0.f;
  for (int64_t o = o0; o < min(total, o0 + outs_per_warp); ++o) {
    const int head = (int)(o % p.heads);
    const int64_t bl = o / p.heads;
    const int lo = (int)(bl % Lo), b = (int)(bl / Lo);
    const int wo = lo % p.Wo, ho = (lo / p.Wo) % p.Ho, to = lo / (p.Wo * p.Ho);
    const T *pdy = dy + ((int64_t)(b * p.heads + head) * Lo + lo) * p.d;
    float g[NC];
#pragma unroll
    for (int j = 0; j < NC; ++j) g[j] = to_f32(pdy[lane + 32 * j]);
    const T *src = x + b * p.x_bs + head * p.x_hs;
#pragma unroll
    for (int tap = 0; tap < 27; ++tap) {
      if (tap >= taps) break;
      const int c = tap % p.kw, bq = (tap / p.kw) % p.kh, a = tap / (p.kw * p.kh);
      const int t = to * p.st - p.pt + a, h = ho * p.sh - p.ph + bq, w = wo * p.sw - p.pw + c;
      if (t < 0 || t >= p.T || h < 0 || h >= p.H || w < 0 || w >= p.W) continue;
      const T *px = src + (int64_t)((t * p.H + h) * p.W + w) * p.x_ls;
#pragma unroll
      for (int j = 0; j < NC; ++j) acc[tap][j] = fmaf(to_f32(px[lane + 32 * j]), g[j], acc[tap][j]);
    }
  }
#pragma unroll
  for (int tap = 0; tap < 27; ++tap) {
    if (tap >= taps) break;
#pragma unroll
    for (int j = 0; j < NC; ++j) atomicAdd(&dw[(lane + 32 * j) * taps + tap], acc[tap][j]);
  }
}

// max-pool backward: the gradient of every output goes to the arg-max input of its window (first maximum in scan
// order, as ATen's max_pool3d_with_indices).  dx must be zero-initialised (fp32); windows overlap -> atomics.
template <typename T, int NC>
__global__ void __launch_bounds__(256) pool_max_bwd_kernel(const T *__restrict__ x, const T *__restrict__ dy,
                                                           float *__restrict__ dx, PoolBwdParams p) {
  const int lane = threadIdx.x & 31;
  const int Lo = p.To * p.Ho * p.Wo, L = p.T * p.H * p.W;
  const int64_t total = (int64_t)p.B * Lo * p.heads;
  const int64_t o = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (o >= total) return;
  const int head = (int)(o % p.heads);
  const int64_t bl = o / p.heads;
  const int lo = (int)(bl % Lo), b = (int)(bl / Lo);
  const int wo = lo % p.Wo, ho = (lo / p.Wo) % p.Ho, to = lo / (p.Wo * p.Ho);
  float best[NC];
  int arg[NC];
#pragma unroll
  for (int j = 0; j < NC; ++j) { best[j] = -INFINITY; arg[j] = -1; }
  const T *src = x + b * p.x_bs + head * p.x_hs;
  for (int a = 0; a < p.kt; ++a) {
    const int t = to * p.st - p.pt + a;
    if (t < 0 || t >= p.T) continue;
    for (int bq = 0; bq < p.kh; ++bq) {
      const int h = ho * p.sh - p.ph + bq;
      if (h < 0 || h >= p.H) continue;
      for (int c = 0; c < p.kw; ++c) {
        const int w = wo * p.sw - p.pw + c;
        if (w < 0 || w >= p.W) continue;
        const int l = (t * p.H + h) * p.W + w;
        const T *px = src + (int64_t)l * p.x_ls;
#pragma unroll
        for (int j = 0; j < NC; ++j) {
          const float v = to_f32(px[lane + 32 * j]);
          if (v > best[j] || arg[j] < 0) { best[j] = v; arg[j] = l; }
        }
      }
    }
  }
  const T *pdy = p.dy_head_major ? dy + ((int64_t)(b * p.heads + head) * Lo + lo) * p.d
                                 : dy + (((int64_t)b * Lo + lo) * p.heads + head) * p.d;   // token layout [B, L', heads*d]
  // dx is a dense fp32 [B, L, heads, d] buffer
#pragma unroll
  for (int j = 0; j < NC; ++j)
    if (arg[j] >= 0) atomicAdd(&dx[(((int64_t)b * L + arg[j]) * p.heads + head) * p.d + lane + 32 * j], to_f32(pdy[lane + 32 * j]));
}

}  // namespace mvit

// ================================================================================================ C ABI
using namespace mvit;

extern "C" int mvit_layernorm_bwd(const void *x, const float *gamma, const void *dy, void *dx, float *dgamma,
                                  float *dbeta, int64_t rows, int channels, float eps, int dtype, void *stream) {
  MVIT_REQUIRE(x && gamma && dy && dx && dgamma && dbeta, "layernorm_bwd: null pointer");
  MVIT_REQUIRE(rows >= 0 && channels > 0 && channels <= 4096, "layernorm_bwd: bad shape");
  MVIT_REQUIRE(dtype == MVIT_F32 || dtype == MVIT_BF16, "layernorm_bwd: unknown dtype");
  if (rows == 0) return 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool aligned = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dx)) & 7) == 0;
  if (channels % 2 == 0 && channels <= 768 && aligned) {
    // one wave of resident CTAs (register-limited: 1 / 2 / 4 per SM for C <= 768 / 384 / 192), at least 64 rows each, so
    // the per-CTA dgamma/dbeta reduction (2C atomics) is paid a few hundred times, not once per 16 rows
    const int per_sm = channels > 384 ? 1 : (channels > 192 ? 2 : 4);
    const unsigned blocks = (unsigned)std::max<int64_t>(1, std::min<int64_t>((rows + 63) / 64, (int64_t)num_sms() * per_sm));
    const size_t smem3 = 3 * (size_t)channels * sizeof(float);
#define LN_BWD_PAIRS(T, NP, LPR)                                                                                       \
  layernorm_bwd_pairs_kernel<T, NP, LPR><<<blocks, 256, smem3, st>>>(static_cast<const T *>(x), gamma,                 \
                                                                     static_cast<const T *>(dy), static_cast<T *>(dx), \
                                                                     dgamma, dbeta, rows, channels, eps)
    if (dtype == MVIT_F32) {
      if (channels <= 96) LN_BWD_PAIRS(float, 3, 16);
      else if (channels <= 192) LN_BWD_PAIRS(float, 3, 32);
      else if (channels <= 384) LN_BWD_PAIRS(float, 6, 32);
      else LN_BWD_PAIRS(float, 12, 32);
    } else {
      if (channels <= 96) LN_BWD_PAIRS(bf16, 3, 16);
      else if (channels <= 192) LN_BWD_PAIRS(bf16, 3, 32);
      else if (channels <= 384) LN_BWD_PAIRS(bf16, 6, 32);
      else LN_BWD_PAIRS(bf16, 12, 32);
    }
#undef LN_BWD_PAIRS
    MVIT_LAUNCH_OK("layernorm_bwd");
    return 0;
  }
  const int rows_per_cta = 64;
  const unsigned blocks = (unsigned)((rows + rows_per_cta - 1) / rows_per_cta);
  const size_t smem = 2 * (size_t)channels * sizeof(float);
  if (dtype == MVIT_F32)
    layernorm_bwd_kernel<float><<<blocks, 256, smem, st>>>(static_cast<const float *>(x), gamma, static_cast<const float *>(dy), static_cast<float *>(dx), dgamma, dbeta, rows, channels, eps, rows_per_cta);
  else
    layernorm_bwd_kernel<bf16><<<blocks, 256, smem, st>>>(static_cast<const bf16 *>(x), gamma, static_cast<const bf16 *>(dy), static_cast<bf16 *>(dx), dgamma, dbeta, rows, channels, eps, rows_per_cta);
  MVIT_LAUNCH_OK("layernorm_bwd");
  return 0;
}

extern "C" int mvit_gelu_bwd(const void *pre, const void *dy, void *dpre, int64_t n, int dtype, void *stream) {
  MVIT_REQUIRE(pre && dy && dpre && n >= 0, "gelu_bwd: bad arguments");
  MVIT_REQUIRE(dtype == MVIT_F32 || dtype == MVIT_BF16, "gelu_bwd: unknown dtype");
  if (n == 0) return 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int vw = dtype == MVIT_F32 ? 4 : 8;
  const bool aligned = ((reinterpret_cast<uintptr_t>(pre) | reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dpre)) & 15) == 0;
  if (aligned && n % vw == 0) {
    const int64_t nvec = n / vw;
    const unsigned vblocks = (unsigned)std::min<int64_t>((nvec + 255) / 256, (int64_t)num_sms() * 16);
    if (dtype == MVIT_F32) gelu_bwd_vec_kernel<float><<<vblocks, 256, 0, st>>>(static_cast<const float *>(pre), static_cast<const float *>(dy), static_cast<float *>(dpre), nvec);
    else gelu_bwd_vec_kernel<bf16><<<vblocks, 256, 0, st>>>(static_cast<const bf16 *>(pre), static_cast<const bf16 *>(dy), static_cast<bf16 *>(dpre), nvec);
    MVIT_LAUNCH_OK("gelu_bwd");
    return 0;
  }
  const unsigned blocks = (unsigned)std::min<int64_t>((n + 255) / 256, (int64_t)num_sms() * 16);
  if (dtype == MVIT_F32) gelu_bwd_kernel<float><<<blocks, 256, 0, st>>>(static_cast<const float *>(pre), static_cast<const float *>(dy), static_cast<float *>(dpre), n);
  else gelu_bwd_kernel<bf16><<<blocks, 256, 0, st>>>(static_cast<const bf16 *>(pre), static_cast<const bf16 *>(dy), static_cast<bf16 *>(dpre), n);
  MVIT_LAUNCH_OK("gelu_bwd");
  return 0;
}

extern "C" int mvit_linear_wgrad(const void *dy, const void *x, float *dw, float *db, int64_t M, int N, int K, int dtype,
                                 int impl, void *stream) {
  MVIT_REQUIRE(dy && x && dw, "linear_wgrad: null pointer");
  MVIT_REQUIRE(M >= 0 && N > 0 && K > 0, "linear_wgrad: bad shape");
  MVIT_REQUIRE(dtype == MVIT_F32 || dtype == MVIT_BF16, "linear_wgrad: unknown dtype");
  if (M == 0) return 0;
  if (impl != MVIT_IMPL_SIMT) {
    const char *why = "fp32 runs on CUDA cores";
    if (dtype == MVIT_BF16 && linear_wgrad_tc_supported(dy, x, dw, M, N, K, &why))
      return linear_wgrad_tc(dy, x, dw, db, M, N, K, static_cast<cudaStream_t>(stream));
    MVIT_REQUIRE(impl != MVIT_IMPL_TCGEN05, "linear_wgrad: tcgen05 path unavailable: %s", why);
  }
  const int gx = (N + WT - 1) / WT, gy = (K + WT - 1) / WT;
  int splits = (int)std::min<int64_t>(std::max<int64_t>(1, (2 * num_sms()) / (gx * gy)), (M + 255) / 256);
  splits = std::max(1, std::min(splits, 65535));
  int64_t m_per_cta = ((M + splits - 1) / splits + WK - 1) / WK * WK;
  dim3 grid(gx, gy, (unsigned)((M + m_per_cta - 1) / m_per_cta));
  MVIT_REQUIRE(gy < 65536, "linear_wgrad: K too large");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == MVIT_F32) linear_wgrad_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float *>(dy), static_cast<const float *>(x), dw, db, M, N, K, m_per_cta);
  else linear_wgrad_kernel<bf16><<<grid, 256, 0, st>>>(static_cast<const bf16 *>(dy), static_cast<const bf16 *>(x), dw, db, M, N, K, m_per_cta);
  MVIT_LAUNCH_OK("linear_wgrad");
  return 0;
}

extern "C" size_t mvit_attention_bwd_workspace_floats(int B, int heads, int Lq) {
  return attention_bwd_workspace_floats(B, heads, Lq);
}

extern "C" int mvit_attention_bwd(const void *q, const void *k, const void *v, const void *out, const void *dout,
                                  const float *lse, void *dq, float *dk, float *dv, float *workspace, int B, int heads,
                                  int Lq, int Lk, int d, float scale, int add_q_residual, int dtype, int impl,
                                  void *stream) {
  MVIT_REQUIRE(q && k && v && out && dout && lse && dq && dk && dv, "attention_bwd: null pointer");
  MVIT_REQUIRE(B >= 0 && heads > 0 && Lq > 0 && Lk > 0, "attention_bwd: bad shape");
  MVIT_REQUIRE(d == 96, "attention_bwd: head_dim %d unsupported", d);
  MVIT_REQUIRE(dtype == MVIT_F32 || dtype == MVIT_BF16, "attention_bwd: unknown dtype");
  MVIT_REQUIRE((int64_t)B * heads < 65536, "attention_bwd: B*heads too large");
  if (B == 0) return 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (impl != MVIT_IMPL_SIMT) {
    AttnBwdArgs a{q, k, v, out, dout, lse, dq, dk, dv, workspace, B, heads, Lq, Lk, scale, add_q_residual ? 1 : 0};
    const char *why = "fp32 runs on CUDA cores";
    if (dtype == MVIT_BF16 && attention_bwd_tc_supported(a, &why)) return attention_bwd_tc(a, st);
    MVIT_REQUIRE(impl != MVIT_IMPL_TCGEN05, "attention_bwd: tcgen05 path unavailable: %s", why);
  }
  dim3 grid((Lq + ABQ - 1) / ABQ, B * heads);
  const size_t smem = sizeof(AttnBwdSmem);
  MVIT_SMEM_OPT_IN(attention_bwd_kernel<float>, smem);
  MVIT_SMEM_OPT_IN(attention_bwd_kernel<bf16>, smem);
  if (dtype == MVIT_F32)
    attention_bwd_kernel<float><<<grid, 128, smem, st>>>(static_cast<const float *>(q), static_cast<const float *>(k), static_cast<const float *>(v), static_cast<const float *>(out), static_cast<const float *>(dout), lse, static_cast<float *>(dq), dk, dv, heads, Lq, Lk, scale, add_q_residual ? 1 : 0);
  else
    attention_bwd_kernel<bf16><<<grid, 128, smem, st>>>(static_cast<const bf16 *>(q), static_cast<const bf16 *>(k), static_cast<const bf16 *>(v), static_cast<const bf16 *>(out), static_cast<const bf16 *>(dout), lse, static_cast<bf16 *>(dq), dk, dv, heads, Lq, Lk, scale, add_q_residual ? 1 : 0);
  MVIT_LAUNCH_OK("attention_bwd");
  return 0;
}

template <typename T>
static int pool_bwd_dispatch(int what, const void *x, const void *dy, const float *weight, void *dx, float *dw,
                             const PoolBwdParams &p, cudaStream_t st) {
  const int nc = p.d / 32;
  const int Lo = p.To * p.Ho * p.Wo, L = p.T * p.H * p.W;
  const size_t wsm = (size_t)p.kt * p.kh * p.kw * p.d * sizeof(float);
  const int64_t vb = 16 / (int64_t)sizeof(T);
  const bool vec_ok = what == 0 && (int64_t)p.B * p.heads < 65536 && (int64_t)p.sh * p.sw * p.T < (1 << 20) &&
                      (int64_t)L * p.d < ((int64_t)1 << 30) &&
                      ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dx)) & 15) == 0 &&
                      p.x_bs % vb == 0 && p.x_ls % vb == 0 && p.x_hs % vb == 0 && p.d % vb == 0;
#define POOL_BWD_CASE(NC)                                                                                             \
  case NC:                                                                                                            \
    if (what == 0 && vec_ok) {                                                                                        \
      const int classes = p.sh * p.sw * p.T, bh = p.B * p.heads;                                                      \
      const int64_t items = (int64_t)((p.H + p.sh - 1) / p.sh) * ((p.W + p.sw - 1) / p.sw) * (p.d / Vec16<T>::N);     \
      const int want = (16 * num_sms() + bh - 1) / bh;                 /* CTAs per (batch, head) */                    \
      const int nsplit = (int)std::max<int64_t>(1, std::min<int64_t>((want + classes - 1) / classes, (items + 511) / 512)); \
      dim3 grid((unsigned)std::max(1, std::min(classes * nsplit, want)), (unsigned)bh);                               \
      pool_conv_dgrad_vec_kernel<T><<<grid, 256, wsm, st>>>(static_cast<const T *>(dy), weight, static_cast<T *>(dx), p, nsplit); \
    } else if (what == 0) {                                                                                           \
      const int64_t blocks = std::min<int64_t>(((int64_t)p.B * L * p.heads + 7) / 8, (int64_t)num_sms() * 16);       \
      pool_conv_dgrad_kernel<T, NC><<<(unsigned)blocks, 256, wsm, st>>>(static_cast<const T *>(dy), weight, static_cast<T *>(dx), p); \
    } else if (what == 1) {                                                                                           \
      const int opw = 64;                                                                                             \
      const int64_t warps = ((int64_t)p.B * Lo * p.heads + opw - 1) / opw;                                            \
      pool_conv_wgrad_kernel<T, NC><<<(unsigned)((warps + 7) / 8), 256, 0, st>>>(static_cast<const T *>(x), static_cast<const T *>(dy), dw, p, opw); \
    } else {                                                                                                          \
      const int64_t blocks = ((int64_t)p.B * Lo * p.heads + 7) / 8;                                                   \
      pool_max_bwd_kernel<T, NC><<<(unsigned)blocks, 256, 0, st>>>(static_cast<const T *>(x), static_cast<const T *>(dy), static_cast<float *>(dx), p); \
    }                                                                                                                 \
    break;
  switch (nc) {
    POOL_BWD_CASE(1)
    POOL_BWD_CASE(2)
    POOL_BWD_CASE(3)
    POOL_BWD_CASE(4)
    default: MVIT_REQUIRE(false, "attention_pool_bwd: head_dim %d unsupported", p.d);
  }
#undef POOL_BWD_CASE
  MVIT_LAUNCH_OK("attention_pool_bwd");
  return 0;
}

// what: 0 = conv input gradient (dx, strided like the forward input), 1 = conv weight gradient (dw[d, taps] +=, needs x),
//       2 = max-pool input gradient (dx: zero-initialised dense fp32 [B, L, heads, d], needs x)
extern "C" int mvit_attention_pool_bwd(int what, const void *x, int64_t x_bs, int64_t x_ls, int64_t x_hs, const void *dy,
                                       const float *weight, void *dx, float *dw, int B, int heads, int d, int T, int H,
                                       int W, int kt, int kh, int kw, int st, int sh, int sw, int dtype, void *stream) {
  MVIT_REQUIRE(dy, "attention_pool_bwd: null pointer");
  MVIT_REQUIRE(what >= 0 && what <= 3, "attention_pool_bwd: unknown gradient kind %d", what);
  MVIT_REQUIRE(d % 32 == 0 && d <= 128, "attention_pool_bwd: head_dim %d unsupported", d);
  MVIT_REQUIRE(kt * kh * kw <= 27, "attention_pool_bwd: kernel larger than 27 taps unsupported");
  MVIT_REQUIRE(dtype == MVIT_F32 || dtype == MVIT_BF16, "attention_pool_bwd: unknown dtype");
  MVIT_REQUIRE((what == 0 && weight && dx) || (what == 1 && x && dw) || (what >= 2 && x && dx), "attention_pool_bwd: missing operand");
  if (B == 0) return 0;
  PoolBwdParams p;
  p.x_bs = x_bs; p.x_ls = x_ls; p.x_hs = x_hs;
  p.B = B; p.heads = heads; p.d = d; p.T = T; p.H = H; p.W = W;
  p.kt = kt; p.kh = kh; p.kw = kw; p.st = st; p.sh = sh; p.sw = sw;
  p.pt = kt / 2; p.ph = kh / 2; p.pw = kw / 2;
  p.To = (T + 2 * p.pt - kt) / st + 1; p.Ho = (H + 2 * p.ph - kh) / sh + 1; p.Wo = (W + 2 * p.pw - kw) / sw + 1;
  p.dy_head_major = what == 3 ? 1 : 0;
  if (what == 3) what = 2;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (what == 1) {
    PoolParams q{};
    q.in_bs = x_bs; q.in_ls = x_ls; q.in_hs = x_hs;
    q.B = B; q.heads = heads; q.d = d; q.T = T; q.H = H; q.W = W;
    q.kt = kt; q.kh = kh; q.kw = kw; q.st = st; q.sh = sh; q.sw = sw;
    q.pt = p.pt; q.ph = p.ph; q.pw = p.pw; q.To = p.To; q.Ho = p.Ho; q.Wo = p.Wo;
    const int r = pool_wgrad_tiled_try(x, dy, dw, q, dtype, s);
    if (r <= 0) return r;
  }
  if (dtype == MVIT_F32) return pool_bwd_dispatch<float>(what, x, dy, weight, dx, dw, p, s);
  return pool_bwd_dispatch<bf16>(what, x, dy, weight, dx, dw, p, s);
}
