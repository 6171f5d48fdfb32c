// Training-step glue: AdamW + global-norm gradient clipping over FLAT parameter / gradient / moment
// arenas, in two launches per step instead of ~700 tiny PyTorch kernels.
//
// Replaces, for the B200 training driver, the tail of the reference's iteration (tools/train_net.py:229-246):
//   optimizer.zero_grad()                                -> one memset of the gradient arena (host side: arena.zero_())
//   torch.nn.utils.clip_grad_norm_(params, 1.0)          -> grad_sqnorm_kernel (+ the scale folded into the update)
//   optimizer.step()   [torch.optim.AdamW, optimizer.py:200-206]  -> adamw_kernel
//   the per-parameter fp32 -> bf16 casts of the next forward      -> the bf16 shadow written by adamw_kernel
//
// Semantics are torch's (decoupled weight decay, bias-corrected moments, eps added after the sqrt / sqrt(bias2)):
//   g' = g * min(1, max_norm / (||g||_2 + 1e-6))            (clip_grad_norm_; max_norm <= 0 disables)
//   p  = p * (1 - lr * wd);  m = b1 m + (1 - b1) g';  v = b2 v + (1 - b2) g'^2
//   p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// The arena is laid out [decayed parameters | non-decayed parameters] so weight decay is a split index, not a table.
// Hyper-parameters that change every iteration (lr, the step counter t) live in a small DEVICE struct: the two launches
// read them from memory, so a captured CUDA graph of the whole training step replays with a new lr / t.
// HBM-bound: 4 fp32 streams read + 3 written + 1 bf16 written = 30 B per parameter (35.3 M parameters -> 1.06 GB,
// ~0.17 ms at the measured 6.45 TB/s).
#include <algorithm>

#include "common.cuh"

namespace mvit {
namespace {

constexpr int kThreads = 256;
constexpr int kMaxPartials = 1024;

struct Hyper {          // mirrors the float32[8] tensor the host keeps on the device
  float lr, beta1, beta2, eps, weight_decay, max_norm;
  int step;             // number of completed steps; incremented by grad_sqnorm_kernel
  float grad_norm;      // ||g||_2 of the last step (before clipping), for logging
};

__global__ void __launch_bounds__(kThreads) grad_sqnorm_kernel(const float *__restrict__ g, int64_t n,
                                                               float *__restrict__ partials, Hyper *hyper) {
  const int64_t nv = n >> 2;
  float acc = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < nv; i += (int64_t)gridDim.x * kThreads) {
    const float4 v = __ldg(reinterpret_cast<const float4 *>(g) + i);
    acc = fmaf(v.x, v.x, acc); acc = fmaf(v.y, v.y, acc); acc = fmaf(v.z, v.z, acc); acc = fmaf(v.w, v.w, acc);
  }
  if (blockIdx.x == 0)
    for (int64_t i = (nv << 2) + threadIdx.x; i < n; i += kThreads) acc = fmaf(g[i], g[i], acc);
  __shared__ float red[kThreads / 32];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < kThreads / 32 ? red[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) {
      partials[blockIdx.x] = t;
      if (blockIdx.x == 0) hyper->step += 1;      // stream order: the update kernel of this step sees t = step
    }
  }
}

__global__ void __launch_bounds__(kThreads)
adamw_kernel(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m, float *__restrict__ v,
             __nv_bfloat16 *__restrict__ shadow, int64_t n_decay, int64_t n, const float *__restrict__ partials,
             int n_partials, Hyper *hyper) {
  __shared__ float s_coef, s_step_size, s_inv_sqrt_bc2, s_decay;
  __shared__ float red[kThreads / 32];
  {   // every CTA re-derives the clip coefficient from the per-CTA partial sums (fixed order: reproducible)
    float t = 0.f;
    for (int i = threadIdx.x; i < n_partials; i += kThreads) t += partials[i];
    t = warp_sum(t);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
      float tot = 0.f;
      for (int i = 0; i < kThreads / 32; ++i) tot += red[i];
      const float norm = sqrtf(tot);
      const Hyper h = *hyper;
      s_coef = h.max_norm > 0.f ? fminf(1.f, h.max_norm / (norm + 1e-6f)) : 1.f;
      const double bc1 = 1.0 - pow((double)h.beta1, (double)h.step), bc2 = 1.0 - pow((double)h.beta2, (double)h.step);
      s_step_size = (float)((double)h.lr / bc1);
      s_inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
      s_decay = 1.f - h.lr * h.weight_decay;
      if (blockIdx.x == 0) hyper->grad_norm = norm;
    }
    __syncthreads();
  }
  const Hyper h = *hyper;
  const float coef = s_coef, step_size = s_step_size, inv_sqrt_bc2 = s_inv_sqrt_bc2, decay = s_decay;
  const float b1 = h.beta1, b2 = h.beta2, eps = h.eps;
  auto upd = [&](float &pp, float gg, float &mm, float &vv, bool wd) {
    gg *= coef;
    if (wd) pp *= decay;
    mm = fmaf(b1, mm, (1.f - b1) * gg);
    vv = fmaf(b2, vv, (1.f - b2) * gg * gg);
    pp -= step_size * (mm / (sqrtf(vv) * inv_sqrt_bc2 + eps));
  };
  const int64_t nv = n >> 2;
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < nv; i += (int64_t)gridDim.x * kThreads) {
    float4 P = reinterpret_cast<float4 *>(p)[i], M = reinterpret_cast<float4 *>(m)[i], V = reinterpret_cast<float4 *>(v)[i];
    const float4 G = __ldg(reinterpret_cast<const float4 *>(g) + i);
    const int64_t e = i << 2;
    upd(P.x, G.x, M.x, V.x, e < n_decay);
    upd(P.y, G.y, M.y, V.y, e + 1 < n_decay);
    upd(P.z, G.z, M.z, V.z, e + 2 < n_decay);
    upd(P.w, G.w, M.w, V.w, e + 3 < n_decay);
    reinterpret_cast<float4 *>(p)[i] = P;
    reinterpret_cast<float4 *>(m)[i] = M;
    reinterpret_cast<float4 *>(v)[i] = V;
    if (shadow) {
      const __nv_bfloat162 lo = __floats2bfloat162_rn(P.x, P.y), hi = __floats2bfloat162_rn(P.z, P.w);
      reinterpret_cast<uint2 *>(shadow)[i] = make_uint2(*reinterpret_cast<const uint32_t *>(&lo), *reinterpret_cast<const uint32_t *>(&hi));
    }
  }
  if (blockIdx.x == 0)
    for (int64_t i = (nv << 2) + threadIdx.x; i < n; i += kThreads) {
      upd(p[i], g[i], m[i], v[i], i < n_decay);
      if (shadow) shadow[i] = __float2bfloat16_rn(p[i]);
    }
}

}  // namespace
}  // namespace mvit

extern "C" size_t mvit_adamw_workspace_floats(void) { return mvit::kMaxPartials; }
extern "C" size_t mvit_adamw_hyper_floats(void) { return sizeof(mvit::Hyper) / sizeof(float); }

extern "C" int mvit_adamw_clip_step(float *params, const float *grads, float *exp_avg, float *exp_avg_sq, void *bf16_shadow,
                                    int64_t n_decay, int64_t n, float *hyper, float *workspace, void *stream) {
  using namespace mvit;
  MVIT_REQUIRE(params && grads && exp_avg && exp_avg_sq && hyper && workspace, "adamw: null pointer");
  MVIT_REQUIRE(n >= 0 && n_decay >= 0 && n_decay <= n, "adamw: bad sizes");
  auto al = [](const void *q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  MVIT_REQUIRE(al(params) && al(grads) && al(exp_avg) && al(exp_avg_sq) && (!bf16_shadow || (reinterpret_cast<uintptr_t>(bf16_shadow) & 7) == 0),
               "adamw: arenas must be 16-byte aligned");
  if (n == 0) return 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int blocks = (int)std::min<int64_t>(kMaxPartials, std::min<int64_t>((n / 4 + kThreads - 1) / kThreads + 1, (int64_t)num_sms() * 4));
  grad_sqnorm_kernel<<<blocks, kThreads, 0, st>>>(grads, n, workspace, reinterpret_cast<Hyper *>(hyper));
  MVIT_LAUNCH_OK("grad_sqnorm");
  const int ublocks = (int)std::min<int64_t>((n / 4 + kThreads - 1) / kThreads + 1, (int64_t)num_sms() * 8);
  adamw_kernel<<<ublocks, kThreads, 0, st>>>(params, grads, exp_avg, exp_avg_sq, static_cast<__nv_bfloat16 *>(bf16_shadow),
                                             n_decay, n, workspace, blocks, reinterpret_cast<Hyper *>(hyper));
  MVIT_LAUNCH_OK("adamw");
  return 0;
}
