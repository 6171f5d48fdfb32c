/*
 * mvit_b200.h — C ABI of libmvit_b200.so: sm_100a kernels for the MViTv2 multiscale-attention path.
 *
 * The reference (JunweiLiang/aicity_action, a PySlowFast fork) has no native layer: its plug
 * point is the Python module API of slowfast/models/attention.py.  This ABI is what the Python
 * drop-in (aicity_action_b200/attention.py, mvit.py) binds with ctypes; each entry point names the
 * reference code it replaces (file:line under /root/reference).
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless its name ends in _host; the caller owns all memory,
 *    the library never allocates or frees device memory and keeps no mutable global state;
 *  - all calls are asynchronous on `stream` (a cudaStream_t passed as void*), re-entrant, and safe
 *    under CUDA-graph capture (no synchronisation, no allocation);
 *  - return value 0 = launched, negative = rejected before launch (bad shape / unsupported
 *    combination / CUDA error); mvit_last_error() returns a thread-local message.  There is no CPU
 *    fallback and no silent fallback of any kind: an unsupported request is an error;
 *  - `dtype` selects the activation type: MVIT_F32 (fp32 storage, fp32 FMA arithmetic) or MVIT_BF16
 *    (bf16 storage, fp32 accumulation; GEMM/attention run on tcgen05 tensor cores);
 *  - LayerNorm gamma/beta, biases and depthwise-conv weights are always fp32; Linear weights have
 *    the activation dtype;
 *  - token tensors are channels-last: [B, L, C] with L = T*H*W (+1 leading cls token if has_cls).
 */
#ifndef MVIT_B200_H_
#define MVIT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MVIT_F32 0
#define MVIT_BF16 1

/* pooling modes of attention_pool (attention.py:136-212) */
#define MVIT_POOL_CONV 0 /* depthwise Conv3d, zero padding, no bias */
#define MVIT_POOL_MAX 1  /* MaxPool3d, -inf padding              */
#define MVIT_POOL_AVG 2  /* AvgPool3d, count_include_pad=True     */

/* epilogues of mvit_linear_fwd */
#define MVIT_EPI_NONE 0
#define MVIT_EPI_GELU 1 /* exact erf GELU (common.py:20 nn.GELU) */
#define MVIT_EPI_GELU_GRAD 2 /* backward helper: y = residual * gelu'(x.w^T + bias)  (residual = upstream gradient) */

/* implementation selector of the GEMM / attention entry points */
#define MVIT_IMPL_AUTO 0    /* bf16 -> tcgen05, f32 -> fp32 FMA kernels */
#define MVIT_IMPL_SIMT 1    /* force the CUDA-core kernels (any dtype)  */
#define MVIT_IMPL_TCGEN05 2 /* force tensor-core kernels (bf16 only)    */

/* Library / build identification. */
int mvit_abi_version(void);
const char *mvit_last_error(void);
/* 1 when the current device is sm_100 (B200); 0 otherwise; negative on CUDA error. */
int mvit_device_supported(void);
/* Number of tensor-core kernel families that abandoned an mbarrier wait since the last call (0 = healthy).  The
 * kernels never trap (a trap would poison the CUDA context): a wait that exceeds ~2^26 polls raises a device flag,
 * the kernel drains, and this call reads and clears the flags.  It SYNCHRONISES the device: call it where the host
 * synchronises anyway (after a step / before trusting results), never inside a captured region. */
int mvit_device_fault(void);

/*
 * LayerNorm over the last dimension.   y[r,:] = (x[r,:]-mean)/sqrt(var+eps)*gamma+beta
 * Replaces: norm1 / norm2 of MultiScaleBlock (attention.py:421,436; eps 1e-6) and MViT.norm
 * (video_model_builder.py:1249).  Statistics in fp32, biased variance (torch semantics).
 */
int mvit_layernorm_fwd(const void *x, const float *gamma, const float *beta, void *y,
                       int64_t rows, int channels, float eps, int dtype, void *stream);

/*
 * Linear with fused epilogue.   y[M,N] = epi(x[M,K] · w[N,K]^T + bias[N]) (+ residual[M,N])
 * Replaces: qkv / proj (attention.py:231,281), proj_max_pool / proj (attention.py:426,443),
 * Mlp.fc1+GELU / fc2 (common.py:27-31) and the residual adds of attention.py:434,445, which are
 * folded into the producing GEMM as `residual` (+ per-sample DropPath scale `row_scale`).
 *   bias, residual, row_scale may be NULL.  row_scale[b] multiplies (x·wᵀ+bias) for rows
 *   [b*rows_per_sample, (b+1)*rows_per_sample) BEFORE the residual is added (common.py:46-59).
 *   ldy / ldr: leading dimensions (elements) of y and residual (>= N).
 *   residual_row_period: 0 = residual has M rows; P > 0 = residual has P rows and row m reads row m % P
 *   (the positional-embedding table broadcast over the batch, video_model_builder.py:1206-1223).
 */
int mvit_linear_fwd(const void *x, const void *w, const float *bias, const void *residual,
                    const float *row_scale, int64_t rows_per_sample, void *y, int64_t M, int N,
                    int K, int64_t ldy, int64_t ldr, int64_t residual_row_period, int epilogue,
                    int dtype, int impl, void *stream);

/*
 * LayerNorm folded around the Linear layers of a block (eval path, bf16, tcgen05 only; there is no other implementation).
 * Replaces: norm1 -> attn.qkv (attention.py:421,231) and norm2 -> mlp.fc1 (attention.py:436, common.py:27) without a
 * LayerNorm kernel and without materialising the normalised tensor.  Exactly one of {stats_out, ln_stats} is set:
 *   producer form (stats_out != NULL): y = (x.w^T + bias)*row_scale + residual as mvit_linear_fwd, and additionally
 *     stats_out[p][m] = (sum_n, sum_n^2) over the columns of N tile p of the bf16 values stored in row m;
 *     stats_out: [mvit_linear_stat_parts(M, N, K)][M][2] fp32.
 *   consumer form (ln_stats != NULL): x is the RAW block stream, ln_stats[ln_parts][M][2] its row statistics from a
 *     producer, w = W.diag(gamma) (bf16), colsum[n] = sum_k w[n,k] (fp32, of the bf16 values), bias = b + W.beta:
 *     y = epi(rstd_m * (x.w^T - mean_m * colsum) + bias),  mean/var over K, biased variance, rstd = 1/sqrt(var + ln_eps)
 *     — algebraically LayerNorm(x).W^T + b.  epilogue: MVIT_EPI_NONE or MVIT_EPI_GELU.
 */
int mvit_linear_stat_parts(int64_t M, int N, int K);
int mvit_linear_ln_fwd(const void *x, const void *w, const float *bias, const float *colsum, const float *ln_stats,
                       int ln_parts, float ln_eps, const void *residual, const float *row_scale,
                       int64_t rows_per_sample, void *y, float *stats_out, int64_t M, int N, int K, int64_t ldy,
                       int64_t ldr, int epilogue, void *stream);

/*
 * Fused MLP of a block for the front stages (eval path, bf16, tcgen05 only):
 *   y = x + fc2(GELU(fc1(LayerNorm(x))))        attention.py:436-445 + common.py:26-33 with norm2 folded as in mvit_linear_ln_fwd
 * x: RAW block stream [M, C] (also the residual), ln_stats its row statistics [ln_parts][M][2]; w1f = fc1.weight*diag(gamma)
 * [4C, C] bf16, b1f = fc1.bias + fc1.weight.beta, colsum1[n] = sum_k w1f[n,k]; w2 = fc2.weight [C, 4C] bf16, b2 = fc2.bias.
 * stats_out (optional) [1][M][2]: row statistics of y for the next block's folded norm1.  The 4C-wide hidden activation stays
 * in tensor / shared memory.  Supported: C in {96, 192}, hidden = 4C, output width C (mvit_mlp_fused_supported).
 */
int mvit_mlp_fused_supported(int C, int H, int C_out);
int mvit_mlp_fused_fwd(const void *x, const float *ln_stats, int ln_parts, float ln_eps, const void *w1f, const float *b1f,
                       const float *colsum1, const void *w2, const float *b2, void *y, float *stats_out, int64_t M, int C,
                       int H, void *stream);

/*
 * im2col for the patch-embedding Conv3d (stem_helper.py:308-338): clip [B, C, T, H, W] (channels-first,
 * as the reference feeds it) -> patch matrix [B*To*Ho*Wo, Kp] with column k = ((c*kt + a)*kh + b)*kw + d
 * (the natural flattening of the Conv3d weight [Cout, C, kt, kh, kw]); zero padding outside the clip and
 * in columns [C*kt*kh*kw, Kp).  The convolution itself is then mvit_linear_fwd on tensor cores with the
 * bias and the positional embedding fused in its epilogue.
 */
int mvit_im2col3d_fwd(const void *clip, void *patches, int B, int C, int T, int H, int W, int kt, int kh,
                      int kw, int st, int sh, int sw, int pt, int ph, int pw, int Kp, int dtype,
                      void *stream);

/*
 * attention_pool (attention.py:12-83) fused with its LayerNorm (attention.py:66-67).
 *
 * Input: a strided view of tokens  in[b, l, head, c]  at  in + b*in_bs + l*in_ls + head*in_hs + c
 * (element strides).  This reads q, k or v straight out of the qkv Linear output
 * [B, N, 3, heads, d] (attention.py:231-236) without the channels-first copy of attention.py:34-36,
 * and equally a plain [B, L, C] tensor (heads = C/d, in_hs = d) for the skip path (attention.py:427).
 * Output: out[b, l', head, c] at out + b*out_bs + l'*out_ls + head*out_hs + c.
 *   mode      MVIT_POOL_*; `weight` = [d, kt*kh*kw] fp32 for CONV (Conv3d weight [d,1,kt,kh,kw]), else NULL
 *   gamma/beta/eps  LayerNorm over d after pooling (NULL = none; eps 1e-5 on this path)
 *   has_cls   first token bypasses pooling and joins the LayerNorm (attention.py:28-29, 62-64)
 *   padding is k/2 per axis, ceil_mode = False:  T' = (T + 2*(kt/2) - kt)/st + 1, etc.
 */
int mvit_attention_pool_fwd(const void *in, int64_t in_bs, int64_t in_ls, int64_t in_hs,
                            const float *weight, const float *gamma, const float *beta, void *out,
                            int64_t out_bs, int64_t out_ls, int64_t out_hs, int B, int heads, int d,
                            int T, int H, int W, int kt, int kh, int kw, int st, int sh, int sw,
                            int mode, int has_cls, float eps, int dtype, void *stream);

/*
 * Pooling attention core (attention.py:267-279):
 *   out[b, i, head, :] = softmax_j( q[b,head,i,:]·k[b,head,j,:] * scale ) · v[b,head,j,:]  (+ q[b,head,i,:])
 * q: [B, heads, Lq, d], k/v: [B, heads, Lk, d] contiguous; out: [B, Lq, heads*d] (the head-merge copy of
 * attention.py:276 is folded into the store).  d must be 96.  The [Lq, Lk] score matrix is never
 * written to memory.  lse (optional, fp32 [B, heads, Lq]) receives log-sum-exp of the scaled scores.
 */
int mvit_attention_fwd(const void *q, const void *k, const void *v, void *out, float *lse, int B,
                       int heads, int Lq, int Lk, int d, float scale, int add_q_residual, int dtype,
                       int impl, void *stream);

/*
 * Decomposed relative-position bias (DEFAULT-OFF; north_star item 2).  NOT part of the reference: it follows
 * upstream PySlowFast's cal_rel_pos_spatial / cal_rel_pos_temporal as restated in SURVEY.md Appendix F and is validated
 * only against the in-repo restatement (oracle/mvit_oracle.py rel_pos_bias) - parity unpinned.
 *   scores[i,j] = scale * q_i.k_j + q_i.Rh[h_i,h'_j] + q_i.Rw[w_i,w'_j] + q_i.Rt[t_i,t'_j],  R*[a,b] = rel_pos_*[dist(a,b)],
 *   dist(a,b) = trunc(a*max(kn/qn,1) - b*max(qn/kn,1) + (kn-1)*max(qn/kn,1))
 * mvit_relpos_operands_fwd turns q [B*heads, qt*qh*qw, 96] (pooled, normalised, UNscaled) and the fp32 tables
 * rel_pos_h [2*max(qh,kh)-1, 96], rel_pos_w, rel_pos_t into two 64-column operands in `dtype`:
 *   q_ext [B*heads, Lq, 64] = [A_h | A_w | A_t | 0] / scale   (A_h[i, h'] = q_i.Rh[h_i, h'] ...)
 *   k_ext [B*heads, Lk, 64] = one-hot(h'_j) | one-hot(w'_j) | one-hot(t'_j) | 0        (kt + kh + kw <= 64)
 * and mvit_attention_rel_fwd is mvit_attention_fwd with the contraction extended over them, scores = scale*(q.k + q_ext.k_ext):
 * the bias rides in the same tensor-core GEMM (10 instead of 6 K16 steps), nothing is added in the softmax and no [Lq, Lk]
 * tensor exists.  Tokens are ordered (t, h, w), w fastest; no cls token.
 */
int mvit_relpos_operands_fwd(const void *q, const float *rel_h, const float *rel_w, const float *rel_t, void *q_ext,
                             void *k_ext, int BH, int qt, int qh, int qw, int kt, int kh, int kw, float scale, int dtype,
                             void *stream);
int mvit_attention_rel_fwd(const void *q, const void *k, const void *v, const void *q_ext, const void *k_ext, void *out,
                           float *lse, int B, int heads, int Lq, int Lk, int d, float scale, int add_q_residual, int dtype,
                           int impl, void *stream);

/*
 * Separable positional embedding add (video_model_builder.py:1206-1223):
 *   x[b, (t*HW + s), c] += pos_spatial[s, c] + pos_temporal[t, c]     (in place, fp32 tables)
 * with an optional dtype conversion: `src` (fp32 or bf16 per src_dtype) -> `dst` (dtype).
 */
int mvit_pos_embed_add(const void *src, int src_dtype, const float *pos_spatial,
                       const float *pos_temporal, void *dst, int B, int T, int HW, int C, int dtype,
                       void *stream);

/*
 * Token mean-pool + classification head (video_model_builder.py:1310-1314, head_helper.py:409-417):
 *   feat[b,:] = mean_l x[b,l,:];  logits = feat·Wᵀ + bias;  probs = softmax(logits) when apply_softmax.
 * x: [B, L, C] (dtype); w: [num_classes, C] fp32; outputs fp32.  feat_out may be NULL.
 * workspace: optional fp32 scratch of mvit_mean_head_workspace_floats(B, L, C) elements; with it the
 * token sum runs as a grid-wide two-stage reduction (fixed order, no atomics) instead of one CTA per clip.
 */
size_t mvit_mean_head_workspace_floats(int B, int L, int C);
int mvit_mean_head_fwd(const void *x, const float *w, const float *bias, float *feat_out,
                       float *out, float *workspace, int B, int L, int C, int num_classes,
                       int apply_softmax, int dtype, void *stream);

/*
 * Clip normalisation (scripts/module_wrapper.py:326-346): uint8 frames [B, T, H, W, 3] (RGB, already
 * resized) -> float clip [B, 3, T, H, W]:  ((x / 255) - mean) / std  evaluated in fp32 with IEEE
 * division and rounded once to `dtype`, i.e. bit-identical to the reference's NumPy float32 pipeline
 * when dtype == MVIT_F32.  Lets the host upload 1 byte per sample instead of 4.
 */
int mvit_preprocess_u8_fwd(const uint8_t *frames, void *clip, int B, int T, int H, int W, float mean,
                           float std, int dtype, void *stream);

/*
 * The three conv poolings of one attention block in one call (attention.py:172-212 pool_q / pool_k / pool_v + their
 * LayerNorms), reading the fused qkv GEMM output IN PLACE: qkv is [B, T*H*W, 3, heads, 96] bf16 contiguous.
 * weights[i] / gammas[i] / betas[i] (i = 0 q, 1 k, 2 v): fp32 [96, 27] depthwise 3x3x3 filter, LayerNorm scale / shift
 * (both NULL: no LayerNorm).  strides_hw[i] = s for stride (1, s, s), s in {1, 2, 4, 8}; padding 1; s = 0 skips tensor i
 * (a caller may pool q on one stream and k, v on another).
 * outs[i]: [B, heads, T*H'*W', 96] bf16 contiguous, H' = (H - 1) / s + 1.  pre_outs (may be NULL, or hold NULL entries):
 * the conv output before the LayerNorm, same layout (saved for the LayerNorm backward in training).
 * Tensors of stride 1 / 2 run on a persistent TMA-fed kernel (equal strides share one launch); strides 4 / 8 on the
 * cp.async kernel that gathers only the touched columns.  All pointer arrays are HOST arrays of device pointers.
 */
int mvit_attention_pool_qkv_fwd(const void *qkv, int B, int heads, int T, int H, int W, const float *const *weights,
                                const float *const *gammas, const float *const *betas, const int *strides_hw,
                                void *const *outs, void *const *pre_outs, float eps, int dtype, void *stream);

/*
 * Training-step glue (tools/train_net.py:229-246: zero_grad / clip_grad_norm_ / AdamW.step; slowfast/models/
 * optimizer.py:200-206): fused AdamW + global-norm gradient clip over FLAT fp32 arenas laid out
 * [decayed parameters (n_decay elements) | non-decayed], torch.optim.AdamW semantics (decoupled weight decay,
 * bias-corrected moments).  Two launches: the squared-norm reduction (which also advances the step counter) and the
 * update, which folds min(1, max_norm / (||g|| + 1e-6)) into the gradient and, when bf16_shadow != NULL, also writes the
 * bf16 copy of every updated parameter (the tensor-core operand of the next forward).
 * hyper: DEVICE float[mvit_adamw_hyper_floats()] = {lr, beta1, beta2, eps, weight_decay, max_norm (<= 0: no clip),
 * int32 step (device-owned, start at 0), grad_norm (written: pre-clip norm of this step)} — read from memory at run
 * time so that a captured CUDA graph of the step replays with a new learning rate.
 * workspace: DEVICE float[mvit_adamw_workspace_floats()].  All arenas 16-byte aligned.
 */
size_t mvit_adamw_workspace_floats(void);
size_t mvit_adamw_hyper_floats(void);
int mvit_adamw_clip_step(float *params, const float *grads, float *exp_avg, float *exp_avg_sq, void *bf16_shadow,
                         int64_t n_decay, int64_t n, float *hyper, float *workspace, void *stream);

/*
 * Frame gather + uint8 bilinear resize, bit-exact with OpenCV `cv2.resize(frame_u8, (out_w, out_h), INTER_LINEAR)`
 * (the reference resizes every uint8 frame of a window before the cast to float: scripts/utils.py:207-211 through
 * scripts/module_wrapper.py:304-331, keep_scale=False).  src: [n_src, H, W, channels] interleaved uint8 frames on the
 * device; frame_idx: device int32[n_out], output frame i is made from source frame frame_idx[i] (NULL: identity,
 * n_out == n_src) — this is the per-window frame-index gather of module_wrapper.py:384-397 done on the device;
 * dst: [n_out, out_h, out_w, channels] uint8.  Fixed-point arithmetic of OpenCV's HResizeLinear / VResizeLinear
 * (11-bit coefficients); see csrc/resize.cu.  channels: 3 or 1.
 */
int mvit_resize_gather_u8(const uint8_t *src, int n_src, int H, int W, const int32_t *frame_idx, int n_out,
                          uint8_t *dst, int out_h, int out_w, int channels, void *stream);

/*
 * Patch embedding as an implicit GEMM (stem_helper.py:308-338 Conv3d + video_model_builder.py:1206-1223 pos-embed).
 *  1. mvit_fold_clip_fwd: space-to-depth of the clip by the conv stride,
 *       folded[b, t/st, h/sh, w/sw, ((t%st*sh + h%sh)*sw + w%sw)*C + c] = x[b, c, t, h, w]      (bf16, zero padded to Cf)
 *     src_kind 0 / 1: fp32 / bf16 channels-first clip [B, C, T, H, W] (what the reference feeds the model);
 *     src_kind 2: uint8 channels-last frames [B, T, H, W, C] with the reference normalisation (x/255 - mean)/std fused.
 *  2. mvit_patch_conv_fwd: out[b, t, h, w, :] = sum_taps folded[b, t+dt, h+dh, w+dw, :] . wf[:, tap, :] + bias + pos[t,h,w,:]
 *     on tcgen05 tensor cores; every tap's operand tile is fetched by a 5-D TMA box shifted by the tap offset
 *     (zero fill outside the clip) - no im2col matrix exists.  wf: [N, nt*nh*nw*Cf] bf16 (the Conv3d weight scattered
 *     into the folded layout by the host), taps dt in [lo_t, lo_t+nt) etc.; pos: [Tf, Hf, Wf, N] bf16 or NULL.
 *     Requires Tf % 2 == 0, Hf % 8 == 0, Wf % 8 == 0, Cf % 64 == 0.
 */
int mvit_fold_clip_fwd(const void *clip, int src_kind, void *folded, int B, int C, int T, int H, int W, int st,
                       int sh, int sw, int Cf, float mean, float std, void *stream);
int mvit_patch_conv_fwd(const void *folded, const void *wf, const float *bias, const void *pos, void *out, int B,
                        int Tf, int Hf, int Wf, int Cf, int nt, int nh, int nw, int lo_t, int lo_h, int lo_w, int N,
                        void *stream);

/* mvit_patch_conv_fwd that also emits the row statistics of its output tokens (producer form of mvit_linear_ln_fwd):
 * stats_out [ceil(N/96)][B*Tf*Hf*Wf][2] fp32, indexed by token. */
int mvit_patch_conv_stats_fwd(const void *folded, const void *wf, const float *bias, const void *pos, void *out,
                              float *stats_out, int B, int Tf, int Hf, int Wf, int Cf, int nt, int nh, int nw, int lo_t,
                              int lo_h, int lo_w, int N, void *stream);

/* Training form of mvit_attention_pool_fwd (conv mode + LayerNorm, no cls token): additionally writes the pooled values
 * BEFORE the LayerNorm to pre_ln_out (contiguous [B, heads, L', d], dtype), which the LayerNorm backward needs — saving
 * them costs one extra store, recomputing them costs the whole convolution.  Only the tuned kernel produces it
 * (3x3x3 depthwise conv, head_dim 96, stride (1,s,s) with s in {1,2,4,8}, 16-byte aligned strides); otherwise an error
 * is returned and the caller recomputes with mvit_attention_pool_fwd(gamma = NULL). */
int mvit_attention_pool_fwd_save(const void *in, int64_t in_bs, int64_t in_ls, int64_t in_hs,
                                 const float *weight, const float *gamma, const float *beta, void *out,
                                 int64_t out_bs, int64_t out_ls, int64_t out_hs, void *pre_ln_out, int B, int heads,
                                 int d, int T, int H, int W, int kt, int kh, int kw, int st, int sh, int sw,
                                 float eps, int dtype, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Backward entry points (the reference has no explicit backward code: tools/train_net.py:229-246 calls
 * loss.backward() and autograd differentiates attention.py / common.py op by op; these are those
 * derivatives, one launch per fused forward op).  Activations / activation gradients are `dtype`;
 * parameter gradients are fp32 and ACCUMULATED (+=) into caller-zeroed buffers.
 * ------------------------------------------------------------------------------------------------ */

/* LayerNorm backward (nn.LayerNorm at attention.py:421,436,66-67): x, dy, dx [rows, channels]; dgamma/dbeta += */
int mvit_layernorm_bwd(const void *x, const float *gamma, const void *dy, void *dx, float *dgamma,
                       float *dbeta, int64_t rows, int channels, float eps, int dtype, void *stream);

/* exact-erf GELU backward (common.py:20): dpre = dy * gelu'(pre), n elements */
int mvit_gelu_bwd(const void *pre, const void *dy, void *dpre, int64_t n, int dtype, void *stream);

/* nn.Linear parameter gradients: dw[N,K] += dy[M,N]^T x[M,K];  db[N] += column sums of dy (db may be NULL).
 * (the input gradient dx = dy . W is mvit_linear_fwd with the transposed weight).  bf16 runs on tcgen05 with both
 * operands consumed MN-major straight from the row-major activations (no transposed copies), split over tokens. */
int mvit_linear_wgrad(const void *dy, const void *x, float *dw, float *db, int64_t M, int N, int K, int dtype,
                      int impl, void *stream);

/* Backward of mvit_attention_fwd (attention.py:267-279).  q/k/v/out as in the forward, dout [B, Lq, heads*d],
 * lse from the forward; dq [B, heads, Lq, d] (dtype, overwritten; includes the +q residual path),
 * dk/dv [B, heads, Lk, d] fp32, caller-zeroed, accumulated.  workspace: fp32 scratch of
 * mvit_attention_bwd_workspace_floats(B, heads, Lq) elements (16-byte aligned) for the tcgen05 path
 * (bf16: three launches - row statistics, dQ, dK/dV - no [Lq, Lk] matrix in memory); NULL selects CUDA cores. */
size_t mvit_attention_bwd_workspace_floats(int B, int heads, int Lq);
int mvit_attention_bwd(const void *q, const void *k, const void *v, const void *out, const void *dout,
                       const float *lse, void *dq, float *dk, float *dv, float *workspace, int B, int heads,
                       int Lq, int Lk, int d, float scale, int add_q_residual, int dtype, int impl, void *stream);

/* Backward pieces of mvit_attention_pool_fwd (attention.py:12-83), geometry arguments as in the forward
 * (padding = kernel/2, no cls token):
 *   what 0: depthwise-conv input gradient, dx written through the forward's input strides (x_bs, x_ls, x_hs);
 *   what 1: depthwise-conv weight gradient dw[d, taps] += (x read through the strides);
 *   what 2: max-pool input gradient into a caller-zeroed dense fp32 dx [B, L, heads, d] (arg-max recomputed from x);
 *   what 3: the same with dy laid out [B, heads, L', d] (max-pooled q / k / v of MVIT.MODE "max").
 * dy is the gradient w.r.t. the pooled (pre-LayerNorm) output: [B, heads, L', d] for what 0/1, token layout
 * [B, L', heads*d] for what 2 (the block's skip path pools [B, L, C] tokens). */
int mvit_attention_pool_bwd(int what, const void *x, int64_t x_bs, int64_t x_ls, int64_t x_hs, const void *dy,
                            const float *weight, void *dx, float *dw, int B, int heads, int d, int T, int H,
                            int W, int kt, int kh, int kw, int st, int sh, int sw, int dtype, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* MVIT_B200_H_ */
