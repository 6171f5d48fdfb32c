"""Drop-in for the `MViT` model of slowfast/models/video_model_builder.py:794-1335 (plus PatchEmbed,
stem_helper.py:308-338; TransformerBasicHead, head_helper.py:369-417; round_width, models/utils.py:8-22).

Same cfg keys (MVIT.*, MODEL.*, DATA.*), same parameter names / shapes / registration order (so
`state_dict()` round-trips with reference checkpoints and a seeded construction draws the same
random numbers), same `forward(x: list[Tensor], ...)` contract.  The 16/24 blocks, the LayerNorms,
the positional-embedding add and the pooled head run in libmvit_b200.so; see attention.py.
"""
from __future__ import annotations

import math
import os
from functools import partial

import torch
import torch.nn as nn
import torch.utils.checkpoint
from torch.nn.init import trunc_normal_

from . import autograd as AG
from . import ops, weights
from .attention import MultiScaleBlock, _compute_dtype


def round_width(width, multiplier, min_width=1, divisor=1, verbose=False):
    """models/utils.py:8-22."""
    if not multiplier:
        return width
    width *= multiplier
    min_width = min_width or divisor
    width_out = max(min_width, int(width + divisor / 2) // divisor * divisor)
    if width_out < 0.9 * width:
        width_out += divisor
    return int(width_out)


class PatchEmbed(nn.Module):
    """Conv3d(3->96, k(3,7,7), s(2,4,4), p(1,3,3)) -> tokens [B, T'H'W', C] (stem_helper.py:308-338).

    The convolution is evaluated channels-last (NDHWC) so its output *is* the token tensor — the
    `flatten(2).transpose(1,2)` of the reference costs nothing."""

    def __init__(self, dim_in=3, dim_out=768, kernel=(1, 16, 16), stride=(1, 4, 4), padding=(1, 7, 7),
                 conv_2d=False):
        super().__init__()
        conv = nn.Conv2d if conv_2d else nn.Conv3d
        self.proj = conv(dim_in, dim_out, kernel_size=kernel, stride=stride, padding=padding)
        self.conv_2d = conv_2d

    def _gemm_weight(self, dtype):
        """Conv weight as a [Cout, Kp] GEMM operand (Kp = C*kt*kh*kw rounded up to 64, zero padded), cached."""
        w = self.proj.weight
        key = (dtype, w._version, w.device, w.data_ptr(), weights.generation())
        slot = getattr(self, "_b200_w", None)
        if slot is None or slot[0] != key:
            k = w[0].numel()
            kp = (k + 63) // 64 * 64
            wp = torch.zeros((w.shape[0], kp), dtype=dtype, device=w.device)
            wp[:, :k] = w.detach().reshape(w.shape[0], k).to(dtype)
            slot = (key, wp)
            self._b200_w = slot
        return slot[1]

    # ---- implicit-GEMM path (bf16): fold the stride into channels, then one tcgen05 GEMM over 5-D TMA tap boxes
    def _fold_geometry(self):
        k, s, p = list(self.proj.kernel_size), list(self.proj.stride), list(self.proj.padding)
        lo = [-((pp + ss - 1) // ss) for pp, ss in zip(p, s)]                 # first folded block a window touches
        hi = [(kk - 1 - pp) // ss for kk, pp, ss in zip(k, p, s)]
        taps = [h - l + 1 for h, l in zip(hi, lo)]
        creal = s[0] * s[1] * s[2] * self.proj.in_channels
        return k, s, p, lo, taps, creal, (creal + 63) // 64 * 64

    def _folded_weight(self):
        """Conv3d weight scattered into the folded layout: [Cout, taps_t*taps_h*taps_w*Cf] bf16 (cached)."""
        w = self.proj.weight
        key = (w._version, w.device, w.data_ptr(), weights.generation())
        slot = getattr(self, "_b200_wf", None)
        if slot is None or slot[0] != key:
            k, s, p, lo, taps, creal, cf = self._fold_geometry()
            C = self.proj.in_channels
            wf = torch.zeros((w.shape[0], taps[0], taps[1], taps[2], cf), dtype=torch.float32, device=w.device)
            wd = w.detach().float()
            for a in range(k[0]):
                bt, ot = divmod(a - p[0], s[0])                # folded block (relative to the output index), offset in block
                for b in range(k[1]):
                    bh, oh = divmod(b - p[1], s[1])
                    for d in range(k[2]):
                        bw, ow = divmod(d - p[2], s[2])
                        ch = ((ot * s[1] + oh) * s[2] + ow) * C
                        wf[:, bt - lo[0], bh - lo[1], bw - lo[2], ch:ch + C] = wd[:, :, a, b, d]
            slot = (key, wf.reshape(w.shape[0], -1).to(torch.bfloat16).contiguous())
            self._b200_wf = slot
        return slot[1]

    def _conv_path_ok(self, x, dtype):
        if self.conv_2d or dtype != torch.bfloat16:
            return False
        if x.dtype == torch.uint8:
            T, H, W = x.shape[1:4]
        else:
            T, H, W = x.shape[2:5]
        s = self.proj.stride
        if T % s[0] or H % s[1] or W % s[2]:
            return False
        k, _, p, _, _, _, _ = self._fold_geometry()
        out = [(n + 2 * pp - kk) // ss + 1 for n, kk, ss, pp in zip((T, H, W), k, s, p)]
        if out != [T // s[0], H // s[1], W // s[2]]:
            return False
        return out[0] % 2 == 0 and out[1] % 8 == 0 and out[2] % 8 == 0 and self.proj.out_channels % 8 == 0

    def forward(self, x, dtype=None, pos=None, pos_period=0, want_stats=False):
        """x: [B, C, T, H, W] clip (or [B, C, H, W] for conv_2d; or uint8 frames [B, T, H, W, C], normalised on the
        fly) -> tokens [B, L, Cout].  bf16: space-to-depth fold + implicit-GEMM convolution on tensor cores (5-D TMA,
        no im2col matrix); otherwise im2col + GEMM.  `pos` ([L, Cout] table in the activation dtype) and the bias are
        added in the GEMM epilogue."""
        if not x.is_cuda:
            raise ops._lib.MvitLibraryError("aicity_action_b200 runs on CUDA tensors only (no CPU fallback)")
        dtype = dtype or (torch.bfloat16 if x.dtype == torch.uint8 else x.dtype)
        if self._conv_path_ok(x, dtype):
            _, s, _, lo, taps, _, cf = self._fold_geometry()
            folded = ops.fold_clip(x, s, cf)
            return ops.patch_conv(folded, self._folded_weight(), self.proj.bias, pos, taps, lo, want_stats=want_stats)
        if x.dtype == torch.uint8:
            x = ops.preprocess_u8(x, dtype)
        t3 = lambda v: [1] + list(v) if self.conv_2d else list(v)
        if self.conv_2d:
            x = x.unsqueeze(2)
        kernel, stride, padding = t3(self.proj.kernel_size), t3(self.proj.stride), t3(self.proj.padding)
        if self.conv_2d:
            padding[0] = 0
        w = self._gemm_weight(dtype)
        patches, _ = ops.im2col3d(x.to(dtype), kernel, stride, padding, w.shape[1])
        y = ops.linear(patches, w, self.proj.bias, residual=pos, residual_row_period=pos_period if pos is not None else 0)
        return y.view(x.shape[0], -1, w.shape[0])


class TransformerBasicHead(nn.Module):
    """Dropout -> Linear -> (Softmax | Sigmoid when not training)  (head_helper.py:369-417)."""

    def __init__(self, dim_in, num_classes, dropout_rate=0.0, act_func="softmax", use_act_in_train=False):
        super().__init__()
        if dropout_rate > 0.0:
            self.dropout = nn.Dropout(dropout_rate)
        self.projection = nn.Linear(dim_in, num_classes, bias=True)
        self.use_act_in_train = use_act_in_train
        if act_func == "softmax":
            self.act = nn.Softmax(dim=1)
        elif act_func == "sigmoid":
            self.act = nn.Sigmoid()
        else:
            raise NotImplementedError(f"{act_func} is not supported as an activation function.")

    def forward_tokens(self, x):
        """x: [B, L, C] tokens; fused mean over L -> projection -> activation, fp32 result."""
        act = self.use_act_in_train or not self.training
        drop = hasattr(self, "dropout") and self.training and self.dropout.p > 0
        softmax = act and isinstance(self.act, nn.Softmax)
        if AG.recording(x, self.projection.weight, self.projection.bias):
            # differentiable path: a [B, C] mean and an 18-way projection — microseconds, left to autograd in fp32
            feat = x.float().mean(1)
            if drop:
                feat = self.dropout(feat)
            out = torch.nn.functional.linear(feat, self.projection.weight, self.projection.bias)
            return self.act(out) if act else out
        if not drop:
            out = ops.mean_head(x, self.projection.weight, self.projection.bias, softmax=softmax)
        else:
            _, feat = ops.mean_head(x, self.projection.weight, self.projection.bias, softmax=False, want_feat=True)
            feat = self.dropout(feat)
            out = ops.mean_head(feat.unsqueeze(1), self.projection.weight, self.projection.bias, softmax=softmax)
        if act and not softmax:
            out = self.act(out)
        return out

    def forward(self, x):
        return self.forward_tokens(x.unsqueeze(1) if x.ndim == 2 else x)


class MViT(nn.Module):
    """video_model_builder.py:794-1335 (classification path; RoI / multi-dataset / contrastive heads are
    out of scope and rejected at construction)."""

    def __init__(self, cfg):
        super().__init__()
        assert cfg.DATA.TRAIN_CROP_SIZE == cfg.DATA.TEST_CROP_SIZE
        self.cfg = cfg
        if cfg.DETECTION.ENABLE or cfg.MODEL.USE_MULTI_HEAD or cfg.CONTRA.ENABLE \
                or cfg.DETECTION.USE_SPATIAL_MAXPOOL_BEFORE_PROJ:
            raise NotImplementedError("aicity_action_b200.MViT implements the TransformerBasicHead path only")
        self.use_query_residual_pool = cfg.MVIT.Q_POOL_RESIDUAL
        self.q_pool_all = cfg.MVIT.Q_POOL_ALL
        self.channel_expand_front = cfg.MVIT.CHANNEL_EXPAND_FRONT
        self.pool_skip_use_conv = cfg.MVIT.POOL_SKIP_USE_CONV
        self.direct_input = cfg.MVIT.DIRECT_INPUT
        pool_first = cfg.MVIT.POOL_FIRST

        spatial_size = cfg.DATA.TRAIN_CROP_SIZE
        temporal_size = cfg.DATA.NUM_FRAMES
        in_chans = cfg.DATA.INPUT_CHANNEL_NUM[0]
        use_2d_patch = cfg.MVIT.PATCH_2D
        self.patch_stride = list(cfg.MVIT.PATCH_STRIDE)
        if use_2d_patch:
            self.patch_stride = [1] + self.patch_stride
        num_classes = cfg.MODEL.NUM_CLASSES
        embed_dim = cfg.MVIT.EMBED_DIM
        dim_out = embed_dim
        num_heads = cfg.MVIT.NUM_HEADS
        mlp_ratio = cfg.MVIT.MLP_RATIO
        qkv_bias = cfg.MVIT.QKV_BIAS
        self.drop_rate = cfg.MVIT.DROPOUT_RATE
        depth = cfg.MVIT.DEPTH
        drop_path_rate = cfg.MVIT.DROPPATH_RATE
        mode = cfg.MVIT.MODE
        self.cls_embed_on = cfg.MVIT.CLS_EMBED_ON
        self.sep_pos_embed = cfg.MVIT.SEP_POS_EMBED
        if cfg.MVIT.NORM != "layernorm":
            raise NotImplementedError("Only supports layernorm.")
        norm_layer = partial(nn.LayerNorm, eps=1e-6)
        self.num_classes = num_classes

        self.patch_embed = PatchEmbed(dim_in=in_chans, dim_out=embed_dim, kernel=cfg.MVIT.PATCH_KERNEL,
                                      stride=cfg.MVIT.PATCH_STRIDE, padding=cfg.MVIT.PATCH_PADDING,
                                      conv_2d=use_2d_patch)
        self.input_dims = [temporal_size, spatial_size, spatial_size]
        self.patch_dims = [self.input_dims[i] // self.patch_stride[i] for i in range(3)]
        num_patches = math.prod(self.patch_dims)

        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, depth)]

        if self.cls_embed_on:
            self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
            pos_embed_dim = num_patches + 1
        else:
            pos_embed_dim = num_patches
        if self.sep_pos_embed:
            self.pos_embed_spatial = nn.Parameter(torch.zeros(1, self.patch_dims[1] * self.patch_dims[2], embed_dim))
            self.pos_embed_temporal = nn.Parameter(torch.zeros(1, self.patch_dims[0], embed_dim))
            if self.cls_embed_on:
                self.pos_embed_class = nn.Parameter(torch.zeros(1, 1, embed_dim))
        else:
            self.pos_embed = nn.Parameter(torch.zeros(1, pos_embed_dim, embed_dim))
        if self.drop_rate > 0.0:
            self.pos_drop = nn.Dropout(p=self.drop_rate)

        # ---- per-block pooling schedule (video_model_builder.py:921-980) ----
        dim_mul, head_mul = [1.0] * (depth + 1), [1.0] * (depth + 1)
        for i, m in cfg.MVIT.DIM_MUL:
            dim_mul[i] = m
        for i, m in cfg.MVIT.HEAD_MUL:
            head_mul[i] = m
        pool_q = [[] for _ in range(depth)]
        pool_kv = [[] for _ in range(depth)]
        stride_q = [[] for _ in range(depth)]
        stride_kv = [[] for _ in range(depth)]
        kvq = cfg.MVIT.POOL_KVQ_KERNEL
        for ent in cfg.MVIT.POOL_Q_STRIDE:
            stride_q[ent[0]] = list(ent[1:])
            pool_q[ent[0]] = list(kvq) if kvq is not None else [s + 1 if s > 1 else s for s in ent[1:]]
        if self.q_pool_all:
            for i in range(depth):
                if not pool_q[i]:
                    pool_q[i] = list(kvq)
                    stride_q[i] = [1, 1, 1]
        if cfg.MVIT.POOL_KV_STRIDE_ADAPTIVE is not None:
            # the reference also writes the derived list back into cfg (video_model_builder.py:958-967)
            _stride_kv = list(cfg.MVIT.POOL_KV_STRIDE_ADAPTIVE)
            cfg.MVIT.POOL_KV_STRIDE = []
            for i in range(depth):
                if len(stride_q[i]) > 0:
                    _stride_kv = [max(_stride_kv[d] // stride_q[i][d], 1) for d in range(len(_stride_kv))]
                cfg.MVIT.POOL_KV_STRIDE.append([i] + _stride_kv)
        for ent in (cfg.MVIT.POOL_KV_STRIDE or []):
            stride_kv[ent[0]] = list(ent[1:])
            pool_kv[ent[0]] = list(kvq) if kvq is not None else [s + 1 if s > 1 else s for s in ent[1:]]

        self.norm_stem = norm_layer(embed_dim) if cfg.MVIT.NORM_STEM else None
        self.act_checkpoint = bool(cfg.MODEL.ACT_CHECKPOINT)
        # MODEL.ACT_CHECKPOINT trades a second forward for activation memory and is honoured as written ("always": every
        # block is recomputed in backward, as fairscale's checkpoint_wrapper does, video_model_builder.py:988-1036).  The
        # fused path never stores a score matrix, so a clip @448 keeps only ~1.6 GB of bf16 activations: on a 180 GB B200
        # MVIT_B200_ACT_CHECKPOINT=auto opts into recomputing only when the estimate does not fit the free memory (the
        # decision is logged once per input shape); =never ignores the flag.
        self.act_checkpoint_policy = os.environ.get("MVIT_B200_ACT_CHECKPOINT", "always")
        self._ckpt_decision = {}

        self.blocks = nn.ModuleList()
        rel_sp = bool(cfg.MVIT.get("REL_POS_SPATIAL", False))
        rel_tm = bool(cfg.MVIT.get("REL_POS_TEMPORAL", False))
        rel_zero = bool(cfg.MVIT.get("REL_POS_ZERO_INIT", False))
        blk_thw = list(self.patch_dims)                      # token grid entering block i (for the rel-pos table sizes)
        for i in range(depth):
            num_heads = round_width(num_heads, head_mul[i])
            if self.channel_expand_front:
                embed_dim = round_width(embed_dim, 1.0 if i == 0 else dim_mul[i - 1], divisor=num_heads)
                dim_out = round_width(dim_out, dim_mul[i], divisor=num_heads)
            else:
                embed_dim = round_width(embed_dim, dim_mul[i], divisor=num_heads)
                dim_out = round_width(embed_dim, dim_mul[i + 1], divisor=round_width(num_heads, head_mul[i + 1]))
            self.blocks.append(MultiScaleBlock(
                dim=embed_dim, dim_out=dim_out, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias,
                drop_rate=self.drop_rate, drop_path=dpr[i], norm_layer=norm_layer, kernel_q=pool_q[i],
                kernel_kv=pool_kv[i], stride_q=stride_q[i], stride_kv=stride_kv[i], mode=mode,
                has_cls_embed=self.cls_embed_on, pool_first=pool_first,
                use_query_residual_pool=self.use_query_residual_pool,
                channel_expand_front=self.channel_expand_front, pool_skip_use_conv=self.pool_skip_use_conv,
                # default-off extension (not in the reference): decomposed relative-position bias
                rel_pos_spatial=rel_sp, rel_pos_temporal=rel_tm, rel_pos_zero_init=rel_zero,
                input_size=list(blk_thw) if (rel_sp or rel_tm) else None))
            if len(stride_q[i]) > 0:
                blk_thw = [n // st for n, st in zip(blk_thw, stride_q[i])]
        embed_dim = dim_out
        self.norm = norm_layer(embed_dim) if not cfg.MVIT.NO_NORM_BEFORE_AVG else None

        if self.sep_pos_embed:
            trunc_normal_(self.pos_embed_spatial, std=0.02)
            trunc_normal_(self.pos_embed_temporal, std=0.02)
            if self.cls_embed_on:
                trunc_normal_(self.pos_embed_class, std=0.02)
        else:
            trunc_normal_(self.pos_embed, std=0.02)
        if self.cls_embed_on:
            trunc_normal_(self.cls_token, std=0.02)

        self.head = TransformerBasicHead(embed_dim, num_classes, dropout_rate=cfg.MODEL.DROPOUT_RATE,
                                         act_func=cfg.MODEL.HEAD_ACT,
                                         use_act_in_train=cfg.MODEL.USE_HEAD_ACT_IN_TRAIN)
        self.apply(self._init_weights)

    def _init_weights(self, m):
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, std=0.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    @torch.jit.ignore
    def no_weight_decay(self):
        if self.cfg.MVIT.ZERO_DECAY_POS_CLS:
            if self.sep_pos_embed:
                names = {"pos_embed_spatial", "pos_embed_temporal", "pos_embed_class"}
                return names | {"cls_token"} if self.cls_embed_on else names
            return {"pos_embed", "cls_token"} if self.cls_embed_on else {"pos_embed"}
        return {}

    # ------------------------------------------------------------------------------------
    def _pos_table(self, dtype):
        """Full [1+N or N, C] fp32 table for the generic (cls / non-separable) path."""
        if self.sep_pos_embed:
            T, H, W = self.patch_dims
            pos = self.pos_embed_spatial.repeat(1, T, 1) + torch.repeat_interleave(
                self.pos_embed_temporal, H * W, dim=1)
            if self.cls_embed_on:
                pos = torch.cat([self.pos_embed_class, pos], 1)
            return pos
        return self.pos_embed

    def _pos_tokens(self, dtype):
        """[N, C] separable positional-embedding table in the activation dtype (cached per parameter version)."""
        ps, pt = self.pos_embed_spatial, self.pos_embed_temporal
        key = (dtype, ps._version, pt._version, ps.device, ps.data_ptr(), pt.data_ptr(), weights.generation())
        slot = getattr(self, "_b200_pos", None)
        if slot is None or slot[0] != key:
            T, H, W = self.patch_dims
            tab = (ps.detach().repeat(1, T, 1) + torch.repeat_interleave(pt.detach(), H * W, dim=1))[0]
            slot = (key, tab.to(dtype).contiguous())
            self._b200_pos = slot
        return slot[1]

    def _checkpoint_needed(self, x):
        """`x`: tokens entering block 0.  Stored activations per token are ~14.5 block-widths of the compute dtype per
        block (LN inputs, qkv, pooled q/k/v, attention output, MLP hidden) summed over the blocks."""
        if self.act_checkpoint_policy == "always":
            return True
        if self.act_checkpoint_policy == "never":
            return False
        key = (tuple(x.shape), x.dtype, x.device)
        if key not in self._ckpt_decision:
            tokens, est = x.shape[0] * x.shape[1], 0
            thw = list(self.patch_dims)
            for blk in self.blocks:
                est += tokens * 14.5 * max(blk.dim, blk.dim_out) * x.element_size()
                new_thw = blk.out_thw(thw)
                tokens = tokens * (new_thw[0] * new_thw[1] * new_thw[2]) // (thw[0] * thw[1] * thw[2])
                thw = new_thw
            free, _ = torch.cuda.mem_get_info(x.device)
            self._ckpt_decision[key] = est * 1.5 > free      # keep activations when they fit with 50 % headroom
            import logging
            logging.getLogger(__name__).info(
                "MVIT_B200_ACT_CHECKPOINT=auto: estimated %.1f GB of activations, %.1f GB free -> %s", est / 2 ** 30,
                free / 2 ** 30, "recompute blocks in backward" if self._ckpt_decision[key] else "keep activations")
        return self._ckpt_decision[key]

    def forward_features(self, x, dtype):
        T, H, W = self.patch_dims
        stem_params = [self.patch_embed.proj.weight, self.patch_embed.proj.bias]
        stem_params += [self.pos_embed_spatial, self.pos_embed_temporal] if self.sep_pos_embed else [self.pos_embed]
        if AG.recording(*stem_params, getattr(self, "cls_token", None)):
            if self.sep_pos_embed and not self.cls_embed_on:
                x = AG.patch_embed(self.patch_embed, x, dtype, self.pos_embed_spatial, self.pos_embed_temporal,
                                   self._pos_tokens(dtype))
            else:
                # cls token / non-separable pos-embed (secondary variants): patch GEMM gradients through the same
                # Function, concatenation and the positional add left to autograd
                tokens = AG.patch_embed(self.patch_embed, x, dtype)
                if self.cls_embed_on:
                    tokens = torch.cat((self.cls_token.to(tokens.dtype).expand(tokens.shape[0], -1, -1), tokens), dim=1)
                x = tokens + self._pos_table(dtype).to(tokens.dtype)
        elif self.sep_pos_embed and not self.cls_embed_on:
            # bias + positional embedding ride in the patch-embed GEMM epilogue
            pos = self._pos_tokens(dtype)
            # (eval / bf16: the epilogue also leaves the tokens' row statistics for block 0's folded norm1)
            x = self.patch_embed(x, dtype, pos=pos, pos_period=pos.shape[0],
                                 want_stats=ops.ln_fold_enabled() and self.norm_stem is None
                                 and not (self.drop_rate and self.training))
        else:
            tokens = self.patch_embed(x, dtype)                  # [B, N, C]
            B = tokens.shape[0]
            if self.cls_embed_on:
                tokens = torch.cat((self.cls_token.to(tokens.dtype).expand(B, -1, -1), tokens), dim=1)
            pos = self._pos_table(dtype).detach()
            zero_t = torch.zeros((1, tokens.shape[-1]), dtype=torch.float32, device=tokens.device)
            x = ops.pos_embed_add(tokens, pos, zero_t, 1, dtype)
        if self.drop_rate and self.training:
            x = self.pos_drop(x)
        if self.norm_stem is not None:
            x = AG.layernorm(x, self.norm_stem.weight, self.norm_stem.bias, self.norm_stem.eps)
        thw = [T, H, W]
        use_ckpt = self.act_checkpoint and x.requires_grad and torch.is_grad_enabled() and self._checkpoint_needed(x)
        for blk in self.blocks:
            if use_ckpt:
                # MODEL.ACT_CHECKPOINT (video_model_builder.py:988-1036 wraps every block in fairscale's checkpoint_wrapper)
                thw_in = list(thw)
                x = torch.utils.checkpoint.checkpoint(lambda t, b=blk, s=thw_in: b(t, s)[0], x, use_reentrant=False)
                thw = blk.out_thw(thw_in)
            else:
                x, thw = blk(x, thw)
        if self.norm is not None:
            x = AG.layernorm(x, self.norm.weight, self.norm.bias, self.norm.eps)
        return x, thw

    def forward(self, x, bboxes=None, dataset_name=None, run_cross_proj=False, use_moco=False,
                moco_momentum=0.9):
        if not self.direct_input:
            x = x[0]
        dtype = _compute_dtype(x)
        x, _ = self.forward_features(x, dtype)
        if self.cls_embed_on:
            x = x[:, :1]
        return self.head.forward_tokens(x)
