"""Drop-in for slowfast/models/attention.py: attention_pool, MultiScaleAttention, MultiScaleBlock.

Same constructor signatures, attribute names, parameter names / shapes / registration order and
`forward(x, thw_shape) -> (tensor, thw)` contract as the reference (attention.py:12, 86, 287), so
`state_dict()` interchanges with reference checkpoints.  The nn.Linear /
nn.Conv3d / nn.LayerNorm sub-modules are *parameter holders only*: all arithmetic is issued to the
sm_100a kernels of libmvit_b200.so through `ops` (no torch math on the hot path, no CPU fallback).

Dataflow of one block (B200 form of SURVEY.md §3.3; ✚ marks fusions that remove reference kernels):
    LN(1e-6) -> qkv GEMM(+bias) -> pool_{q,k,v}: depthwise conv3d ✚ LN(1e-5), read straight from the
    [B,N,3,h,96] GEMM output (no channels-first copies) -> fused attention softmax(qkᵀ/√d)v ✚ +q ✚
    head-merge -> proj GEMM ✚ bias ✚ skip-residual ✚ DropPath scale -> LN -> fc1 GEMM ✚ bias ✚ GELU ->
    fc2 GEMM ✚ bias ✚ residual ✚ DropPath scale.  Identity MaxPool3d([1,1,1]) skips are elided (D6).
"""
from __future__ import annotations

import numpy
import torch
import torch.nn as nn

from . import autograd as AG
from . import ops
from . import weights
from .common import DropPath, Mlp, drop_path_scale


_COMPUTE_POLICY = None       # None: read $MVIT_B200_COMPUTE at call time; else "auto" | "bf16" | "fp32"


def set_compute_dtype(policy):
    """Arithmetic used for fp32 input tensors: "auto" (fp32 kernels unless a 16-bit autocast region is active — the
    reference's own contract), "bf16" (bf16 storage + tcgen05 kernels, as if the call sat in a bf16 autocast region; what
    the reference's inference scripts need to reach the tensor cores: they pass fp32 clips with no autocast,
    scripts/module_wrapper.py:606-608), "fp32" (fp32 kernels even under autocast).  None re-enables $MVIT_B200_COMPUTE."""
    global _COMPUTE_POLICY
    if policy not in (None, "auto", "bf16", "fp32"):
        raise ValueError(f"compute dtype policy must be auto, bf16 or fp32, got {policy!r}")
    _COMPUTE_POLICY = policy


def compute_policy() -> str:
    import os
    pol = _COMPUTE_POLICY or os.environ.get("MVIT_B200_COMPUTE", "auto").lower()
    if pol not in ("auto", "bf16", "fp32"):
        raise ValueError(f"MVIT_B200_COMPUTE must be auto, bf16 or fp32, got {pol!r}")
    return pol


def _compute_dtype(x: torch.Tensor) -> torch.dtype:
    """fp32 tensors run the fp32 kernels unless a bf16 autocast region is active (the reference's
    TRAIN.MIXED_PRECISION / torch.autocast contract) or the compute policy says bf16; bf16 tensors run the
    tensor-core kernels."""
    if x.dtype == torch.bfloat16 or x.dtype == torch.uint8:     # uint8 = raw frames, normalised on the device
        return torch.bfloat16
    if x.dtype == torch.float32:
        pol = compute_policy()
        if pol != "auto":
            return torch.bfloat16 if pol == "bf16" else torch.float32
        # any reduced-precision autocast region selects the tensor-core path; fp16 autocast (what the reference's
        # `torch.cuda.amp.autocast(enabled=TRAIN.MIXED_PRECISION)` requests, train_net.py:126) is served in bf16:
        # same 16-bit storage, wider exponent, so the GradScaler that accompanies it is harmless
        if torch.is_autocast_enabled("cuda") and torch.get_autocast_dtype("cuda") in (torch.bfloat16, torch.float16):
            return torch.bfloat16
        return torch.float32
    raise TypeError(f"unsupported activation dtype {x.dtype} (float32 or bfloat16)")


_SIDE = {}


def _side_streams(device):
    """Two auxiliary CUDA streams per device for the K / V pooling launches."""
    key = (device.type, device.index)
    if key not in _SIDE:
        _SIDE[key] = (torch.cuda.Stream(device=device), torch.cuda.Stream(device=device))
    return _SIDE[key]


def _pool_desc(pool):
    """(mode, kernel, stride, weight) of a reference-style pool module, or None."""
    if pool is None:
        return None
    t3 = lambda v: [int(v)] * 3 if isinstance(v, int) else [int(a) for a in v]
    if isinstance(pool, nn.Conv3d):
        k, s, p = t3(pool.kernel_size), t3(pool.stride), t3(pool.padding)
        if pool.groups != pool.in_channels or pool.in_channels != pool.out_channels or pool.bias is not None:
            raise NotImplementedError("attention_pool: only depthwise bias-free Conv3d pooling is supported")
        mode, w = "conv", pool.weight
    elif isinstance(pool, nn.MaxPool3d):
        k, s, p = t3(pool.kernel_size), t3(pool.stride), t3(pool.padding)
        if pool.ceil_mode:
            raise NotImplementedError("attention_pool: ceil_mode=True is not supported")
        mode, w = "max", None
    elif isinstance(pool, nn.AvgPool3d):
        k, s, p = t3(pool.kernel_size), t3(pool.stride), t3(pool.padding)
        if pool.ceil_mode or not pool.count_include_pad:
            raise NotImplementedError("attention_pool: AvgPool3d variant not supported")
        mode, w = "avg", None
    else:
        raise NotImplementedError(f"attention_pool: unsupported pool module {type(pool).__name__}")
    if p != [x // 2 for x in k]:
        raise NotImplementedError("attention_pool: padding must be kernel//2 (as built by the reference)")
    return mode, k, s, w


def attention_pool(tensor, pool, thw_shape, has_cls_embed=True, norm=None, pool2d=None):
    """attention.py:12-83.  tensor: [B, heads, L, d] (any strides with unit channel stride, e.g. a
    slice of the qkv output) or [B, L, C]; returns (pooled tensor, [T', H', W'])."""
    if pool is None:
        return tensor, thw_shape
    if pool2d is not None:
        raise NotImplementedError("pool2d (the reference's ONNX/TNN export aid) is not supported")
    desc = _pool_desc(pool)
    mode, k, s, w = desc
    if AG.recording(tensor, w, getattr(norm, "weight", None)):
        # differentiable form: the block's skip-path max pool; q/k/v pooling trains through MultiScaleAttention
        if tensor.ndim == 3 and mode == "max" and norm is None:
            return AG.maxpool_tokens(tensor, list(thw_shape), k, s, has_cls=has_cls_embed)
        if tensor.ndim == 4:
            ln = (norm.weight, norm.bias, norm.eps) if norm is not None else None
            return AG.pool_heads_generic(tensor, thw_shape, k, s, mode, w, ln, has_cls_embed)
        raise NotImplementedError("attention_pool: [B, L, C] input is differentiable for max pooling without norm only")
    ln = None
    if norm is not None:
        ln = (norm.weight, norm.bias, norm.eps)
    if tensor.ndim == 4:
        if tensor.stride(3) != 1:
            tensor = tensor.contiguous()
        out, thw = ops.attention_pool_heads(tensor, list(thw_shape), k, s, mode=mode, weight=w, ln=ln,
                                            has_cls=has_cls_embed)
        return out, thw
    if tensor.ndim == 3:
        if ln is None and mode != "conv":
            out, thw = ops.attention_pool_tokens(tensor, list(thw_shape), k, s, mode=mode, has_cls=has_cls_embed)
            return out, thw
        out, thw = ops.attention_pool_heads(tensor.unsqueeze(1), list(thw_shape), k, s, mode=mode, weight=w,
                                            ln=ln, has_cls=has_cls_embed)
        return out.squeeze(1), thw
    raise NotImplementedError(f"Unsupported input dimension {tensor.shape}")


class MultiScaleAttention(nn.Module):
    """attention.py:86-284."""

    def __init__(self, dim, num_heads=8, qkv_bias=False, drop_rate=0.0, kernel_q=(1, 1, 1),
                 kernel_kv=(1, 1, 1), stride_q=(1, 1, 1), stride_kv=(1, 1, 1), norm_layer=nn.LayerNorm,
                 has_cls_embed=True, mode="conv", pool_first=False, use_query_residual_pool=False,
                 expand_channel=False, expand_to_dim=None, rel_pos_spatial=False, rel_pos_temporal=False,
                 rel_pos_zero_init=False, input_size=None):
        super().__init__()
        self.drop_rate = drop_rate
        self.num_heads = num_heads
        dim_out = expand_to_dim if expand_channel else dim
        self.dim_out = dim_out
        head_dim = dim_out // num_heads
        self.scale = head_dim ** -0.5
        self.has_cls_embed = has_cls_embed
        pad_q = [int(q // 2) for q in kernel_q]
        pad_kv = [int(kv // 2) for kv in kernel_kv]

        self.qkv = nn.Linear(dim, dim_out * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim_out, dim_out)
        if drop_rate > 0.0:
            self.proj_drop = nn.Dropout(drop_rate)

        # kernel (1,1,1) with stride (1,1,1) means "no pooling" (attention.py:131-134)
        if numpy.prod(kernel_q) == 1 and numpy.prod(stride_q) == 1:
            kernel_q = ()
        if numpy.prod(kernel_kv) == 1 and numpy.prod(stride_kv) == 1:
            kernel_kv = ()

        def make(kind, kernel, stride, pad):
            if len(kernel) == 0:
                return None
            if kind == "conv":
                return nn.Conv3d(head_dim, head_dim, kernel, stride=stride, padding=pad, groups=head_dim,
                                 bias=False)
            cls = nn.AvgPool3d if kind == "avg" else nn.MaxPool3d
            return cls(kernel, stride, pad, ceil_mode=False)

        if mode in ("avg", "max"):
            self.pool_q = make(mode, kernel_q, stride_q, pad_q)
            self.pool_k = make(mode, kernel_kv, stride_kv, pad_kv)
            self.pool_v = make(mode, kernel_kv, stride_kv, pad_kv)
        elif mode == "conv":
            self.pool_q = make("conv", kernel_q, stride_q, pad_q)
            self.norm_q = norm_layer(head_dim) if len(kernel_q) > 0 else None
            self.pool_k = make("conv", kernel_kv, stride_kv, pad_kv)
            self.norm_k = norm_layer(head_dim) if len(kernel_kv) > 0 else None
            self.pool_v = make("conv", kernel_kv, stride_kv, pad_kv)
            self.norm_v = norm_layer(head_dim) if len(kernel_kv) > 0 else None
        else:
            raise NotImplementedError(f"Unsupported model {mode}")
        self.use_query_residual_pool = use_query_residual_pool

        # Decomposed relative-position bias: a DEFAULT-OFF extension that the reference does not have.
        # Parameter names / shapes / init follow upstream PySlowFast's MViTv2 (Appendix F): tables shared across heads,
        # [2*max(q_size, kv_size) - 1, head_dim].  Absent unless requested, so the default state_dict is the reference's.
        self.rel_pos_spatial, self.rel_pos_temporal = bool(rel_pos_spatial), bool(rel_pos_temporal)
        if self.rel_pos_spatial or self.rel_pos_temporal:
            if input_size is None or has_cls_embed or head_dim != 96:
                raise NotImplementedError("relative-position bias needs input_size=(T,H,W), no cls token and head_dim 96")
            t, hh, ww = input_size
            sq = stride_q if len(stride_q) > 0 else (1, 1, 1)
            skv = stride_kv if len(stride_kv) > 0 else (1, 1, 1)
            def table(n_q, n_kv):
                p = nn.Parameter(torch.zeros(2 * max(n_q, n_kv) - 1, head_dim))
                if not rel_pos_zero_init:
                    nn.init.trunc_normal_(p, std=0.02)
                return p
            self._rel_sizes = ((hh // sq[1], hh // skv[1]), (ww // sq[2], ww // skv[2]), (t // sq[0], t // skv[0]))
            if self.rel_pos_spatial:                      # upstream creates a table only for the term that is on
                self.rel_pos_h = table(*self._rel_sizes[0])
                self.rel_pos_w = table(*self._rel_sizes[1])
            if self.rel_pos_temporal:
                self.rel_pos_t = table(*self._rel_sizes[2])

    def _rel_operands(self, q, q_thw, k_thw):
        """(q_ext, k_ext) of ops.attention(rel=...) or None.  A table that is switched off contributes zeros."""
        if not (self.rel_pos_spatial or self.rel_pos_temporal):
            return None
        tabs = []
        for name, (nq, nk) in zip(("rel_pos_h", "rel_pos_w", "rel_pos_t"), self._rel_sizes):
            p = getattr(self, name, None)
            tabs.append(p if p is not None else torch.zeros(2 * max(nq, nk) - 1, 96, device=q.device))
        if AG.recording(q, *tabs):
            raise NotImplementedError("the relative-position bias is a forward-only (eval / no_grad) extension")
        return ops.relpos_operands(q, q_thw, k_thw, tabs[0], tabs[1], tabs[2], self.scale)

    # -- B200 forward ----------------------------------------------------------------------
    def _pooled(self, qkv5, which, pool, norm, thw_shape):
        """qkv5: [B, N, 3, h, d] GEMM output; returns contiguous [B, h, L', d] for q/k/v #which."""
        t = qkv5[:, :, which].permute(0, 2, 1, 3)          # [B, h, N, d] strided view, no copy
        if pool is None:
            # un-pooled operand: the head-major gather the attention kernel needs is a 1x1x1 "avg" pool
            out, _ = ops.attention_pool_heads(t, [1, 1, t.shape[2]], [1, 1, 1], [1, 1, 1], mode="avg")
            return out, list(thw_shape)
        return attention_pool(t, pool, thw_shape, has_cls_embed=self.has_cls_embed, norm=norm)

    def _train_descs(self):
        """Pooling description of q / k / v for the differentiable path (AG.pool_qkv)."""
        descs, params = [], []
        for pool, norm in ((self.pool_q, getattr(self, "norm_q", None)), (self.pool_k, getattr(self, "norm_k", None)),
                           (self.pool_v, getattr(self, "norm_v", None))):
            if pool is None:
                descs.append(None)
                params += [None, None, None]
                continue
            mode, k, s, w = _pool_desc(pool)
            if mode != "conv" or self.has_cls_embed:
                return None, None                      # secondary variants: generic differentiable path
            descs.append((tuple(k), tuple(s), norm.eps if norm is not None else None))
            params += [w, getattr(norm, "weight", None), getattr(norm, "bias", None)]
        return tuple(descs), params

    def _fused_pool_args(self, qkv):
        """(weights, lns, strides) for ops.attention_pool_qkv when all three pools are the shipped 3x3x3 depthwise convs with
        LayerNorm, no cls token, head_dim 96, bf16 — else None (generic per-tensor path)."""
        if self.has_cls_embed or self.dim_out // self.num_heads != 96:
            return None
        ws, lns, strides = [], [], []
        for pool, norm in ((self.pool_q, getattr(self, "norm_q", None)), (self.pool_k, getattr(self, "norm_k", None)),
                           (self.pool_v, getattr(self, "norm_v", None))):
            if pool is None or norm is None:
                return None
            mode, k, s, w = _pool_desc(pool)
            if mode != "conv" or list(k) != [3, 3, 3]:
                return None
            ws.append(w)
            lns.append((norm.weight, norm.bias, norm.eps))
            strides.append(tuple(s))
        if len({ln[2] for ln in lns}) != 1 or not ops.pool_qkv_supported(qkv, self.num_heads, strides):
            return None
        return ws, lns, strides

    def attend(self, x, thw_shape, ln_in=None):
        """Everything before `proj`: returns (y [B, Lq, C], out_thw) with y = softmax(qkᵀ·scale)v (+q).
        ln_in = (norm1, row statistics of x): `x` is the RAW block stream and norm1 is folded into the qkv GEMM."""
        B, N, _ = x.shape
        C, h = self.dim_out, self.num_heads
        if ln_in is not None:
            norm, stats = ln_in
            wf, bf, colsum = weights.folded_ln_linear(self.qkv.weight, self.qkv.bias, norm.weight, norm.bias)
            qkv = ops.linear_ln(x, stats, wf, bf, colsum, norm.eps)
        else:
            qkv = AG.linear(x, self.qkv.weight, self.qkv.bias)
        pool_params = [p for m in (self.pool_q, self.pool_k, self.pool_v, getattr(self, "norm_q", None),
                                   getattr(self, "norm_k", None), getattr(self, "norm_v", None)) if m is not None
                       for p in m.parameters()]
        if AG.recording(qkv, *pool_params):
            if self.rel_pos_spatial or self.rel_pos_temporal:
                raise NotImplementedError("the relative-position bias is a forward-only (eval / no_grad) extension")
            descs, params = self._train_descs()
            if descs is not None:
                (q, k, v), shapes = AG.pool_qkv(qkv, h, thw_shape, descs, params)
                return AG.attention(q, k, v, self.scale, self.use_query_residual_pool), shapes[0]
            # MVIT.MODE avg / max or a cls token: the same dataflow from generic differentiable pieces
            qkv5g = qkv.view(B, N, 3, h, C // h)
            parts, out_shape = [], list(thw_shape)
            for which, (pool, norm) in enumerate(((self.pool_q, getattr(self, "norm_q", None)),
                                                  (self.pool_k, getattr(self, "norm_k", None)),
                                                  (self.pool_v, getattr(self, "norm_v", None)))):
                t = qkv5g[:, :, which].permute(0, 2, 1, 3)
                if pool is None:
                    parts.append(t.contiguous())
                    continue
                t, shp = attention_pool(t, pool, thw_shape, has_cls_embed=self.has_cls_embed, norm=norm)
                if which == 0:
                    out_shape = shp
                parts.append(t.contiguous())
            return AG.attention(parts[0], parts[1], parts[2], self.scale, self.use_query_residual_pool), out_shape
        fused = self._fused_pool_args(qkv)
        if fused is not None:
            # q, k and v pooled (+ LayerNorm) by one call that reads the qkv GEMM output in place: persistent TMA-fed kernel,
            # tensors of equal stride share a launch
            ws, lns, strides = fused
            if strides[0] == strides[1]:
                (q, k, v), grids, _ = ops.attention_pool_qkv(qkv, h, list(thw_shape), ws, lns, strides)
            else:
                # two persistent launches (q; k + v): the second runs on a side stream so that its CTAs fill the SMs the first
                # one's last wave leaves idle (fork / join with events, graph-capturable)
                cur = torch.cuda.current_stream()
                side = _side_streams(x.device)[0]
                side.wait_stream(cur)
                with torch.cuda.stream(side):
                    (_, k, v), _, _ = ops.attention_pool_qkv(qkv, h, list(thw_shape), ws, lns, strides, only=(1, 2))
                    k.record_stream(cur)
                    v.record_stream(cur)
                (q, _, _), grids, _ = ops.attention_pool_qkv(qkv, h, list(thw_shape), ws, lns, strides, only=(0,))
                cur.wait_stream(side)
                qkv.record_stream(side)
            rel = self._rel_operands(q, grids[0], ops.pooled_thw(list(thw_shape), [3, 3, 3], list(strides[1])))
            return ops.attention(q, k, v, self.scale, self.use_query_residual_pool, rel=rel), grids[0]
        qkv5 = qkv.view(B, N, 3, h, C // h)
        # the three pooling launches are independent: K and V run on side streams next to Q so the small
        # deep-stage launches overlap instead of queueing (fork / join with events, graph-capturable)
        cur = torch.cuda.current_stream()
        side = _side_streams(x.device)
        fork = torch.cuda.Event()
        fork.record(cur)
        outs = [None, None]
        k_thw = list(thw_shape)
        for n, (which, pool, norm) in enumerate(((1, self.pool_k, getattr(self, "norm_k", None)),
                                                 (2, self.pool_v, getattr(self, "norm_v", None)))):
            with torch.cuda.stream(side[n]):
                side[n].wait_event(fork)
                outs[n], k_thw = self._pooled(qkv5, which, pool, norm, thw_shape)
                outs[n].record_stream(cur)
        q, out_shape = self._pooled(qkv5, 0, self.pool_q, getattr(self, "norm_q", None), thw_shape)
        for n in range(2):
            cur.wait_stream(side[n])
        qkv.record_stream(side[0])
        qkv.record_stream(side[1])
        k, v = outs
        y = ops.attention(q, k, v, self.scale, self.use_query_residual_pool, rel=self._rel_operands(q, out_shape, k_thw))
        return y, out_shape

    def forward(self, x, thw_shape, residual=None, row_scale=None, ln_in=None, want_stats=False):
        """Reference contract: (x [B,N,dim], thw) -> (proj(attn) [B,Lq,dim_out], thw').
        `residual` / `row_scale` (used by MultiScaleBlock) fold `x_res + drop_path(.)` into the proj GEMM.
        ln_in / want_stats (eval, bf16; MultiScaleBlock only): norm1 folded into the qkv GEMM, and the proj epilogue emits
        the row statistics norm2's folding needs."""
        dt = _compute_dtype(x)
        y, out_shape = self.attend(x.to(dt), thw_shape, ln_in)
        if want_stats:
            out = ops.linear_stats(y, weights.cached_weight(self.proj.weight, y.dtype), self.proj.bias, residual=residual,
                                   row_scale=row_scale)
            return out, out_shape
        if self.drop_rate > 0.0 and self.training:
            # MVIT.DROPOUT_RATE > 0 (0.0 in every shipped config): the dropout mask sits between the GEMM and the
            # residual add, so the epilogue fusion is split and the elementwise tail is left to PyTorch
            out = self.proj_drop(AG.linear(y, self.proj.weight, self.proj.bias))
            return AG.scale_add(out, residual, row_scale), out_shape
        out = AG.linear(y, self.proj.weight, self.proj.bias, residual=residual, row_scale=row_scale)
        return out, out_shape


class MultiScaleBlock(nn.Module):
    """attention.py:287-446."""

    def __init__(self, dim, dim_out, num_heads, mlp_ratio=4.0, qkv_bias=False, qk_scale=None, drop_rate=0.0,
                 drop_path=0.0, act_layer=nn.GELU, norm_layer=nn.LayerNorm, up_rate=None,
                 kernel_q=(1, 1, 1), kernel_kv=(1, 1, 1), stride_q=(1, 1, 1), stride_kv=(1, 1, 1),
                 mode="conv", has_cls_embed=True, pool_first=False, use_query_residual_pool=False,
                 channel_expand_front=False, pool_skip_use_conv=False, rel_pos_spatial=False, rel_pos_temporal=False,
                 rel_pos_zero_init=False, input_size=None):
        super().__init__()
        self.dim = dim
        self.dim_out = dim_out
        self.norm1 = norm_layer(dim)
        kernel_skip = [s + 1 if s > 1 else s for s in stride_q]
        stride_skip = stride_q
        padding_skip = [int(skip // 2) for skip in kernel_skip]

        dim_in = dim
        self.expand_channel = bool(channel_expand_front and dim != dim_out)
        self.pool_skip_use_conv = pool_skip_use_conv

        # the reference hands the bare nn.LayerNorm (eps 1e-5) to the attention (attention.py:338, D5)
        self.attn = MultiScaleAttention(
            dim, num_heads=num_heads, qkv_bias=qkv_bias, drop_rate=drop_rate, kernel_q=kernel_q,
            kernel_kv=kernel_kv, stride_q=stride_q, stride_kv=stride_kv, norm_layer=nn.LayerNorm,
            has_cls_embed=has_cls_embed, mode=mode, use_query_residual_pool=use_query_residual_pool,
            expand_channel=self.expand_channel, expand_to_dim=dim_out, rel_pos_spatial=rel_pos_spatial,
            rel_pos_temporal=rel_pos_temporal, rel_pos_zero_init=rel_pos_zero_init, input_size=input_size)
        if self.expand_channel:
            dim = dim_out
            self.dim = dim_out

        self.drop_path = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()
        self.norm2 = norm_layer(dim)
        mlp_hidden_dim = int(dim * mlp_ratio)
        self.has_cls_embed = has_cls_embed
        mlp_dim_out = dim * up_rate if up_rate is not None and up_rate > 1 else dim_out
        self.mlp = Mlp(in_features=dim, hidden_features=mlp_hidden_dim, out_features=mlp_dim_out,
                       act_layer=act_layer, drop_rate=drop_rate)
        if dim != dim_out:
            self.proj = nn.Linear(dim, dim_out)
        if self.expand_channel:
            self.proj_max_pool = nn.Linear(dim_in, dim_out)
        self.pool_skip = (nn.MaxPool3d(kernel_skip, stride_skip, padding_skip, ceil_mode=False)
                          if len(kernel_skip) > 0 else None)
        self.pool_skip_norm = None

    def _skip_is_identity(self):
        p = self.pool_skip
        if p is None:
            return True
        k = [p.kernel_size] * 3 if isinstance(p.kernel_size, int) else list(p.kernel_size)
        s = [p.stride] * 3 if isinstance(p.stride, int) else list(p.stride)
        return all(v == 1 for v in k) and all(v == 1 for v in s)   # MaxPool3d([1,1,1]) == identity (D6)

    def out_thw(self, thw_shape):
        """Token grid after this block (the q pooling decides it), without running it."""
        desc = _pool_desc(self.attn.pool_q)
        return list(thw_shape) if desc is None else ops.pooled_thw(list(thw_shape), desc[1], desc[2])

    def forward(self, x, thw_shape):
        dt = _compute_dtype(x)
        x = x.to(dt).contiguous()
        B = x.shape[0]
        p_drop = getattr(self.drop_path, "drop_prob", 0.0) or 0.0
        # one draw per drop_path call, in the reference's order (attention branch, then MLP branch)
        s_attn = drop_path_scale(B, p_drop, self.training, x.device)
        s_mlp = drop_path_scale(B, p_drop, self.training, x.device)

        # Eval / bf16: neither LayerNorm runs as a kernel.  The GEMM that produced the block stream (patch embed, the previous
        # block's fc2, this block's proj) left per-row (sum, sum of squares) next to it; qkv and fc1 read the raw stream with
        # gamma folded into their weights and finish the normalisation in their epilogues (gemm_tc.cu, kLnIn / kLnStats).
        fold = self._ln_fold_ok(x)
        stats1 = ops.row_stats_of(x) if fold else None
        xn = x if stats1 is not None else AG.layernorm(x, self.norm1.weight, self.norm1.bias, self.norm1.eps)
        if self.expand_channel and not self.pool_skip_use_conv:
            x = AG.linear(x, self.proj_max_pool.weight, self.proj_max_pool.bias)
        if self._skip_is_identity():
            x_res = x
        else:
            x_res, _ = attention_pool(x, self.pool_skip, thw_shape, has_cls_embed=self.has_cls_embed)
        # x = x_res + drop_path(proj(attn))  — residual and DropPath scale live in the proj GEMM epilogue
        x, thw_new = self.attn(xn, thw_shape, residual=x_res, row_scale=s_attn,
                               ln_in=(self.norm1, stats1) if stats1 is not None else None, want_stats=fold)
        stats2 = ops.row_stats_of(x) if fold and self.dim == self.dim_out else None
        if stats2 is not None:
            return self.mlp(x, residual=x, row_scale=s_mlp, ln_in=(self.norm2, stats2), want_stats=True), thw_new
        x_norm = AG.layernorm(x, self.norm2.weight, self.norm2.bias, self.norm2.eps)
        if self.dim != self.dim_out:
            x = AG.linear(x_norm, self.proj.weight, self.proj.bias)
        # out = x + drop_path(mlp(x_norm)) — residual and scale in the fc2 GEMM epilogue
        out = self.mlp(x_norm, residual=x, row_scale=s_mlp, want_stats=fold)
        return out, thw_new

    def _ln_fold_ok(self, x):
        if not (ops.ln_fold_enabled() and x.is_cuda and x.dtype == torch.bfloat16 and not self.has_cls_embed):
            return False
        if self.attn.drop_rate > 0.0 and self.training:
            return False
        params = [self.norm1.weight, self.norm1.bias, self.norm2.weight, self.norm2.bias, self.attn.qkv.weight,
                  self.attn.proj.weight, self.mlp.fc1.weight, self.mlp.fc2.weight]
        if AG.recording(x, *params):
            return False
        # 16-byte TMA rows for every GEMM on the path
        return all(d % 8 == 0 for d in (self.attn.qkv.in_features, self.attn.proj.out_features, self.mlp.fc1.out_features,
                                        self.mlp.fc2.out_features))
