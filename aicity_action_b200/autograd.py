"""Differentiable forms of the fused ops: every forward launch of `ops` paired with its backward launches.

The reference has no backward code of its own — `tools/train_net.py:229-246` calls `loss.backward()` and autograd
differentiates attention.py / common.py op by op.  Here each *fused* forward op is one `torch.autograd.Function`
whose backward issues the derivative kernels of libmvit_b200.so (csrc/backward.cu; input-gradient GEMMs reuse the
forward GEMM with the transposed weight).  The functional wrappers below take the no-graph fast path when nothing
requires grad (inference under `torch.no_grad()` is unchanged).  Recompute-instead-of-store is used where the forward
fuses an activation (GELU) or a normalisation (the pooling LayerNorm): the pre-activation is rebuilt in backward by
re-running the cheap producer rather than written to HBM in forward.

Parameter gradients are fp32 (the master dtype), activations / activation gradients are the compute dtype.
There is no CPU implementation.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch
from torch.autograd import Function

from . import ops
from .weights import cached_weight, cached_weight_t


def recording(*tensors) -> bool:
    """True when autograd must see this op (grad mode on and some operand requires grad)."""
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)


def _like_param(g: Optional[torch.Tensor], p: Optional[torch.Tensor]):
    if g is None or p is None:
        return None
    return g.reshape(p.shape).to(p.dtype)


# ------------------------------------------------------------------------------------------------ LayerNorm
class _LayerNorm(Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, eps):
        x = x.contiguous()
        ctx.save_for_backward(x, gamma)
        ctx.eps = eps
        ctx.beta = beta
        return ops.layernorm(x, gamma, beta, eps)

    @staticmethod
    def backward(ctx, dy):
        x, gamma = ctx.saved_tensors
        dx, dg, db = ops.layernorm_bwd(x, gamma, dy.to(x.dtype), ctx.eps)
        return dx, _like_param(dg, gamma), _like_param(db, ctx.beta), None


def layernorm(x, gamma, beta, eps):
    if recording(x, gamma, beta):
        return _LayerNorm.apply(x, gamma, beta, eps)
    return ops.layernorm(x, gamma, beta, eps)


# ------------------------------------------------------------------------------------------------ Linear (+ epilogue)
class _Linear(Function):
    """y = act(x·Wᵀ + b) · row_scale + residual   (nn.Linear + GELU + DropPath + residual add, one GEMM)."""

    @staticmethod
    def forward(ctx, x, weight, bias, residual, row_scale, gelu):
        x = x.contiguous()
        ctx.save_for_backward(x, weight, bias, row_scale)
        ctx.gelu = gelu
        ctx.has_residual = residual is not None
        return ops.linear(x, cached_weight(weight, x.dtype), bias, residual=residual, row_scale=row_scale, gelu=gelu)

    @staticmethod
    def backward(ctx, dy):
        x, weight, bias, row_scale = ctx.saved_tensors
        dy = dy.to(x.dtype).contiguous()
        d_res = dy if ctx.has_residual else None
        dz = dy
        if row_scale is not None:                       # DropPath: a per-sample constant multiplier
            nb = row_scale.numel()
            dz = (dy.view(nb, -1) * row_scale.to(dy.dtype).view(nb, 1)).view(dy.shape)
        if ctx.gelu:
            # the pre-activation is rebuilt (not stored in forward) and differentiated in the same GEMM's epilogue:
            # dz <- dz * gelu'(x.W^T + b), with dz riding in as the epilogue's "residual" tile
            dz = ops.linear(x, cached_weight(weight, x.dtype), bias, residual=dz, gelu_grad=True)
        dx = ops.linear(dz, cached_weight_t(weight, x.dtype)) if ctx.needs_input_grad[0] else None
        dw = db = None
        if ctx.needs_input_grad[1] or (bias is not None and ctx.needs_input_grad[2]):
            dw, db = ops.linear_wgrad(dz, x, bias is not None)
        return dx, _like_param(dw, weight), _like_param(db, bias), d_res, None, None


def linear(x, weight, bias=None, *, residual=None, row_scale=None, gelu=False):
    """`weight` / `bias` are the fp32 master parameters (the compute-dtype copy is cached)."""
    if recording(x, weight, bias, residual):
        return _Linear.apply(x, weight, bias, residual, row_scale, gelu)
    return ops.linear(x, cached_weight(weight, x.dtype), bias, residual=residual, row_scale=row_scale, gelu=gelu)


def scale_add(y, residual=None, row_scale=None):
    """residual + y * row_scale with plain tensor ops (only used when a dropout mask forbids the epilogue fusion)."""
    if row_scale is not None:
        y = y * row_scale.to(y.dtype).view(-1, *([1] * (y.ndim - 1)))
    return y if residual is None else residual + y


# ------------------------------------------------------------------------------------------------ attention
class _Attention(Function):
    @staticmethod
    def forward(ctx, q, k, v, scale, add_q):
        out, lse = ops.attention(q, k, v, scale, add_q, want_lse=True)
        ctx.save_for_backward(q, k, v, out, lse)
        ctx.scale, ctx.add_q = scale, add_q
        return out

    @staticmethod
    def backward(ctx, dout):
        q, k, v, out, lse = ctx.saved_tensors
        dq, dk, dv = ops.attention_bwd(q, k, v, out, dout.to(q.dtype), lse, ctx.scale, ctx.add_q)
        return dq, dk, dv, None, None


def attention(q, k, v, scale, add_q):
    if recording(q, k, v):
        return _Attention.apply(q, k, v, scale, add_q)
    return ops.attention(q, k, v, scale, add_q)


# ------------------------------------------------------------------------------------------------ q/k/v pooling
class _PoolQKV(Function):
    """qkv GEMM output [B, N, 3·h·d] -> pooled (q, k, v) [B, h, L', d]: depthwise Conv3d ✚ LayerNorm read in place from
    the GEMM output (attention.py:172-212); an operand whose pool is None is only re-laid out head-major.
    Backward writes the three input gradients straight into one [B, N, 3, h, d] buffer."""

    @staticmethod
    def forward(ctx, qkv, heads, thw, descs, *params):
        # descs[i] = None | (kernel, stride, eps | None);  params = (w, gamma, beta) x 3 (None where absent)
        B, N, C3 = qkv.shape
        d = C3 // (3 * heads)
        qkv5 = qkv.view(B, N, 3, heads, d)
        outs, shapes, pres = [], [], [None, None, None]
        strides3 = [tuple(dsc[1]) if dsc is not None else None for dsc in descs]
        if (d == 96 and all(dsc is not None and tuple(dsc[0]) == (3, 3, 3) and dsc[2] is not None for dsc in descs)
                and len({dsc[2] for dsc in descs}) == 1 and all(params[3 * i + 1] is not None for i in range(3))
                and ops.pool_qkv_supported(qkv, heads, strides3)):
            # the shipped configuration: one call pools q, k and v (persistent TMA kernel) and keeps their pre-LayerNorm values
            outs, shapes, pres = ops.attention_pool_qkv(qkv, heads, list(thw), [params[0], params[3], params[6]],
                                                        [(params[3 * i + 1], params[3 * i + 2], descs[i][2]) for i in range(3)],
                                                        strides3, save_pre=True)
            ctx.save_for_backward(qkv, *params, *pres)
            ctx.meta = (heads, list(thw), descs)
            ctx.out_tokens = [sh[0] * sh[1] * sh[2] for sh in shapes]
            return tuple(outs)
        for i in range(3):
            t = qkv5[:, :, i].permute(0, 2, 1, 3)
            w, g, b = params[3 * i:3 * i + 3]
            if descs[i] is None:
                o, _ = ops.attention_pool_heads(t, [1, 1, N], [1, 1, 1], [1, 1, 1], mode="avg")
                shapes.append(list(thw))
            else:
                kernel, stride, eps = descs[i]
                if g is not None and ops.pool_save_supported(t, kernel, stride):
                    # keep the conv output for the LayerNorm backward (one extra store) instead of recomputing it
                    o, pres[i], s = ops.attention_pool_heads_save(t, list(thw), kernel, stride, w, (g, b, eps))
                else:
                    ln = (g, b, eps) if g is not None else None
                    o, s = ops.attention_pool_heads(t, list(thw), kernel, stride, mode="conv", weight=w, ln=ln)
                shapes.append(s)
            outs.append(o)
        ctx.save_for_backward(qkv, *params, *pres)
        ctx.meta = (heads, list(thw), descs)
        ctx.out_tokens = [s[0] * s[1] * s[2] for s in shapes]
        return tuple(outs)

    @staticmethod
    def backward(ctx, *douts):
        qkv, *rest = ctx.saved_tensors
        params, pres = rest[:9], rest[9:]
        heads, thw, descs = ctx.meta
        B, N, C3 = qkv.shape
        d = C3 // (3 * heads)
        qkv5 = qkv.view(B, N, 3, heads, d)
        dqkv = torch.empty_like(qkv)
        dqkv5 = dqkv.view(B, N, 3, heads, d)
        strides = (N * C3, C3, d)                       # (batch, token, head) element strides of a q/k/v slice
        grads = []
        for i in range(3):
            w, g, b = params[3 * i:3 * i + 3]
            dy = douts[i]
            if dy is None:
                dy = torch.zeros((B, heads, ctx.out_tokens[i], d), dtype=qkv.dtype, device=qkv.device)
            dy = dy.to(qkv.dtype).contiguous()
            if descs[i] is None:
                dqkv5[:, :, i].copy_(dy.permute(0, 2, 1, 3))
                grads += [None, None, None]
                continue
            kernel, stride, eps = descs[i]
            x_view = qkv5[:, :, i]
            dg = db = None
            if g is not None:                           # differentiate the LayerNorm on the saved (or rebuilt) conv output
                conv = pres[i]
                if conv is None:
                    conv, _ = ops.attention_pool_heads(x_view.permute(0, 2, 1, 3), thw, kernel, stride, mode="conv",
                                                       weight=w)
                dy, dg, db = ops.layernorm_bwd(conv, g, dy, eps)
            dw = torch.zeros((d, kernel[0] * kernel[1] * kernel[2]), dtype=torch.float32, device=qkv.device)
            ops.attention_pool_bwd(1, x_view, strides, dy, None, None, dw, B, heads, d, thw, kernel, stride)
            if tuple(stride) == (1, 1, 1) and all(k % 2 == 1 for k in kernel):
                # unit stride: the input gradient is the same depthwise convolution with the taps reversed -> the tuned
                # forward kernel, reading dy and writing straight into the q/k/v slice of dqkv
                Lo = dy.shape[2]
                ops.attention_pool_strided(dy, 0, (heads * Lo * d, d, Lo * d), B, heads, d, thw, kernel, stride, "conv",
                                           w.detach().reshape(d, -1).flip(1).contiguous(), None, None, 0.0, False,
                                           dqkv5[:, :, i], strides)
            else:
                ops.attention_pool_bwd(0, None, strides, dy, w.reshape(d, -1), dqkv5[:, :, i], None, B, heads, d, thw,
                                       kernel, stride)
            grads += [_like_param(dw, w), _like_param(dg, g), _like_param(db, b)]
        return (dqkv, None, None, None, *grads)


def pool_qkv(qkv, heads, thw, descs, params):
    """-> ((q, k, v), (thw_q, thw_k, thw_v)).  `descs`/`params` as in _PoolQKV.forward."""
    if recording(qkv, *params):
        outs = _PoolQKV.apply(qkv, heads, list(thw), descs, *params)
        shapes = []
        for i in range(3):
            if descs[i] is None:
                shapes.append(list(thw))
            else:
                shapes.append(ops.pooled_thw(list(thw), descs[i][0], descs[i][1]))
        return outs, shapes
    raise RuntimeError("pool_qkv is the training-path entry; inference uses attention.attention_pool directly")


# ------------------------------------------------------------------------------------------------ generic pooling
class _PoolHeads(Function):
    """Pool3d of one [B, heads, L, d] operand (any strided view with unit channel stride) without cls token or LayerNorm:
    mode "conv" (depthwise, weight [d,1,kt,kh,kw]), "avg" or "max".  The secondary variants of the reference
    (MVIT.MODE avg / max, a cls token) are assembled from this, `layernorm` and plain tensor slicing / concatenation —
    correct, on the generic kernels, not tuned (the shipped configs take the fused `_PoolQKV` path)."""

    @staticmethod
    def forward(ctx, x, weight, mode, thw, kernel, stride):
        out, _ = ops.attention_pool_heads(x, list(thw), list(kernel), list(stride), mode=mode, weight=weight)
        ctx.save_for_backward(x, weight)
        ctx.meta = (mode, list(thw), list(kernel), list(stride))
        return out

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        mode, thw, kernel, stride = ctx.meta
        B, heads, L, d = x.shape
        dy = dy.to(x.dtype).contiguous()
        xs = (x.stride(0), x.stride(2), x.stride(1))                   # (batch, token, head)
        dxb = torch.empty((B, L, heads, d), dtype=x.dtype, device=x.device)
        ds = (L * heads * d, heads * d, d)
        dw = None
        if mode == "max":
            dx32 = torch.zeros((B, L, heads, d), dtype=torch.float32, device=x.device)
            ops.attention_pool_bwd(3, x, xs, dy, None, dx32, None, B, heads, d, thw, kernel, stride)
            dxb.copy_(dx32)
        else:
            taps = kernel[0] * kernel[1] * kernel[2]
            if mode == "conv":
                dw = torch.zeros((d, taps), dtype=torch.float32, device=x.device)
                ops.attention_pool_bwd(1, x, xs, dy, None, None, dw, B, heads, d, thw, kernel, stride)
                wmat = weight.reshape(d, -1)
            else:                                                      # avg: count_include_pad -> every tap weighs 1/taps
                wmat = torch.full((d, taps), 1.0 / taps, dtype=torch.float32, device=x.device)
            ops.attention_pool_bwd(0, None, ds, dy, wmat, dxb, None, B, heads, d, thw, kernel, stride)
        return dxb.permute(0, 2, 1, 3), _like_param(dw, weight), None, None, None, None


def pool_heads_generic(t, thw, kernel, stride, mode, weight, ln, has_cls):
    """Differentiable attention_pool of one [B, heads, L(+1), d] operand for the secondary variants."""
    cls = None
    if has_cls:
        cls, t = t[:, :, :1], t[:, :, 1:]
    out = _PoolHeads.apply(t, weight, mode, list(thw), list(kernel), list(stride))
    if cls is not None:
        out = torch.cat((cls, out), dim=2)
    if ln is not None:
        out = layernorm(out, ln[0], ln[1], ln[2])
    return out, ops.pooled_thw(list(thw), kernel, stride)


# ------------------------------------------------------------------------------------------------ skip-path max pool
class _MaxPoolTokens(Function):
    @staticmethod
    def forward(ctx, x, thw, kernel, stride):
        x = x.contiguous()
        ctx.save_for_backward(x)
        ctx.meta = (list(thw), list(kernel), list(stride))
        out, _ = ops.attention_pool_tokens(x, list(thw), kernel, stride, mode="max")
        return out

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        thw, kernel, stride = ctx.meta
        B, L, C = x.shape
        d = 96 if C % 96 == 0 else 32
        dx = torch.zeros((B, L, C), dtype=torch.float32, device=x.device)
        ops.attention_pool_bwd(2, x, (L * C, C, d), dy.to(x.dtype).contiguous(), None, dx, None, B, C // d, d, thw, kernel,
                               stride)
        return dx.to(x.dtype), None, None, None


def maxpool_tokens(x, thw, kernel, stride, has_cls=False):
    if recording(x):
        if has_cls:                                   # the cls row bypasses pooling (attention.py:28-29, 62-64)
            pooled = _MaxPoolTokens.apply(x[:, 1:], thw, kernel, stride)
            return torch.cat((x[:, :1], pooled), dim=1), ops.pooled_thw(list(thw), kernel, stride)
        return _MaxPoolTokens.apply(x, thw, kernel, stride), ops.pooled_thw(list(thw), kernel, stride)
    return ops.attention_pool_tokens(x, list(thw), kernel, stride, mode="max", has_cls=has_cls)


# ------------------------------------------------------------------------------------------------ patch embedding
class _PatchEmbed(Function):
    """Conv3d patch embedding + separable positional embedding (stem_helper.py:336, video_model_builder.py:1206-1223).
    Forward = the module's fused inference path; backward = im2col (recomputed) + the Linear weight-gradient kernel, and
    two reductions for the positional tables.  The clip itself never needs a gradient."""

    @staticmethod
    def forward(ctx, clip, weight, bias, pos_spatial, pos_temporal, module, dtype, pos_table):
        ctx.save_for_backward(clip, weight, bias, pos_spatial, pos_temporal)
        ctx.module, ctx.dtype = module, dtype
        return module(clip, dtype, pos=pos_table, pos_period=0 if pos_table is None else pos_table.shape[0])

    @staticmethod
    def backward(ctx, dy):
        clip, weight, bias, pos_spatial, pos_temporal = ctx.saved_tensors
        m, dtype = ctx.module, ctx.dtype
        dy = dy.to(dtype).contiguous()
        B, L, N = dy.shape
        x = ops.preprocess_u8(clip, dtype) if clip.dtype == torch.uint8 else clip.to(dtype)
        kernel, stride, padding = list(m.proj.kernel_size), list(m.proj.stride), list(m.proj.padding)
        k = weight[0].numel()
        patches, _ = ops.im2col3d(x, kernel, stride, padding, (k + 63) // 64 * 64)
        dw, db = ops.linear_wgrad(dy.view(B * L, N), patches, bias is not None)
        dps = dpt = None
        if pos_spatial is not None:
            T, HW = pos_temporal.shape[1], pos_spatial.shape[1]
            g = dy.view(B, T, HW, N).float()
            dps, dpt = g.sum((0, 1)), g.sum((0, 2))
        return (None, _like_param(dw[:, :k].contiguous(), weight), _like_param(db, bias), _like_param(dps, pos_spatial),
                _like_param(dpt, pos_temporal), None, None, None)


def patch_embed(module, clip, dtype, pos_spatial=None, pos_temporal=None, pos_table=None):
    if module.conv_2d:
        raise NotImplementedError("training with MVIT.PATCH_2D is not supported by the B200 path")
    return _PatchEmbed.apply(clip, module.proj.weight, module.proj.bias, pos_spatial, pos_temporal, module, dtype,
                             pos_table)
