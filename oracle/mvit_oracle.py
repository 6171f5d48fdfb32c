"""CPU oracle for the MViTv2 multiscale-attention path.   *** TEST INFRASTRUCTURE ***

A plain-PyTorch (CPU, fp32 or fp64) restatement of the algorithm the reference
runs for `attention_pool`, `MultiScaleAttention`, `MultiScaleBlock` and `MViT`.
It is NOT the product: only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it.  The
product package (`aicity_action_b200`) never does.

Parity pin: `oracle/make_golden.py` executes the UNMODIFIED reference modules
(`/root/reference`, imported through `oracle/ref_shims.py`) on seeded inputs,
asserts this restatement agrees with them, and writes the fixtures under
`tests/golden/`.  `tests/test_oracle_golden.py` re-checks the restatement against
those fixtures on every CPU run (the reference tree does not travel to the GPU
box).  The reference ships no tests of its own, so the fixtures
generated from its code are the only pin there is.

One function has NO pin: `rel_pos_bias` (the default-off relative-position term of north_star item 2) does not exist
in the reference and restates upstream PySlowFast from SURVEY.md Appendix F — parity unpinned.

Everything is written functionally over a flat `state_dict` (names/shapes of
SURVEY.md Appendix B) so the same weights drive the reference, this oracle and
the CUDA path.  Reference citations are `file:line` under /root/reference.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# ----------------------------------------------------------------------------
# host-side integer logic
# ----------------------------------------------------------------------------
def round_width(width, multiplier, min_width=1, divisor=1):
    """slowfast/models/utils.py:8-22 — channel / head rounding."""
    if not multiplier:
        return width
    width *= multiplier
    min_width = min_width or divisor
    out = max(min_width, int(width + divisor / 2) // divisor * divisor)
    if out < 0.9 * width:
        out += divisor
    return int(out)


@dataclass
class BlockSpec:
    dim: int            # block input channels
    dim_out: int        # block output channels
    heads: int
    kernel_q: List[int]
    kernel_kv: List[int]
    stride_q: List[int]
    stride_kv: List[int]
    drop_path: float
    expand: bool        # attention produces dim_out channels (CHANNEL_EXPAND_FRONT)
    attn_dim: int = 0   # channels inside attention (= dim_out if expand else dim)

    def __post_init__(self):
        self.attn_dim = self.dim_out if self.expand else self.dim


@dataclass
class MViTSpec:
    blocks: List[BlockSpec]
    patch_dims: List[int]
    embed_dim: int
    patch_kernel: List[int]
    patch_stride: List[int]
    patch_padding: List[int]
    cls_embed_on: bool
    sep_pos_embed: bool
    mode: str
    q_pool_residual: bool
    num_classes: int
    final_norm: bool
    head_act: str = "softmax"
    mlp_ratio: float = 4.0


def derive_spec(cfg) -> MViTSpec:
    """Per-block (dim, dim_out, heads, pool kernels/strides) from the MVIT.* keys.

    video_model_builder.py:846-1038.  `cfg` is anything with attribute/key access
    to DATA / MVIT / MODEL (the reference CfgNode, or the product's own)."""
    g = lambda node, key: node[key] if isinstance(node, dict) else getattr(node, key)
    M, D, MO = g(cfg, "MVIT"), g(cfg, "DATA"), g(cfg, "MODEL")
    depth = g(M, "DEPTH")
    size = g(D, "TRAIN_CROP_SIZE")
    assert size == g(D, "TEST_CROP_SIZE")                      # :805
    patch_stride = list(g(M, "PATCH_STRIDE"))
    if g(M, "PATCH_2D"):
        patch_stride = [1] + patch_stride                      # :829-830
    in_dims = [g(D, "NUM_FRAMES"), size, size]
    patch_dims = [in_dims[i] // patch_stride[i] for i in range(3)]   # :867-870
    dpr = [x.item() for x in torch.linspace(0, g(M, "DROPPATH_RATE"), depth)]  # :880-882

    dim_mul = [1.0] * (depth + 1)
    head_mul = [1.0] * (depth + 1)
    for i, m in g(M, "DIM_MUL"):
        dim_mul[i] = m                                         # :921-924
    for i, m in g(M, "HEAD_MUL"):
        head_mul[i] = m

    kq = [[] for _ in range(depth)]
    kkv = [[] for _ in range(depth)]
    sq = [[] for _ in range(depth)]
    skv = [[] for _ in range(depth)]
    kvq_kernel = g(M, "POOL_KVQ_KERNEL")
    for ent in g(M, "POOL_Q_STRIDE"):                          # :932-948
        i, s = ent[0], list(ent[1:])
        sq[i] = s
        kq[i] = list(kvq_kernel) if kvq_kernel is not None else [x + 1 if x > 1 else x for x in s]
    if g(M, "Q_POOL_ALL"):                                     # :951-955
        for i in range(depth):
            if not kq[i]:
                kq[i] = list(kvq_kernel)
                sq[i] = [1, 1, 1]
    kv_stride = g(M, "POOL_KV_STRIDE")
    adaptive = g(M, "POOL_KV_STRIDE_ADAPTIVE")
    if adaptive is not None:                                   # :958-967
        cur = list(adaptive)
        kv_stride = []
        for i in range(depth):
            if len(sq[i]) > 0:
                cur = [max(cur[d] // sq[i][d], 1) for d in range(len(cur))]
            kv_stride.append([i] + cur)
    for ent in kv_stride:                                      # :969-980
        i, s = ent[0], list(ent[1:])
        skv[i] = s
        kkv[i] = list(kvq_kernel) if kvq_kernel is not None else [x + 1 if x > 1 else x for x in s]

    expand_front = g(M, "CHANNEL_EXPAND_FRONT")
    heads = g(M, "NUM_HEADS")
    embed = g(M, "EMBED_DIM")
    dim_out = embed
    blocks = []
    for i in range(depth):                                     # :997-1038
        heads = round_width(heads, head_mul[i])
        if expand_front:
            mul = 1.0 if i == 0 else dim_mul[i - 1]
            embed = round_width(embed, mul, divisor=heads)
            dim_out = round_width(dim_out, dim_mul[i], divisor=heads)
        else:
            embed = round_width(embed, dim_mul[i], divisor=heads)
            dim_out = round_width(embed, dim_mul[i + 1], divisor=round_width(heads, head_mul[i + 1]))
        blocks.append(BlockSpec(embed, dim_out, heads, kq[i], kkv[i], sq[i], skv[i], dpr[i],
                                expand=bool(expand_front and embed != dim_out)))
    return MViTSpec(
        blocks=blocks, patch_dims=patch_dims, embed_dim=g(M, "EMBED_DIM"),
        patch_kernel=list(g(M, "PATCH_KERNEL")), patch_stride=list(g(M, "PATCH_STRIDE")),
        patch_padding=list(g(M, "PATCH_PADDING")), cls_embed_on=bool(g(M, "CLS_EMBED_ON")),
        sep_pos_embed=bool(g(M, "SEP_POS_EMBED")), mode=g(M, "MODE"),
        q_pool_residual=bool(g(M, "Q_POOL_RESIDUAL")), num_classes=g(MO, "NUM_CLASSES"),
        final_norm=not g(M, "NO_NORM_BEFORE_AVG"), head_act=g(MO, "HEAD_ACT"),
        mlp_ratio=g(M, "MLP_RATIO"))


def pooled_thw(thw: Sequence[int], kernel: Sequence[int], stride: Sequence[int]) -> List[int]:
    """Conv3d / MaxPool3d output size with pad = k//2, ceil_mode=False (attention.py:58, 121-122)."""
    return [(n + 2 * (k // 2) - k) // s + 1 for n, k, s in zip(thw, kernel, stride)]


# ----------------------------------------------------------------------------
# attention_pool                                       attention.py:12-83
# ----------------------------------------------------------------------------
def attention_pool(x: Tensor, thw: Sequence[int], *, mode: str, kernel: Sequence[int],
                   stride: Sequence[int], weight: Optional[Tensor] = None,
                   has_cls: bool = False, ln: Optional[Tuple[Tensor, Tensor, float]] = None):
    """x: [B, h, L, d] (or [B, L, C]); returns (pooled, thw').

    mode 'conv' = depthwise Conv3d (weight [d,1,kt,kh,kw], zero pad k//2, no bias);
    'max' / 'avg' = MaxPool3d / AvgPool3d with pad k//2.  `ln=(gamma,beta,eps)` is the
    LayerNorm over d applied after pooling (attention.py:66-67).  A cls token (first
    row) bypasses pooling and joins the LayerNorm (attention.py:28-29, 62-64)."""
    squeeze = x.ndim == 3
    if squeeze:
        x = x.unsqueeze(1)                                     # :21-23
    cls = None
    if has_cls:
        cls, x = x[:, :, :1], x[:, :, 1:]
    B, h, L, d = x.shape
    T, H, W = thw
    assert L == T * H * W
    vol = x.reshape(B * h, T, H, W, d).permute(0, 4, 1, 2, 3)  # :34-36 channels-first
    pad = [k // 2 for k in kernel]
    if mode == "conv":
        vol = F.conv3d(vol, weight, None, stride=tuple(stride), padding=tuple(pad), groups=d)
    elif mode == "max":
        vol = F.max_pool3d(vol, tuple(kernel), tuple(stride), tuple(pad))
    elif mode == "avg":
        vol = F.avg_pool3d(vol, tuple(kernel), tuple(stride), tuple(pad))
    else:
        raise NotImplementedError(mode)
    thw2 = list(vol.shape[2:])
    out = vol.reshape(B, h, d, -1).transpose(2, 3)             # :58-60
    if cls is not None:
        out = torch.cat((cls, out), dim=2)
    if ln is not None:
        out = F.layer_norm(out, (d,), ln[0], ln[1], ln[2])
    if squeeze:
        out = out.reshape(B, out.shape[2], d)
    return out, thw2


# ----------------------------------------------------------------------------
# MultiScaleAttention.forward                          attention.py:222-284
# ----------------------------------------------------------------------------
POOL_LN_EPS = 1e-5   # attention.py:338 passes the bare nn.LayerNorm
BLOCK_LN_EPS = 1e-6  # video_model_builder.py:848-850


def rel_pos_bias(q: Tensor, q_thw, k_thw, rel_h: Optional[Tensor], rel_w: Optional[Tensor],
                 rel_t: Optional[Tensor]) -> Tensor:
    """Decomposed relative-position bias [B, h, Lq, Lk] of upstream PySlowFast's MViTv2 (cal_rel_pos_spatial /
    cal_rel_pos_temporal).  *** NOT in /root/reference: parity unpinned. ***  Restated from the formula in
    SURVEY.md Appendix F, in its dense form: gather R[a, b] = table[dist(a, b)], contract with the unscaled pooled q, and
    broadcast over the axes each term does not depend on.  q: [B, h, qt*qh*qw, d], tokens ordered (t, h, w); no cls."""
    B, nh, Lq, d = q.shape
    qt, qh, qw = q_thw
    kt, kh, kw = k_thw
    q6 = q.reshape(B, nh, qt, qh, qw, d)
    bias = q.new_zeros(B, nh, qt, qh, qw, kt, kh, kw)

    def gathered(table, nq, nk):
        q_ratio, k_ratio = max(nk / nq, 1.0), max(nq / nk, 1.0)
        dist = torch.arange(nq)[:, None] * q_ratio - torch.arange(nk)[None, :] * k_ratio + (nk - 1) * k_ratio
        assert table.shape[0] == 2 * max(nq, nk) - 1
        return table[dist.long()]                                    # [nq, nk, d]

    if rel_h is not None:
        bias = bias + torch.einsum("bnthwc,hkc->bnthwk", q6, gathered(rel_h, qh, kh))[:, :, :, :, :, None, :, None]
    if rel_w is not None:
        bias = bias + torch.einsum("bnthwc,wkc->bnthwk", q6, gathered(rel_w, qw, kw))[:, :, :, :, :, None, None, :]
    if rel_t is not None:
        bias = bias + torch.einsum("bnthwc,tkc->bnthwk", q6, gathered(rel_t, qt, kt))[:, :, :, :, :, :, None, None]
    return bias.reshape(B, nh, Lq, kt * kh * kw)


def multiscale_attention(x: Tensor, thw, sd: Dict[str, Tensor], pfx: str, spec: BlockSpec,
                         mvit: MViTSpec, return_parts: bool = False):
    B, N, _ = x.shape
    C, h = spec.attn_dim, spec.heads
    d = C // h
    qkv = F.linear(x, sd[pfx + "qkv.weight"], sd.get(pfx + "qkv.bias"))
    qkv = qkv.reshape(B, N, 3, h, d).permute(2, 0, 3, 1, 4)    # :231-236
    parts = []
    out_thw = list(thw)
    k_thw = list(thw)
    for name, t, kern, strd in (("q", qkv[0], spec.kernel_q, spec.stride_q),
                                ("k", qkv[1], spec.kernel_kv, spec.stride_kv),
                                ("v", qkv[2], spec.kernel_kv, spec.stride_kv)):
        pooled = bool(kern) and not (math.prod(kern) == 1 and math.prod(strd) == 1)   # :131-134
        if pooled:
            ln = None
            if mvit.mode == "conv":
                ln = (sd[pfx + f"norm_{name}.weight"], sd[pfx + f"norm_{name}.bias"], POOL_LN_EPS)
            t, t_thw = attention_pool(t, thw, mode=mvit.mode, kernel=kern, stride=strd,
                                      weight=sd.get(pfx + f"pool_{name}.weight"),
                                      has_cls=mvit.cls_embed_on, ln=ln)
            if name == "q":
                out_thw = t_thw
            elif name == "k":
                k_thw = t_thw
        parts.append(t)
    q, k, v = parts
    scale = d ** -0.5                                          # :118-119
    attn = (q @ k.transpose(-2, -1)) * scale                   # :267
    if any(pfx + n in sd for n in ("rel_pos_h", "rel_pos_w", "rel_pos_t")):
        # default-off extension, absent from the reference (see rel_pos_bias): added to the scaled scores
        attn = attn + rel_pos_bias(q, out_thw, k_thw, sd.get(pfx + "rel_pos_h"), sd.get(pfx + "rel_pos_w"),
                                   sd.get(pfx + "rel_pos_t"))
    attn = attn.softmax(dim=-1)                                # :269
    Lq = q.shape[2]
    y = (attn @ v).transpose(1, 2).reshape(B, Lq, C)           # :276
    if mvit.q_pool_residual:
        y = y + q.transpose(1, 2).reshape(B, Lq, C)            # :277-279
    out = F.linear(y, sd[pfx + "proj.weight"], sd[pfx + "proj.bias"])   # :281
    if return_parts:
        return out, out_thw, dict(q=q, k=k, v=v, y=y)
    return out, out_thw


# ----------------------------------------------------------------------------
# MultiScaleBlock.forward                              attention.py:412-446
# ----------------------------------------------------------------------------
def multiscale_block(x: Tensor, thw, sd: Dict[str, Tensor], pfx: str, spec: BlockSpec,
                     mvit: MViTSpec, drop_mask: Optional[Tuple[Tensor, Tensor]] = None):
    """`drop_mask=(m_attn, m_mlp)`: per-sample DropPath multipliers already divided by
    keep-prob (common.py:46-59); None = eval."""
    xn = F.layer_norm(x, (spec.dim,), sd[pfx + "norm1.weight"], sd[pfx + "norm1.bias"], BLOCK_LN_EPS)
    x_block, thw_new = multiscale_attention(xn, thw, sd, pfx + "attn.", spec, mvit)
    if spec.expand:
        x = F.linear(x, sd[pfx + "proj_max_pool.weight"], sd[pfx + "proj_max_pool.bias"])   # :424-426
    if spec.stride_q:                                          # pool_skip is None only for stride_q == []
        kskip = [s + 1 if s > 1 else s for s in spec.stride_q]                              # :316-318
        x_res, _ = attention_pool(x, thw, mode="max", kernel=kskip, stride=spec.stride_q,
                                  has_cls=mvit.cls_embed_on)
    else:
        x_res = x
    if drop_mask is not None:
        x_block = x_block * drop_mask[0]
    x = x_res + x_block                                        # :434
    c = spec.attn_dim
    xn2 = F.layer_norm(x, (c,), sd[pfx + "norm2.weight"], sd[pfx + "norm2.bias"], BLOCK_LN_EPS)
    hdn = F.gelu(F.linear(xn2, sd[pfx + "mlp.fc1.weight"], sd[pfx + "mlp.fc1.bias"]))       # common.py:27-28
    x_mlp = F.linear(hdn, sd[pfx + "mlp.fc2.weight"], sd[pfx + "mlp.fc2.bias"])
    if c != spec.dim_out:                                      # :441-443
        x = F.linear(xn2, sd[pfx + "proj.weight"], sd[pfx + "proj.bias"])
    if drop_mask is not None:
        x_mlp = x_mlp * drop_mask[1]
    return x + x_mlp, thw_new                                  # :445


# ----------------------------------------------------------------------------
# MViT.forward                                         video_model_builder.py:1161-1335
# ----------------------------------------------------------------------------
def mvit_forward(x: Tensor, sd: Dict[str, Tensor], mvit: MViTSpec, *, training: bool = False,
                 return_features: bool = False):
    """x: [B, 3, T, S, S] clip.  Eval returns softmax probabilities (head_helper.py:409-417)."""
    x = F.conv3d(x, sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"],
                 stride=tuple(mvit.patch_stride), padding=tuple(mvit.patch_padding))
    x = x.flatten(2).transpose(1, 2)                           # stem_helper.py:336-338
    B = x.shape[0]
    T, H, W = mvit.patch_dims
    if mvit.cls_embed_on:
        x = torch.cat((sd["cls_token"].expand(B, -1, -1), x), dim=1)
    if mvit.sep_pos_embed:                                     # :1196-1223
        pos = sd["pos_embed_spatial"].repeat(1, T, 1) + torch.repeat_interleave(
            sd["pos_embed_temporal"], H * W, dim=1)
        if mvit.cls_embed_on:
            pos = torch.cat([sd["pos_embed_class"], pos], 1)
        x = x + pos
    else:
        x = x + sd["pos_embed"]
    thw = [T, H, W]
    for i, spec in enumerate(mvit.blocks):
        x, thw = multiscale_block(x, thw, sd, f"blocks.{i}.", spec, mvit)
    if mvit.final_norm:
        x = F.layer_norm(x, (x.shape[-1],), sd["norm.weight"], sd["norm.bias"], BLOCK_LN_EPS)
    feat = x[:, 0] if mvit.cls_embed_on else x.mean(1)         # :1305-1310
    logits = F.linear(feat, sd["head.projection.weight"], sd["head.projection.bias"])
    out = logits
    if not training:
        out = logits.softmax(dim=1) if mvit.head_act == "softmax" else logits.sigmoid()
    if return_features:
        return out, dict(feat=feat, logits=logits, thw=thw)
    return out
