#!/usr/bin/env python3
"""Launch ONE representative shape of a kernel family a few times (for `ncu -k regex:... -s N -c 1`).

    python tools/profile_one.py attn|attn0|gemm_big|gemm_gelu|gemm_res|pool_q|pool_kv|pool_qkv0|pool_qkv4|mlp96|mlp192|ln
Shapes are MViTv2-B 16x4@448 batch-8 stage shapes.
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aicity_action_b200 import ops  # noqa: E402

what = sys.argv[1]
dt = torch.bfloat16
B = 8
torch.manual_seed(0)
if what == "attn":          # block 1: h=2, Lq=25088, Lk=6272
    q = torch.randn(B, 2, 25088, 96, device="cuda", dtype=dt)
    k = torch.randn(B, 2, 6272, 96, device="cuda", dtype=dt)
    v = torch.randn(B, 2, 6272, 96, device="cuda", dtype=dt)
    fn = lambda: ops.attention(q, k, v, 96 ** -0.5, True)
elif what == "attn0":       # block 0: h=1, Lq=100352, Lk=1568
    q = torch.randn(B, 1, 100352, 96, device="cuda", dtype=dt)
    k = torch.randn(B, 1, 1568, 96, device="cuda", dtype=dt)
    v = torch.randn(B, 1, 1568, 96, device="cuda", dtype=dt)
    fn = lambda: ops.attention(q, k, v, 96 ** -0.5, True)
elif what in ("gemm_gelu", "gemm_res", "gemm_big"):
    M, N, K = {"gemm_gelu": (802816, 384, 96), "gemm_res": (802816, 96, 384), "gemm_big": (12544, 3072, 768)}[what]
    x = torch.randn(M, K, device="cuda", dtype=dt)
    w = torch.randn(N, K, device="cuda", dtype=dt) * K ** -0.5
    b = torch.randn(N, device="cuda")
    r = torch.randn(M, N, device="cuda", dtype=dt) if what == "gemm_res" else None
    y = torch.empty(M, N, device="cuda", dtype=dt)
    fn = lambda: ops.linear(x, w, b, residual=r, gelu=what != "gemm_res", out=y)
elif what in ("pool_q", "pool_kv"):
    qkv = torch.randn(B, 8 * 112 * 112, 3, 1, 96, device="cuda", dtype=dt)
    w = torch.randn(96, 27, device="cuda")
    g, bb = torch.ones(96, device="cuda"), torch.zeros(96, device="cuda")
    which, st = (0, [1, 1, 1]) if what == "pool_q" else (1, [1, 8, 8])
    view = qkv[:, :, which].permute(0, 2, 1, 3)
    fn = lambda: ops.attention_pool_heads(view, [8, 112, 112], [3, 3, 3], st, mode="conv", weight=w, ln=(g, bb, 1e-5))
elif what in ("pool_qkv0", "pool_qkv4"):       # the fused q+k+v call at the block-0 / block-4 shapes
    thw, h, sq, skv = ([8, 112, 112], 1, 1, 8) if what == "pool_qkv0" else ([8, 28, 28], 4, 1, 2)
    N = thw[0] * thw[1] * thw[2]
    qkv = torch.randn(B, N, 3 * h * 96, device="cuda", dtype=dt)
    w = torch.randn(96, 1, 3, 3, 3, device="cuda")
    g, bb = torch.ones(96, device="cuda"), torch.zeros(96, device="cuda")
    st3 = [(1, sq, sq), (1, skv, skv), (1, skv, skv)]
    fn = lambda: ops.attention_pool_qkv(qkv, h, thw, [w] * 3, [(g, bb, 1e-5)] * 3, st3)
elif what in ("mlp96", "mlp192"):              # the fused MLP of block 0 / blocks 1-2
    from aicity_action_b200.weights import folded_ln_linear
    Cc, M = (96, 802816) if what == "mlp96" else (192, 200704)
    base = torch.randn(M, Cc, device="cuda", dtype=dt)
    x = ops.linear_stats(base, torch.eye(Cc, device="cuda", dtype=dt), None)
    stats = ops.row_stats_of(x)
    w1, b1 = torch.randn(4 * Cc, Cc, device="cuda") * Cc ** -0.5, torch.randn(4 * Cc, device="cuda") * 0.1
    w2 = (torch.randn(Cc, 4 * Cc, device="cuda") * (4 * Cc) ** -0.5).to(dt)
    b2 = torch.randn(Cc, device="cuda") * 0.1
    wf, bf, cs = folded_ln_linear(torch.nn.Parameter(w1), torch.nn.Parameter(b1), torch.nn.Parameter(torch.ones(Cc, device="cuda")),
                                  torch.nn.Parameter(torch.zeros(Cc, device="cuda")))
    fn = lambda: ops.mlp_fused(x, stats, wf, bf, cs, w2, b2, 1e-6)
elif what == "ln":
    x = torch.randn(802816, 96, device="cuda", dtype=dt)
    g, bb = torch.ones(96, device="cuda"), torch.zeros(96, device="cuda")
    fn = lambda: ops.layernorm(x, g, bb, 1e-6)
else:
    raise SystemExit(f"unknown target {what}")
for _ in range(3):
    fn()
torch.cuda.synchronize()
