#!/usr/bin/env python3
"""A/B of the folded-LayerNorm GEMM forms against the plain ones at the MViTv2-B @448 batch-8 stage shapes.

    python tools/lnfold_bench.py [out.json]
Per stage: qkv / fc1 as plain GEMM, plain GEMM + LayerNorm launch, and the LN-folded consumer; proj / fc2 (+residual)
with and without the row-statistics output.  L2 flushed between launches."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aicity_action_b200 import ops  # noqa: E402
from aicity_action_b200.weights import folded_ln_linear  # noqa: E402

_flush = None


def timed(fn, iters=8, warmup=3):
    global _flush
    if _flush is None:
        _flush = torch.empty(160 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    total = 0.0
    for _ in range(iters):
        _flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        total += e0.elapsed_time(e1)
    return total / iters


def main():
    dt = torch.bfloat16
    B = 8
    rows = []
    # (name, tokens per clip, C, hidden)
    for name, L, C in (("blk0", 100352, 96), ("blk2", 25088, 192), ("blk4", 6272, 384), ("blk15", 1568, 768)):
        M = B * L
        x = torch.randn(M, C, device="cuda", dtype=dt)
        res = torch.randn(M, C, device="cuda", dtype=dt)
        g = torch.nn.Parameter(torch.rand(C, device="cuda") + 0.5)
        be = torch.nn.Parameter(torch.randn(C, device="cuda") * 0.1)
        for lname, N, gelu in (("qkv", 3 * C, False), ("fc1", 4 * C, True)):
            w = torch.nn.Parameter(torch.randn(N, C, device="cuda") * C ** -0.5)
            b = torch.nn.Parameter(torch.randn(N, device="cuda") * 0.1)
            w16 = w.detach().to(dt)
            xs = ops.linear_stats(x, torch.eye(C, device="cuda", dtype=dt), None, residual=res)
            st = ops.row_stats_of(xs)
            wf, bf, cs = folded_ln_linear(w, b, g, be)
            t_plain = timed(lambda: ops.linear(xs, w16, b.detach(), gelu=gelu))
            t_ln = timed(lambda: ops.layernorm(xs, g.detach(), be.detach(), 1e-6))
            t_fold = timed(lambda: ops.linear_ln(xs, st, wf, bf, cs, 1e-6, gelu=gelu))
            rows.append(dict(shape=f"{name} {lname} M={M} N={N} K={C}", plain_ms=t_plain, layernorm_ms=t_ln, folded_ms=t_fold))
        if ops.mlp_fused_supported(C, 4 * C, C):
            # the whole MLP: folded fc1 + GELU, fc2 + residual + statistics as two launches against the single fused kernel
            w1 = torch.nn.Parameter(torch.randn(4 * C, C, device="cuda") * C ** -0.5)
            b1 = torch.nn.Parameter(torch.randn(4 * C, device="cuda") * 0.1)
            w2 = (torch.randn(C, 4 * C, device="cuda") * (4 * C) ** -0.5).to(dt)
            b2 = torch.randn(C, device="cuda") * 0.1
            xs = ops.linear_stats(x, torch.eye(C, device="cuda", dtype=dt), None, residual=res)
            st = ops.row_stats_of(xs)
            wf, bf, cs = folded_ln_linear(w1, b1, g, be)
            t_two = timed(lambda: ops.linear_stats(ops.linear_ln(xs, st, wf, bf, cs, 1e-6, gelu=True), w2, b2, residual=xs))
            t_fused = timed(lambda: ops.mlp_fused(xs, st, wf, bf, cs, w2, b2, 1e-6))
            rows.append(dict(shape=f"{name} mlp M={M} C={C}", two_launch_ms=t_two, fused_ms=t_fused))
        for lname, K in (("proj", C), ("fc2", 4 * C)):
            h = torch.randn(M, K, device="cuda", dtype=dt)
            w16 = (torch.randn(C, K, device="cuda") * K ** -0.5).to(dt)
            b = torch.randn(C, device="cuda") * 0.1
            t_plain = timed(lambda: ops.linear(h, w16, b, residual=res))
            t_stats = timed(lambda: ops.linear_stats(h, w16, b, residual=res))
            rows.append(dict(shape=f"{name} {lname}+res M={M} N={C} K={K}", plain_ms=t_plain, stats_ms=t_stats))
    for r in rows:
        print("  ".join(f"{k}={v:.4f}" if isinstance(v, float) else f"{v:44s}" for k, v in r.items()))
    if len(sys.argv) > 1:
        json.dump(rows, open(sys.argv[1], "w"), indent=1)


if __name__ == "__main__":
    main()
