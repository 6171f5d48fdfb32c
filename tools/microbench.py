#!/usr/bin/env python3
"""Per-kernel microbenchmark over the MViTv2-B 16x4 stage shapes (BASELINE config 5).

    python tools/microbench.py [--size 448] [--batch 8] [--what pool,attn,gemm,ln] [--json out.json]

Every kernel is timed alone with CUDA events (3 warm-up + N timed launches, an L2-flushing write between
launches), and reported against the roofline that bounds it: attention_pool / LayerNorm in GB/s of
ALGORITHMIC bytes vs the measured HBM peak, attention / Linear in TFLOP/s vs the measured bf16 peak.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from aicity_action_b200 import ops  # noqa: E402
from aicity_action_b200.config import aicity_cfg  # noqa: E402
from aicity_action_b200.mvit import MViT  # noqa: E402


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d["bf16_tflops"]
    return 6650.0, 1590.0


_flush = None


def timeit(fn, iters=10, flush=True):
    global _flush
    if _flush is None:
        _flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        if flush:
            _flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters


def stage_shapes(size):
    """(blk, thw, Cin, C, heads, stride_q, stride_kv) for every distinct block of MVITV2_FULL_B_16x4."""
    cfg = aicity_cfg("MVITV2_FULL_B_16x4_CONV_448" if size == 448 else "MVITV2_FULL_B_16x4_CONV")
    m = MViT(cfg)
    thw = list(m.patch_dims)
    out, seen = [], set()
    for i, blk in enumerate(m.blocks):
        a = blk.attn
        sq = list(a.pool_q.stride) if a.pool_q is not None else [1, 1, 1]
        skv = list(a.pool_k.stride) if a.pool_k is not None else [1, 1, 1]
        cin, c = a.qkv.in_features, a.dim_out
        key = (tuple(thw), cin, c, a.num_heads, tuple(sq), tuple(skv))
        if key not in seen:
            seen.add(key)
            out.append(dict(blk=i, thw=list(thw), cin=cin, c=c, heads=a.num_heads, sq=sq, skv=skv))
        thw = ops.pooled_thw(thw, [3, 3, 3], sq)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=448)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--what", default="pool,attn,gemm,ln")
    ap.add_argument("--json", default="")
    ap.add_argument("--iters", type=int, default=10)
    args = ap.parse_args()
    what = set(args.what.split(","))
    hbm, tf = peaks()
    B = args.batch
    dt = torch.bfloat16
    rows = []

    def rec(kind, name, ms, work, bound):
        ach = work / (ms * 1e-3) / (1e9 if bound == "hbm" else 1e12)
        peak = hbm if bound == "hbm" else tf
        rows.append(dict(kernel=kind, shape=name, ms=ms, achieved=ach, unit="GB/s" if bound == "hbm" else "TFLOP/s",
                         frac=ach / peak))
        print(f"{kind:9s} {name:58s} {ms:8.3f} ms  {ach:9.1f} {'GB/s' if bound == 'hbm' else 'TF/s'}  {100 * ach / peak:5.1f}%",
              flush=True)

    for s in stage_shapes(args.size):
        T, H, W = s["thw"]
        N = T * H * W
        C, h = s["c"], s["heads"]
        tag = f"blk{s['blk']} thw={T}x{H}x{W} C={s['cin']}->{C} h={h}"
        if "pool" in what:
            qkv = torch.randn(B, N, 3, h, 96, device="cuda", dtype=dt)
            w = torch.randn(96, 27, device="cuda")
            g, b = torch.ones(96, device="cuda"), torch.zeros(96, device="cuda")
            for which, st in ((0, s["sq"]), (1, s["skv"])):
                view = qkv[:, :, which].permute(0, 2, 1, 3)
                Lo = math.prod(ops.pooled_thw(s["thw"], [3, 3, 3], st))
                work = (B * h * N * 96 + B * h * Lo * 96) * 2.0
                ms = timeit(lambda: ops.attention_pool_heads(view, s["thw"], [3, 3, 3], st, mode="conv", weight=w,
                                                             ln=(g, b, 1e-5)), args.iters)
                rec("pool", f"{tag} {'q' if which == 0 else 'k/v'} s={st}", ms, work, "hbm")
            # the block's three pools in one call (mvit_attention_pool_qkv_fwd): all of qkv read + q, k, v written
            q2 = qkv.view(B, N, 3 * h * 96)
            st3 = [tuple(s["sq"]), tuple(s["skv"]), tuple(s["skv"])]
            if ops.pool_qkv_supported(q2, h, st3):
                w5 = w.view(96, 1, 3, 3, 3)
                lo = [math.prod(ops.pooled_thw(s["thw"], [3, 3, 3], list(x))) for x in st3]
                work = (3 * B * h * N * 96 + B * h * sum(lo) * 96) * 2.0
                ms = timeit(lambda: ops.attention_pool_qkv(q2, h, list(s["thw"]), [w5] * 3, [(g, b, 1e-5)] * 3, st3), args.iters)
                rec("pool_qkv", f"{tag} q+k+v s_q={s['sq'][1]} s_kv={s['skv'][1]}", ms, work, "hbm")
            del qkv
        if "attn" in what:
            Lq = math.prod(ops.pooled_thw(s["thw"], [3, 3, 3], s["sq"]))
            Lk = math.prod(ops.pooled_thw(s["thw"], [3, 3, 3], s["skv"]))
            q = torch.randn(B, h, Lq, 96, device="cuda", dtype=dt)
            k = torch.randn(B, h, Lk, 96, device="cuda", dtype=dt)
            v = torch.randn(B, h, Lk, 96, device="cuda", dtype=dt)
            ms = timeit(lambda: ops.attention(q, k, v, 96 ** -0.5, True), args.iters)
            rec("attention", f"{tag} Lq={Lq} Lk={Lk}", ms, 4.0 * B * h * Lq * Lk * 96, "tensor")
            del q, k, v
        if "gemm" in what:
            Lq = math.prod(ops.pooled_thw(s["thw"], [3, 3, 3], s["sq"]))
            gemms = [("qkv", B * N, 3 * C, s["cin"], False, False), ("proj+res", B * Lq, C, C, False, True),
                     ("fc1+gelu", B * Lq, 4 * C, C, True, False), ("fc2+res", B * Lq, C, 4 * C, False, True)]
            if s["cin"] != C:
                gemms.append(("proj_max_pool", B * N, C, s["cin"], False, False))
            for name, M, Nn, K, gelu, res in gemms:
                x = torch.randn(M, K, device="cuda", dtype=dt)
                wt = torch.randn(Nn, K, device="cuda", dtype=dt) * K ** -0.5
                bias = torch.randn(Nn, device="cuda")
                r = torch.randn(M, Nn, device="cuda", dtype=dt) if res else None
                y = torch.empty(M, Nn, device="cuda", dtype=dt)
                ms = timeit(lambda: ops.linear(x, wt, bias, residual=r, gelu=gelu, out=y), args.iters)
                byt = 2.0 * (M * K + Nn * K + M * Nn * (2 if res else 1))
                rec("linear", f"{tag} {name} M={M} N={Nn} K={K}  [{byt / (ms * 1e-3) / 1e9:6.0f} GB/s]", ms,
                    2.0 * M * Nn * K, "tensor")
                del x, wt, r, y
        if "ln" in what:
            x = torch.randn(B * N, s["cin"], device="cuda", dtype=dt)
            g, b = torch.ones(s["cin"], device="cuda"), torch.zeros(s["cin"], device="cuda")
            y = torch.empty_like(x)
            ms = timeit(lambda: ops.layernorm(x, g, b, 1e-6, out=y), args.iters)
            rec("layernorm", f"{tag} rows={B * N} C={s['cin']}", ms, 2.0 * x.numel() * 2, "hbm")
            del x, y
    if args.json:
        with open(args.json, "w") as f:
            json.dump({"size": args.size, "batch": B, "hbm_gbs_peak": hbm, "bf16_tflops_peak": tf, "rows": rows}, f, indent=1)


if __name__ == "__main__":
    main()
