#!/usr/bin/env python3
"""Diagnostic driver for the tcgen05 kernels (run on the GPU box): compares them with fp32 torch math on
the device, shape by shape, each family in its own process so a trap in one does not mask the other.

    python tools/tc_check.py gemm | attn | all
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def rel(a, b):
    return float((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-30))


def describe_err(got, ref, rows=128, cols=32):
    """Coarse error map: max-abs error per (row-block, col-block)."""
    import torch
    e = (got.float() - ref.float()).abs()
    M, N = e.shape
    out = []
    for r0 in range(0, min(M, 4 * rows), rows):
        out.append(" ".join(f"{float(e[r0:r0 + rows, c0:c0 + cols].max()):8.2e}" for c0 in range(0, min(N, 8 * cols), cols)))
    return "\n      ".join(out)


def gemm():
    import torch
    from aicity_action_b200 import ops
    from aicity_action_b200._lib import IMPL_TCGEN05
    torch.manual_seed(0)
    bad = 0
    shapes = [(128, 96, 64), (128, 96, 96), (128, 192, 128), (256, 192, 192), (300, 288, 96), (1000, 576, 192),
              (6272, 1152, 384), (6272, 384, 1536), (25088, 192, 96), (12544, 3072, 768), (12544, 768, 3072),
              (50, 2304, 768), (100352, 288, 96)]
    for (M, N, K) in shapes:
        x = torch.randn(M, K, device="cuda").bfloat16()
        w = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16()
        b = torch.randn(N, device="cuda") * 0.1
        r = torch.randn(M, N, device="cuda").bfloat16()
        ref = x.float() @ w.float().t() + b
        for name, kw, refv in (("bias", {}, ref), ("gelu", {"gelu": True}, torch.nn.functional.gelu(ref)),
                               ("res", {"residual": r}, ref + r.float())):
            got = ops.linear(x, w, b, impl=IMPL_TCGEN05, **kw)
            torch.cuda.synchronize()
            e = rel(got, refv)
            ok = e < 1e-2
            bad += not ok
            print(f"gemm M={M:6d} N={N:4d} K={K:4d} {name:5s} rel={e:.3e} {'ok' if ok else 'FAIL'}", flush=True)
            if not ok:
                print("      " + describe_err(got, refv), flush=True)
    return bad


def attn():
    import torch
    from aicity_action_b200 import ops
    from aicity_action_b200._lib import IMPL_TCGEN05
    torch.manual_seed(0)
    bad = 0
    shapes = [(1, 1, 128, 128), (1, 1, 256, 128), (1, 1, 256, 256), (1, 2, 256, 512), (2, 2, 300, 200), (1, 4, 72, 72),
              (1, 1, 1024, 16), (1, 8, 1568, 1568), (1, 2, 6272, 392), (2, 1, 25088, 1568), (1, 4, 6272, 6272)]
    for (B, h, Lq, Lk) in shapes:
        q = torch.randn(B, h, Lq, 96, device="cuda").bfloat16()
        k = torch.randn(B, h, Lk, 96, device="cuda").bfloat16()
        v = torch.randn(B, h, Lk, 96, device="cuda").bfloat16()
        scale = 96 ** -0.5
        o = torch.nn.functional.scaled_dot_product_attention(q.float(), k.float(), v.float(), scale=scale)
        for add_q in (False, True):
            ref = (o + q.float() if add_q else o).transpose(1, 2).reshape(B, Lq, h * 96)
            got, lse = ops.attention(q, k, v, scale, add_q, want_lse=True, impl=IMPL_TCGEN05)
            torch.cuda.synchronize()
            e = rel(got, ref)
            ok = e < 2e-2
            bad += not ok
            print(f"attn B={B} h={h} Lq={Lq:6d} Lk={Lk:5d} add_q={int(add_q)} rel={e:.3e} {'ok' if ok else 'FAIL'}", flush=True)
            if not ok:
                print("      " + describe_err(got[0], ref[0]), flush=True)
    return bad


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what == "all":
        rc = 0
        for w in ("gemm", "attn"):
            r = subprocess.run([sys.executable, os.path.abspath(__file__), w], timeout=600)
            rc |= r.returncode
        sys.exit(rc)
    sys.exit(1 if {"gemm": gemm, "attn": attn}[what]() else 0)
