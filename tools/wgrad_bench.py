"""Per-shape timing of the Linear weight-gradient kernel (the MViTv2-B @448 layer shapes at batch 8)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aicity_action_b200 import ops  # noqa: E402

B = 8
shapes = []
for tokens, C, blocks in ((100352, 96, 1), (25088, 192, 3), (6272, 384, 16), (1568, 768, 3)):
    M = B * tokens
    shapes += [("qkv", M, 3 * C, C, blocks), ("proj", M, C, C, blocks), ("fc1", M, 4 * C, C, blocks),
               ("fc2", M, C, 4 * C, blocks)]
flush = torch.empty(256 * 2 ** 20, dtype=torch.uint8, device="cuda")
tot = 0.0
for name, M, N, K, blocks in shapes:
    dy = torch.randn(M, N, device="cuda", dtype=torch.bfloat16)
    x = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
    for bias in (False, True):
        ts = []
        for _ in range(4):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.linear_wgrad(dy, x, bias)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[1]
        print(f"{name:5s} M={M:7d} N={N:5d} K={K:5d} bias={int(bias)}  {ms:7.3f} ms  {2.0 * M * N * K / ms / 1e9:7.1f} TFLOP/s  x{blocks}")
        if bias:
            tot += ms * blocks
print(f"sum over the model (with bias): {tot:.2f} ms")
