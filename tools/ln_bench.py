"""LayerNorm forward / backward timing at the MViTv2-B @448 shapes (batch 8, bf16)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aicity_action_b200 import ops  # noqa: E402

flush = torch.empty(256 * 2 ** 20, dtype=torch.uint8, device="cuda")


def timed(fn, reps=5):
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[reps // 2]


for rows, C in ((802816, 96), (200704, 192), (50176, 384), (12544, 768), (200704, 96), (50176, 96)):
    x = torch.randn(rows, C, device="cuda", dtype=torch.bfloat16)
    dy = torch.randn_like(x)
    g, b = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
    tf = timed(lambda: ops.layernorm(x, g, b, 1e-6))
    tb = timed(lambda: ops.layernorm_bwd(x, g, dy, 1e-6))
    mb = rows * C * 2 / 1e6
    print(f"rows {rows:7d} C {C:4d}: fwd {tf * 1e3:7.1f} us ({2 * mb / tf / 1e3:5.2f} TB/s)   bwd {tb * 1e3:7.1f} us ({3 * mb / tb / 1e3:5.2f} TB/s)")
