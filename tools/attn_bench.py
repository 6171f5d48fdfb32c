"""Forward / backward timing of the fused attention at the four MViTv2-B @448 stage shapes (batch 8)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aicity_action_b200 import ops  # noqa: E402

B = int(os.environ.get("B", "8"))
reps = int(os.environ.get("REPS", "3"))
flush = torch.empty(256 * 2 ** 20, dtype=torch.uint8, device="cuda")
for name, h, Lq, Lk, blocks in (("stage1", 1, 100352, 1568, 1), ("block1", 2, 25088, 6272, 1), ("stage2", 2, 25088, 1568, 2),
                                ("stage3", 4, 6272, 1568, 16), ("stage4", 8, 1568, 1568, 3)):
    q = torch.randn(B, h, Lq, 96, device="cuda", dtype=torch.bfloat16)
    k = torch.randn(B, h, Lk, 96, device="cuda", dtype=torch.bfloat16)
    v = torch.randn(B, h, Lk, 96, device="cuda", dtype=torch.bfloat16)
    do = torch.randn(B, Lq, h * 96, device="cuda", dtype=torch.bfloat16)
    out, lse = ops.attention(q, k, v, 96 ** -0.5, True, want_lse=True)
    tf, tb = [], []
    for _ in range(reps):
        flush.zero_()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        ops.attention(q, k, v, 96 ** -0.5, True, want_lse=True)
        e[1].record()
        ops.attention_bwd(q, k, v, out, do, lse, 96 ** -0.5, True)
        e[2].record()
        torch.cuda.synchronize()
        tf.append(e[0].elapsed_time(e[1]))
        tb.append(e[1].elapsed_time(e[2]))
    f, b = sorted(tf)[reps // 2], sorted(tb)[reps // 2]
    fl = 4.0 * B * h * Lq * Lk * 96
    print(f"{name} h={h} Lq={Lq} Lk={Lk}: fwd {f:.3f} ms {fl / f / 1e9:.0f} TF/s | bwd {b:.3f} ms "
          f"{2.5 * fl / b / 1e9:.0f} TF/s algorithmic ({3.5 * fl / b / 1e9:.0f} executed)  x{blocks}")
