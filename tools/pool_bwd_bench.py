"""Per-shape timing of the attention_pool backward pieces at the MViTv2-B @448 shapes (batch 8, bf16)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aicity_action_b200 import ops  # noqa: E402

B, d = 8, 96
flush = torch.empty(256 * 2 ** 20, dtype=torch.uint8, device="cuda")


def timed(fn, reps=3):
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


# (name, heads, thw, stride)
shapes = [("blk0 q s1", 1, (8, 112, 112), (1, 1, 1)), ("blk0 kv s8", 1, (8, 112, 112), (1, 8, 8)),
          ("blk1 q s2", 2, (8, 112, 112), (1, 2, 2)), ("blk1 kv s4", 2, (8, 112, 112), (1, 4, 4)),
          ("blk2 q s1", 2, (8, 56, 56), (1, 1, 1)), ("blk2 kv s4", 2, (8, 56, 56), (1, 4, 4)),
          ("blk3 q s2", 4, (8, 56, 56), (1, 2, 2)), ("blk4 q s1", 4, (8, 28, 28), (1, 1, 1)),
          ("blk4 kv s2", 4, (8, 28, 28), (1, 2, 2)), ("blk14 q s2", 8, (8, 28, 28), (1, 2, 2)),
          ("blk15 q s1", 8, (8, 14, 14), (1, 1, 1))]
for name, h, thw, st in shapes:
    N = thw[0] * thw[1] * thw[2]
    qkv = torch.randn(B, N, 3 * h * d, device="cuda", dtype=torch.bfloat16)
    dqkv = torch.empty_like(qkv)
    w = torch.randn(d, 27, device="cuda")
    oth = ops.pooled_thw(list(thw), [3, 3, 3], list(st))
    Lo = oth[0] * oth[1] * oth[2]
    dy = torch.randn(B, h, Lo, d, device="cuda", dtype=torch.bfloat16)
    strides = (N * 3 * h * d, 3 * h * d, d)
    x_view = qkv.view(B, N, 3, h, d)[:, :, 0]
    dx_view = dqkv.view(B, N, 3, h, d)[:, :, 0]
    dw = torch.zeros(d, 27, device="cuda")
    t_w = timed(lambda: ops.attention_pool_bwd(1, x_view, strides, dy, None, None, dw, B, h, d, list(thw), [3, 3, 3], list(st)))
    if st == (1, 1, 1):
        t_d = timed(lambda: ops.attention_pool_strided(dy, 0, (h * Lo * d, d, Lo * d), B, h, d, list(thw), [3, 3, 3], list(st),
                                                       "conv", w, None, None, 0.0, False, dx_view, strides))
    else:
        t_d = timed(lambda: ops.attention_pool_bwd(0, None, strides, dy, w, dx_view, None, B, h, d, list(thw), [3, 3, 3],
                                                   list(st)))
    t_f = timed(lambda: ops.attention_pool_heads(x_view.permute(0, 2, 1, 3), list(thw), [3, 3, 3], list(st), mode="conv", weight=w))
    mb_in, mb_out = B * N * h * d * 2 / 1e6, B * Lo * h * d * 2 / 1e6
    print(f"{name:11s} in {mb_in:6.1f} MB out {mb_out:6.1f} MB | fwd(noLN) {t_f * 1e3:7.1f} us  wgrad {t_w * 1e3:7.1f} us  "
          f"dgrad {t_d * 1e3:7.1f} us ({(mb_in + mb_out) / t_d / 1e3:5.2f} TB/s)")
