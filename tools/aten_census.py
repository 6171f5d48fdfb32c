#!/usr/bin/env python3
"""Which kernels of one eager inference forward are NOT from libmvit_b200.so (ATen copies, fills, casts)?

    python tools/aten_census.py [--size 448] [--batch 8]      -> gpurun_out/aten_census.txt
Uses the torch profiler (CUPTI) on one forward after warm-up and groups device kernels by name."""
import argparse
import collections
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from aicity_action_b200.config import aicity_cfg  # noqa: E402
from aicity_action_b200.mvit import MViT  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=448)
ap.add_argument("--batch", type=int, default=8)
a = ap.parse_args()
cfg = aicity_cfg("MVITV2_FULL_B_16x4_CONV_448" if a.size == 448 else "MVITV2_FULL_B_16x4_CONV")
torch.manual_seed(0)
model = MViT(cfg).cuda().eval()
x = torch.randint(0, 256, (a.batch, cfg.DATA.NUM_FRAMES, a.size, a.size, 3), dtype=torch.uint8, device="cuda")
with torch.no_grad():
    for _ in range(3):
        model([x])
    torch.cuda.synchronize()
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA, torch.profiler.ProfilerActivity.CPU],
                                with_stack=True) as prof:
        model([x])
        torch.cuda.synchronize()
ours, other = collections.Counter(), collections.Counter()
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        (ours if "mvit" in ev.name else other)[ev.name[:110]] += 1
lines = [f"{sum(ours.values())} kernels from libmvit_b200.so, {sum(other.values())} others per forward"]
lines += [f"  {n:4d}  {k}" for k, n in other.most_common()]
# CPU-side ops that launched the foreign kernels, with the Python frame that issued them
lines.append("aten ops (CPU side) with a device kernel:")
for ev in prof.key_averages(group_by_stack_n=6):
    if ev.key.startswith("aten::") and ev.device_time_total > 0 and ev.key not in ("aten::empty", "aten::empty_like"):
        stack = [s for s in ev.stack if "aicity_action_b200" in s][:2]
        lines.append(f"  {ev.count:4d}  {ev.key:28s} {' <- '.join(s.strip()[-90:] for s in stack)}")
out = "\n".join(lines)
print(out)
d = os.path.join(ROOT, "gpurun_out")
if os.path.isdir(d):
    open(os.path.join(d, "aten_census.txt"), "w").write(out + "\n")
