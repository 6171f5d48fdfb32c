#!/usr/bin/env python3
"""Does running the batch as TWO half-batches on two streams (one captured graph, fork / join) beat one batch-8 forward?

The forward is a strict chain of ~170 kernels, each with a ramp and a tail wave; two independent chains could fill each
other's tails.  Prints ms per batch-8 step for: one batch-8 graph, two batch-4 chains on two streams, two batch-4 chains
back to back on one stream.
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aicity_action_b200.config import aicity_cfg  # noqa: E402
from aicity_action_b200.graphed import GraphedForward  # noqa: E402
from aicity_action_b200.mvit import MViT  # noqa: E402


def timed(fn, steps=20, heat=2.0):
    t_end = time.perf_counter() + heat
    while time.perf_counter() < t_end:
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    cfg = aicity_cfg("MVITV2_FULL_B_16x4_CONV_448")
    torch.manual_seed(0)
    m = MViT(cfg).eval().cuda()
    B, T, S = 8, cfg.DATA.NUM_FRAMES, cfg.DATA.TRAIN_CROP_SIZE
    frames = torch.randint(0, 256, (B, T, S, S, 3), dtype=torch.uint8).cuda()
    with torch.no_grad():
        g8 = GraphedForward(m, frames)
        ref = g8().clone()
        print("one batch-8 graph      ", round(timed(g8), 3), "ms", flush=True)

        halves = [frames[:4].contiguous(), frames[4:].contiguous()]
        for h in halves:
            for _ in range(2):
                m([h])
        torch.cuda.synchronize()
        for mode in ("two streams", "one stream"):
            s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
            graph = torch.cuda.CUDAGraph()
            outs = [None, None]
            with torch.cuda.graph(graph):
                cur = torch.cuda.current_stream()
                if mode == "two streams":
                    for n, st in enumerate((s1, s2)):
                        st.wait_stream(cur)
                        with torch.cuda.stream(st):
                            outs[n] = m([halves[n]])
                    cur.wait_stream(s1)
                    cur.wait_stream(s2)
                else:
                    for n in range(2):
                        outs[n] = m([halves[n]])
            graph.replay()
            torch.cuda.synchronize()
            same = torch.equal(torch.cat(outs), ref)
            print(f"two batch-4 chains, {mode:11s}", round(timed(graph.replay), 3), "ms   equal to batch-8:", same, flush=True)


if __name__ == "__main__":
    main()
