#!/usr/bin/env python3
"""BASELINE config 3: sliding-window temporal localisation over synthetic 3-view driver videos, windows sharded
across the ranks of one box (torchrun --nproc-per-node N), scores gathered with one NCCL all_gather per video.

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/sliding_window_bench.py \
        [--frames 18000] [--views 3] [--size 448] [--batch 8] [--check]
Prints one JSON line (windows/s whole job, device-timed, max over ranks).  --check re-runs video 0 on rank 0
alone and verifies the sharded result is identical.
"""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from aicity_action_b200 import sliding_window as SW  # noqa: E402
from aicity_action_b200.config import aicity_cfg  # noqa: E402
from aicity_action_b200.mvit import MViT  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=18000)
    ap.add_argument("--views", type=int, default=3)
    ap.add_argument("--size", type=int, default=448)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--no-cuda-graph", action="store_true")
    a = ap.parse_args()
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", 1), ("RANK", 0), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    torch.set_num_threads(max(1, (os.cpu_count() or 1) // world))   # torchrun pins OMP to 1 thread; frame gathering is host work
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = aicity_cfg("MVITV2_FULL_B_16x4_CONV_448" if a.size == 448 else "MVITV2_FULL_B_16x4_CONV")
    torch.manual_seed(0)
    model = MViT(cfg).eval().to(dev)
    runner = SW.SlidingWindowRunner(model, batch_size=a.batch, device=dev, rank=rank, world=world,
                                    use_cuda_graph=not a.no_cuda_graph)
    videos = [SW.SyntheticVideo(seed=100 + v, num_frames=a.frames, size=a.size) for v in range(a.views)]
    runner.run_video(SW.SyntheticVideo(seed=1, num_frames=16 * a.batch * world, size=a.size), cfg.MODEL.NUM_CLASSES)  # warm-up
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    results = [runner.run_video(v, cfg.MODEL.NUM_CLASSES) for v in videos]
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    n_win = sum(len(r) for r in results)
    ok = None
    if a.check and rank == 0:
        solo = SW.SlidingWindowRunner(model, batch_size=a.batch, device=dev).run_video(videos[0], cfg.MODEL.NUM_CLASSES)
        ok = all(x[0] == y[0] and x[1] == y[1] and (x[2] == y[2]).all() for x, y in zip(solo, results[0]))
        if not ok:
            bad = [i for i, (x, y) in enumerate(zip(solo, results[0])) if not (x[2] == y[2]).all()]
            worst = max(float(abs(x[2] - y[2]).max()) for x, y in zip(solo, results[0]))
            print(f"check: {len(bad)} of {len(solo)} windows differ (first {bad[:8]}), max |diff| {worst:.3e}", file=sys.stderr)
    if rank == 0:
        print(json.dumps({"metric": "sliding-window windows/s (synthetic 3-view videos, host frame synthesis included)",
                          "value": n_win / float(dt.item()), "unit": "windows/s", "n_gpus": world, "windows": n_win,
                          "seconds": float(dt.item()), "sharded_equals_single_rank": ok}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
