"""Training-step timing of the B200 path: forward + backward + AdamW on synthetic clips.

    python tools/train_bench.py --batch 8 --size 448 --steps 5 [--fp32] [--no-checkpoint]

Follows the reference's own training recipe (README.md:100-112): ACT_CHECKPOINT True, DROPPATH_RATE 0.4,
head DROPOUT_RATE 0.5, AdamW(lr 1e-4, wd 1e-4, eps 1e-8), clip_grad_norm 1.0, cross-entropy on random labels.

Prints one JSON line with ms/step, clips/s and the per-kernel-category device time (CUDA events on the launching stream).
Under torchrun it wraps the model in DistributedDataParallel (NCCL gradient all-reduce, as build.py:44-53 does)."""
import argparse
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from aicity_action_b200 import ops  # noqa: E402
from aicity_action_b200.config import aicity_cfg  # noqa: E402
from aicity_action_b200.mvit import MViT  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--size", type=int, default=448, choices=[224, 448])
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--fp32", action="store_true")
    ap.add_argument("--no-checkpoint", action="store_true", help="MODEL.ACT_CHECKPOINT False")
    ap.add_argument("--checkpoint-policy", default="auto", choices=["auto", "always", "never"],
                    help="how MODEL.ACT_CHECKPOINT True is honoured (auto: only when activations do not fit)")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        torch.distributed.init_process_group("nccl")
    yaml = "MVITV2_FULL_B_16x4_CONV_448.yaml" if a.size == 448 else "MVITV2_FULL_B_16x4_CONV.yaml"
    cfg = aicity_cfg(yaml, ["MODEL.ACT_CHECKPOINT", not a.no_checkpoint, "MVIT.DROPPATH_RATE", 0.4,
                            "MODEL.DROPOUT_RATE", 0.5])
    torch.manual_seed(0)
    model = MViT(cfg).cuda().train()
    model.act_checkpoint_policy = a.checkpoint_policy
    net = model
    if world > 1:
        net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local])
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=1e-4, eps=1e-8)
    g = torch.Generator(device="cuda").manual_seed(1 + rank)
    x = torch.randn((a.batch, 3, cfg.DATA.NUM_FRAMES, a.size, a.size), device="cuda", generator=g)
    y = torch.randint(0, cfg.MODEL.NUM_CLASSES, (a.batch,), device="cuda", generator=g)

    def step():
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=not a.fp32):
            logits = net([x])
        loss = F.cross_entropy(logits.float(), y)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        opt.step()
        return loss

    for _ in range(a.warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / a.steps], device="cuda")
    if world > 1:
        torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
    # per-category kernel time of one more (instrumented) step
    ops.event_log = {}
    step()
    torch.cuda.synchronize()
    cats = {k: round(sum(s.elapsed_time(e) for s, e, _ in v), 3) for k, v in ops.event_log.items()}
    ops.event_log = None
    if rank == 0:
        print(json.dumps({"metric": "train_clips_per_sec", "value": round(a.batch * world / (ms.item() / 1e3), 2),
                          "ms_per_step": round(ms.item(), 2), "n_gpus": world, "batch_per_gpu": a.batch, "size": a.size,
                          "dtype": "f32" if a.fp32 else "bf16", "act_checkpoint": not a.no_checkpoint, "checkpoint_policy": a.checkpoint_policy,
                          "loss": round(loss.item(), 4), "kernel_ms": cats,
                          "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 2)}))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
