#!/usr/bin/env python3
"""Run-to-run and batch-composition determinism of the eval forward (bf16, uint8 frames in), per C-ABI call.

    python tools/determinism_probe.py [--iters 60] [--size 448] [--out gpurun_out/determinism.json]

Legs:
  replay   one captured graph replayed `iters` times on the same frames: every output must equal the first bit for bit
  eager    the same frames through eager launches, every tensor an `ops.*` call returns is checksummed (exact integer
           hash): reports the first call whose hash differs from iteration 0 -> names the kernel
  permute  the batch in reversed clip order: row b of the result must equal row B-1-b of the straight run
  ragged   the first 5 clips alone (B=5) against rows 0..4 of the B=8 run
The sliding-window identity check of bench.py compares exactly such pairs (eager / replay, different batch mates).
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aicity_action_b200 import ops  # noqa: E402
from aicity_action_b200.config import aicity_cfg  # noqa: E402
from aicity_action_b200.graphed import GraphedForward  # noqa: E402
from aicity_action_b200.mvit import MViT  # noqa: E402


def tensor_hash(t: torch.Tensor):
    """Exact (integer) position-sensitive hash of the bits of a tensor."""
    t = t.detach().contiguous()
    es = t.element_size()
    v = t.view({1: torch.uint8, 2: torch.int16, 4: torch.int32, 8: torch.int64}[es]).reshape(-1).to(torch.int64)
    n = v.numel()
    if n == 0:
        return (0, 0)
    w = (torch.arange(n, device=v.device, dtype=torch.int64) % 65521) + 1
    return (int(v.sum()), int((v * w).sum()))


def walk(o, out):
    if isinstance(o, torch.Tensor):
        if o.is_cuda:
            out.append(o)
    elif isinstance(o, (list, tuple)):
        for x in o:
            walk(x, out)


class Tracer:
    """Wraps every public function of `ops` that returns CUDA tensors and records (name, shape, hash) per call."""

    def __init__(self, per_clip=False):
        self.log, self.saved, self.per_clip = [], {}, per_clip

    def __enter__(self):
        for name in dir(ops):
            fn = getattr(ops, name)
            if name.startswith("_") or not callable(fn) or isinstance(fn, type) or getattr(fn, "__module__", "") != ops.__name__:
                continue
            self.saved[name] = fn
            setattr(ops, name, self._wrap(name, fn))
        return self

    def _wrap(self, name, fn):
        def inner(*a, **k):
            r = fn(*a, **k)
            ts = []
            walk(r, ts)
            for n, t in enumerate(ts):
                if self.per_clip and t.dim() >= 2:
                    h = tuple(tensor_hash(t[b]) for b in range(t.shape[0]))
                else:
                    h = tensor_hash(t)
                self.log.append((f"{name}[{n}]", tuple(t.shape), h))
            return r
        return inner

    def __exit__(self, *exc):
        for name, fn in self.saved.items():
            setattr(ops, name, fn)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=60)
    ap.add_argument("--size", type=int, default=448)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--out", default="gpurun_out/determinism.json")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    cfg = aicity_cfg("MVITV2_FULL_B_16x4_CONV_448" if args.size == 448 else "MVITV2_FULL_B_16x4_CONV")
    torch.manual_seed(0)
    model = MViT(cfg).eval().to(dev)
    B, T, S = args.batch, cfg.DATA.NUM_FRAMES, cfg.DATA.TRAIN_CROP_SIZE
    g = torch.Generator().manual_seed(7)
    frames = torch.randint(0, 256, (B, T, S, S, 3), dtype=torch.uint8, generator=g).to(dev)
    res = {"size": S, "batch": B, "iters": args.iters}

    with torch.no_grad():
        # ---- replay ---------------------------------------------------------------------------
        buf = frames.clone()
        gf = GraphedForward(model, buf)
        first = gf().clone()
        bad = 0
        worst = 0.0
        for _ in range(args.iters):
            o = gf()
            if not torch.equal(o, first):
                bad += 1
                worst = max(worst, float((o.float() - first.float()).abs().max()))
        res["replay"] = {"mismatching_replays": bad, "max_abs_diff": worst}
        print("replay", res["replay"], flush=True)

        # ---- eager, untraced ----------------------------------------------------------------------
        e_first = model([frames]).clone()
        bad = sum(0 if torch.equal(model([frames]), e_first) else 1 for _ in range(args.iters))
        res["eager"] = {"mismatching_runs": bad, "equals_replay": bool(torch.equal(e_first, first))}
        print("eager", res["eager"], flush=True)

        # ---- eager, traced ------------------------------------------------------------------------
        ref_log, first_bad = None, {}
        for it in range(max(4, args.iters // 4)):
            with Tracer() as tr:
                model([frames])
            if ref_log is None:
                ref_log = tr.log
                continue
            for n, (a, b) in enumerate(zip(ref_log, tr.log)):
                if a != b:
                    key = f"call {n}: {a[0]} {list(a[1])}"
                    first_bad[key] = first_bad.get(key, 0) + 1
                    break
        res["eager_traced"] = {"runs": max(4, args.iters // 4), "calls_per_forward": len(ref_log),
                               "first_divergent_call": first_bad}
        print("eager_traced", res["eager_traced"], flush=True)

        # ---- permute / ragged, traced per clip ----------------------------------------------------
        def traced(x):
            with Tracer(per_clip=True) as tr:
                y = model([x]).clone()
            return y, tr.log

        y0, log0 = traced(frames)
        yp, logp = traced(frames.flip(0).contiguous())
        perm_bad = None
        for n, (a, b) in enumerate(zip(log0, logp)):
            if isinstance(a[2], tuple) and len(a[2]) == B and isinstance(a[2][0], tuple) and a[1][0] == B:
                if tuple(reversed(b[2])) != a[2]:
                    perm_bad = f"call {n}: {a[0]} {list(a[1])} clips " + str(
                        [i for i in range(B) if a[2][i] != b[2][B - 1 - i]])
                    break
        res["permute"] = {"output_equal": bool(torch.equal(y0, yp.flip(0))), "first_divergent_call": perm_bad}
        print("permute", res["permute"], flush=True)

        nr = min(5, B)
        yr, logr = traced(frames[:nr].contiguous())
        rag_bad = None
        for n, (a, b) in enumerate(zip(log0, logr)):
            if isinstance(a[2], tuple) and len(a[2]) == B and isinstance(a[2][0], tuple) and a[1][0] == B and b[1][0] == nr:
                if tuple(b[2]) != a[2][:nr]:
                    rag_bad = f"call {n}: {a[0]} {list(a[1])} clips " + str([i for i in range(nr) if a[2][i] != b[2][i]])
                    break
        res["ragged"] = {"output_equal": bool(torch.equal(y0[:nr], yr)), "first_divergent_call": rag_bad,
                         "calls_match": len(log0) == len(logr)}
        print("ragged", res["ragged"], flush=True)

    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
