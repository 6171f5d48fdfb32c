#!/usr/bin/env python3
"""Benchmark of the MViTv2 multiscale-attention path (BASELINE.json metric: MViTv2-B 16x4@448 clips/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch 8]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

One "step" = one forward of the drop-in MViT (configs/Aicity/MVITV2_FULL_B_16x4_CONV_448, random
init, eval, bf16) over a batch of 8 synthetic clips per GPU (BASELINE config 2).  Rank 0 prints ONE
JSON line:
  value      clips/s, inputs resident in HBM when the timed region starts (CUDA events, max over ranks)
  e2e        clips/s through the public API from pinned HOST uint8 frames (H2D + on-device normalise +
             forward + D2H of the probabilities inside the timed region, copies double-buffered)
  roofline   dominant kernel (fused tcgen05 attention): algorithmic FLOP / CUDA-event time vs the
             measured bf16 peak of MEASURED_PEAKS.json;  `kernels` carries the other kernel families
             (attention_pool in GB/s vs the measured HBM peak, GEMMs in TFLOP/s)
  cpu_baseline  the CPU oracle port (oracle/mvit_oracle.py, fp32 torch on the host cores) on a bounded
             sample of the same workload
`--impl reference` times that CPU port alone (the reference is pure PyTorch-on-CPU for this path and its
source tree cannot travel to the GPU box; the port is pinned to it bit-for-bit by tests/golden).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIG_NAME = "MVITV2_FULL_B_16x4_CONV_448"
# algorithmic FLOP per clip @448, 2 FLOP per MAC
FLOP_PER_CLIP = 856.4e9
METRIC = "MViTv2-B 16x4@448 inference clips/s"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.t.join(timeout=2)
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v == "Active"})
        pw = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "power_w_max": max(pw) if pw else None, "samples": len(self.rows)}


def cpu_port_clips_per_s(steps, warmup, size=448, batch=1):
    """The CPU oracle port on all host cores: fp32, batch 1, same cfg and init scheme as the GPU arm."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import mvit_oracle as O
    from aicity_action_b200.config import aicity_cfg
    from aicity_action_b200.mvit import MViT

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = aicity_cfg(CONFIG_NAME)
    torch.manual_seed(0)
    sd = {k: v.detach() for k, v in MViT(cfg).state_dict().items()}   # parameter container only (no compute)
    spec = O.derive_spec(cfg)
    torch.manual_seed(1)
    x = torch.randn(batch, 3, cfg.DATA.NUM_FRAMES, size, size)
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            O.mvit_forward(x, sd, spec)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    total = sum(times)
    return batch * len(times) / total, cores, 1e3 * total / len(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    steps, warmup = max(1, min(args.steps, 10)), max(1, min(args.warmup, 2))
    v, cores, ms = cpu_port_clips_per_s(steps, warmup)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "clips/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{CONFIG_NAME} eval forward, batch 1 per step, fp32, host cores"},
            "cpu_baseline": {"value": v, "unit": "clips/s", "cores": cores, "kind": "port",
                             "sample": f"{steps} timed forwards of 1 clip @448 after {warmup} warm-up"},
            "e2e": {"value": v, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)
    return 0


def summarize_events(log, peaks, steps, base):
    """Per kernel family: launches, device time (union of the [start, end] intervals, so launches that overlap on
    side streams are not double counted) and achieved throughput of the algorithmic work."""
    out = {}
    for cat, evs in log.items():
        spans = sorted((base.elapsed_time(a), base.elapsed_time(b)) for a, b, _ in evs)
        ms, cur_a, cur_b = 0.0, None, None
        for a, b in spans:
            if cur_b is None or a > cur_b:
                if cur_b is not None:
                    ms += cur_b - cur_a
                cur_a, cur_b = a, b
            else:
                cur_b = max(cur_b, b)
        if cur_b is not None:
            ms += cur_b - cur_a
        work = sum(w for _, _, w in evs)
        ent = {"launches_per_step": len(evs) / steps, "ms_per_step": ms / steps}
        if cat.startswith("pool") or cat in ("layernorm", "fold_clip", "im2col"):
            ent.update(bound="hbm", achieved=work / (ms * 1e-3) / 1e9, peak=peaks["hbm_gbs"], unit="GB/s")
        else:
            ent.update(bound="tensor", achieved=work / (ms * 1e-3) / 1e12, peak=peaks["bf16_tflops_sustained"],
                       unit="TFLOP/s")
        ent["frac"] = ent["achieved"] / ent["peak"]
        out[cat] = ent
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    from aicity_action_b200 import ops
    from aicity_action_b200.config import aicity_cfg
    from aicity_action_b200.mvit import MViT

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    B, K, W = args.batch, args.steps, max(3, args.warmup)
    cfg = aicity_cfg(CONFIG_NAME)
    size, frames = cfg.DATA.TRAIN_CROP_SIZE, cfg.DATA.NUM_FRAMES
    torch.manual_seed(0)
    model = MViT(cfg).eval().to(dev)
    peaks = load_peaks()

    # synthetic uint8 frames [B, T, H, W, 3] (what the sliding-window reader produces after resize), pinned
    g = torch.Generator().manual_seed(1 + rank)
    n_host = 2
    host = [torch.randint(0, 256, (B, frames, size, size, 3), dtype=torch.uint8, generator=g).pin_memory()
            for _ in range(n_host)]
    dev_clip = ops.preprocess_u8(host[0].to(dev), torch.bfloat16)     # resident input, 154 MB > 126 MB L2
    probs_host = torch.empty((B, cfg.MODEL.NUM_CLASSES), dtype=torch.float32).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    from aicity_action_b200.graphed import GraphedForward
    use_graph = not args.no_cuda_graph
    with torch.no_grad():
        for _ in range(W):
            out = model([dev_clip])
        # the public GraphedForward wrapper: the forward (171 launches) is captured once per input buffer and replayed
        # with one host call, so the step rate does not depend on how many ranks share the host's cores
        fwd = GraphedForward(model, dev_clip) if use_graph else (lambda: model([dev_clip]))
        for _ in range(2):
            out = fwd()
        barrier()
        # ---- timed region 1: device-resident inputs ----------------------------------------
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        n0 = ops.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(K):
            out = fwd()
        e1.record()
        barrier()
        launches = (ops.launch_count - n0)
        ms = e0.elapsed_time(e1)
        # ---- same K steps again with every C-ABI launch bracketed by CUDA events on its stream: the per-kernel
        # durations behind `roofline` / `kernels` (kept out of region 1 so event records do not perturb `value`)
        ops.event_log = {}
        i0, i1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        i0.record()
        for _ in range(K):
            out = model([dev_clip])
        i1.record()
        barrier()
        log, ops.event_log = ops.event_log, None
        ms_instr = i0.elapsed_time(i1)
        clocks = sampler.stop() if rank == 0 else None
        assert torch.isfinite(out).all()

        # ---- timed region 2: end to end from pinned host memory ----------------------------
        copy_stream = torch.cuda.Stream()
        dbuf = [torch.empty_like(host[0], device=dev) for _ in range(2)]
        for b_ in dbuf:
            b_.copy_(host[0])
        if use_graph:
            g0 = GraphedForward(model, dbuf[0])
            fwd_u8 = [g0, GraphedForward(model, dbuf[1], pool=g0.pool)]
        else:
            fwd_u8 = [lambda b_=b_: model([b_]) for b_ in dbuf]
        ready = [torch.cuda.Event() for _ in range(2)]
        freed = [torch.cuda.Event() for _ in range(2)]

        def upload(i):
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(freed[i % 2])
                dbuf[i % 2].copy_(host[i % n_host], non_blocking=True)
                ready[i % 2].record(copy_stream)

        def e2e_loop(n):
            cur = torch.cuda.current_stream()
            for i in range(2):
                freed[i].record(cur)
            upload(0)
            for i in range(n):
                if i + 1 < n:
                    upload(i + 1)
                cur.wait_event(ready[i % 2])
                probs = fwd_u8[i % 2]()              # uint8 frames: normalise + fold + patch-embed on the device
                freed[i % 2].record(cur)
                probs_host.copy_(probs, non_blocking=True)
            cur.synchronize()

        e2e_loop(2)
        barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        e2e_loop(K)
        t1.record()
        barrier()
        ms_e2e = t0.elapsed_time(t1)

    # ---- secondary figure: one training step (forward + backward + AdamW) of the same model on the same clips, bf16
    # autocast as TRAIN.MIXED_PRECISION does, DistributedDataParallel over NCCL when world > 1 (build.py:44-53)
    ms_train, train_launches, train_steps = 0.0, 0, 0
    if not args.no_train:
        import torch.nn.functional as F
        # the reference's recipe (README.md:100-112): activation checkpointing, DropPath 0.4, head
        # dropout 0.5, AdamW(1e-4, wd 1e-4, eps 1e-8), gradient clipping at 1.0
        tcfg = aicity_cfg(CONFIG_NAME, ["MODEL.ACT_CHECKPOINT", True, "MVIT.DROPPATH_RATE", 0.4, "MODEL.DROPOUT_RATE", 0.5])
        tmodel = MViT(tcfg).to(dev)
        tmodel.load_state_dict(model.state_dict())
        model = tmodel.train()
        net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local]) if world > 1 else model
        opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=1e-4, eps=1e-8)
        labels = torch.randint(0, cfg.MODEL.NUM_CLASSES, (B,), device=dev)
        train_steps = max(2, min(K, 5))

        def train_step():
            loss = F.cross_entropy(net([dev_clip]).float(), labels)
            opt.zero_grad(set_to_none=True)
            loss.backward()
            torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
            opt.step()
            return loss

        torch.cuda.empty_cache()                     # inference-phase blocks go back before the allocator re-plans
        for _ in range(3):
            train_step()
        barrier()
        n0 = ops.launch_count
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r0.record()
        for _ in range(train_steps):
            loss = train_step()
        r1.record()
        barrier()
        ms_train = r0.elapsed_time(r1)
        train_launches = ops.launch_count - n0
        assert torch.isfinite(loss)
    times = torch.tensor([ms, ms_e2e, ms_train], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms, ms_e2e, ms_train = times.tolist()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    kernels = summarize_events(log, peaks, K, i0)
    attn = kernels.get("attention", {})
    roofline = {"kernel": "attention_tc_kernel (fused tcgen05 pooling attention)", "bound": "tensor",
                "achieved": attn.get("achieved"), "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                "frac": attn.get("frac"), "traffic": None, "peak_source": peaks["source"] + " (sustained bf16)",
                "share_of_step": attn.get("ms_per_step", 0.0) / (ms_instr / K) if ms_instr else None,
                "instrumented_ms_per_step": ms_instr / K}
    # DRAM bytes per attention launch from the committed ncu pass of this same command (profiles/README.md); algorithmic
    # bytes per launch (q, k, v read + out written once) printed beside it
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic_r1.json")) as f:
            tr = json.load(f)
        roofline["traffic"] = tr["avg_dram_bytes_per_launch"]
        roofline["traffic_source"] = tr["source"]
    except (OSError, KeyError, ValueError):
        pass
    value = world * B * K / (ms * 1e-3)
    line = {
        "metric": METRIC, "value": value, "unit": "clips/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": f"{CONFIG_NAME} eval forward, batch {B} per GPU, random init",
                   "l2": "inputs (154 MB bf16 clip batch) and activations exceed the 126 MB L2",
                   "parallelism": f"replicated model, {world} independent clip batches",
                   "launch": "CUDA graph replay (aicity_action_b200.graphed.GraphedForward)" if use_graph else "eager"},
        "model_tflops": value * FLOP_PER_CLIP / 1e12 / world,
        "frac_of_bf16_peak_whole_model": value * FLOP_PER_CLIP / 1e12 / world / peaks["bf16_tflops_sustained"],
        "e2e": {"value": world * B * K / (ms_e2e * 1e-3), "unit": "clips/s",
                "h2d_bytes_per_step": host[0].numel(), "d2h_bytes_per_step": probs_host.numel() * 4,
                "input": "pinned uint8 frames [B,16,448,448,3], normalised on device"},
        "gpu_launches": launches,
        "roofline": roofline, "kernels": kernels, "clocks": clocks,
    }
    if train_steps:
        line["train"] = {"value": world * B * train_steps / (ms_train * 1e-3), "unit": "clips/s",
                         "ms_per_step": ms_train / train_steps, "steps": train_steps, "batch_per_gpu": B,
                         "gpu_launches": train_launches,
                         "what": "forward + backward (MODEL.ACT_CHECKPOINT True under the 'auto' policy: activations fit in HBM so "
                                 "nothing is recomputed; DropPath 0.4, head dropout 0.5) + "
                                 "grad-clip + AdamW, bf16 activations / fp32 master weights"
                                 + (", DDP gradient all-reduce over NCCL" if world > 1 else "")}
    if world == 1 and not args.no_cpu_baseline:
        v, cores, cpu_ms = cpu_port_clips_per_s(steps=2, warmup=1)
        line["cpu_baseline"] = {"value": v, "unit": "clips/s", "cores": cores, "kind": "port",
                                "sample": "2 timed fp32 forwards of 1 clip @448 after 1 warm-up (oracle port)"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the secondary training-step measurement")
    ap.add_argument("--no-cuda-graph", action="store_true", help="issue every launch from Python instead of graph replay")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
