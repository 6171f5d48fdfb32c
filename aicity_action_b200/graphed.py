"""CUDA-graph replay of the eval forward.

One MViTv2-B forward is ~170 kernel launches (~15 ms of GPU time at batch 8); issued from Python they cost a few ms of
host time per step, which is invisible on an idle host and becomes the limiter when eight ranks share one box's cores.
The C ABI is capture-safe by construction (stream-ordered, no allocation, no synchronisation, TMA descriptors passed as
kernel parameters), so the whole forward — the side-stream K/V pooling fork/join included — is captured once per input
buffer and replayed with a single host call.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops


class GraphedForward:
    """`y = GraphedForward(model, x)()` replays `model([x])` on the contents `x` holds at replay time.

    `x` is the caller's static input buffer (clip [B, 3, T, H, W] float/bf16, or uint8 frames [B, T, H, W, 3]); the result
    is returned in a static output tensor that the next replay overwrites.  Several instances may share one memory pool
    (`pool=other.pool`), e.g. one per upload buffer of a double-buffered pipeline."""

    def __init__(self, model: torch.nn.Module, x: torch.Tensor, *, warmup: int = 2, pool=None):
        if not x.is_cuda:
            raise ops._lib.MvitLibraryError("GraphedForward needs a CUDA input buffer (no CPU fallback)")
        if model.training:
            raise RuntimeError("GraphedForward captures the eval forward; call model.eval() first")
        self.model, self.x = model, x
        side = torch.cuda.Stream(device=x.device)
        side.wait_stream(torch.cuda.current_stream(x.device))
        with torch.no_grad(), torch.cuda.stream(side):
            for _ in range(max(1, warmup)):            # first-use work (cudaFuncSetAttribute, weight casts) stays outside
                model([x])
        torch.cuda.current_stream(x.device).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        n0 = ops.launch_count
        with torch.no_grad(), torch.cuda.graph(self.graph, pool=pool):
            self.out = model([x])
        self.launches = ops.launch_count - n0          # kernels of libmvit_b200.so replayed per call
        self.pool = self.graph.pool()

    def __call__(self, x: Optional[torch.Tensor] = None) -> torch.Tensor:
        if x is not None and x.data_ptr() != self.x.data_ptr():
            self.x.copy_(x, non_blocking=True)
        self.graph.replay()
        ops.launch_count += self.launches
        return self.out
