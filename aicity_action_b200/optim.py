"""Training-step glue for the B200 path: flat parameter / gradient arenas, a fused AdamW + global-norm
clip (two kernel launches per step), a bf16 weight shadow maintained by the optimizer kernel, an arena all-reduce for
data-parallel training and whole-step CUDA-graph capture.

Reference behaviour being replaced (tools/train_net.py:229-246, slowfast/models/optimizer.py:71-75,200-206):
    optimizer.zero_grad(); loss.backward(); clip_grad_norm_(model.parameters(), 1.0); optimizer.step()
with `torch.optim.AdamW(params, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=wd)` over two parameter groups — the
reference gives weight decay to the >= 2-D parameters only (119 tensors) and 0 to the 231 1-D ones; `decay_filter` below
defaults to that rule.  On the stock path this tail costs ~700 tiny launches (zero fills, foreach norm / clip, the
optimizer, 130 fp32 -> bf16 weight casts of the next forward); here it is `arena.zero_()` + mvit_adamw_clip_step.

`FusedAdamW` is a `torch.optim.Optimizer`: `param_groups[i]["lr"]` can be set per iteration exactly as the reference's
`optim.set_lr` does (optimizer.py:250-260), `state_dict()` / `load_state_dict()` round-trip, `zero_grad()` and `step()`
keep their meaning.  Differences a caller can observe: parameters and their `.grad` are views into flat buffers (so
`zero_grad(set_to_none=True)` zeroes instead of dropping them), and gradients are NOT rescaled in place by the clip (the
scale is folded into the update; `last_grad_norm()` returns the pre-clip norm `clip_grad_norm_` would have returned).
"""
from __future__ import annotations

from typing import Callable, Iterable, Optional

import torch

from . import _lib, weights

_ALIGN = 64          # elements: every tensor starts on a 256-byte (fp32) / 128-byte (bf16 shadow) boundary — TMA needs 16


def _default_decay(name: str, p: torch.Tensor) -> bool:
    """optimizer.py:71-75 as it actually behaves: only 1-D parameters skip weight decay."""
    return p.ndim > 1


class ParamArena:
    """Re-homes the parameters of `module` into one flat fp32 buffer laid out [decayed | non-decayed], with matching flat
    gradient buffer (each `p.grad` is a view) and bf16 shadow (each Linear / Conv weight's tensor-core operand is a view).
    Parameter objects, names, shapes and values are unchanged: `state_dict()` is the reference's."""

    def __init__(self, module: torch.nn.Module, decay_filter: Callable[[str, torch.Tensor], bool] = _default_decay):
        named = [(n, p) for n, p in module.named_parameters() if p.requires_grad]
        if not named:
            raise ValueError("ParamArena: module has no trainable parameters")
        dev = named[0][1].device
        if dev.type != "cuda":
            raise _lib.MvitLibraryError("ParamArena needs CUDA parameters (no CPU fallback); call model.cuda() first")
        if any(p.dtype != torch.float32 or p.device != dev for _, p in named):
            raise TypeError("ParamArena: master parameters must be fp32 on one device")
        decayed = [(n, p) for n, p in named if decay_filter(n, p)]
        plain = [(n, p) for n, p in named if not decay_filter(n, p)]
        self.names, self.offsets, self.params = [], [], []
        off = 0
        for group in (decayed, plain):
            for n, p in group:
                self.names.append(n)
                self.offsets.append(off)
                self.params.append(p)
                off += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
            if group is decayed:
                self.n_decay = off
        self.numel = off
        self.flat = torch.zeros(self.numel, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(self.numel, dtype=torch.float32, device=dev)
        self.shadow = torch.zeros(self.numel, dtype=torch.bfloat16, device=dev)
        with torch.no_grad():
            for p, o in zip(self.params, self.offsets):
                view = self.flat[o:o + p.numel()].view_as(p)
                view.copy_(p)
                p.data = view
                p.grad = self.grad[o:o + p.numel()].view_as(p)
                weights.register_shadow(p, self.shadow[o:o + p.numel()].view_as(p))
        self.sync_shadow()

    def sync_shadow(self):
        """bf16 shadow <- master weights (after load_state_dict or any update the fused optimizer did not make)."""
        self.shadow.copy_(self.flat)
        for p in self.params:
            weights.mark_shadow_current(p)

    def zero_grad(self):
        self.grad.zero_()
        for p, o in zip(self.params, self.offsets):       # a caller may have dropped .grad (set_to_none): re-attach
            if p.grad is None or p.grad.data_ptr() != self.grad.data_ptr() + 4 * o:
                p.grad = self.grad[o:o + p.numel()].view_as(p)


class FusedAdamW(torch.optim.Optimizer):
    """AdamW + optional global-norm clip over a ParamArena: `step()` = mvit_adamw_clip_step (two launches)."""

    def __init__(self, module_or_arena, lr: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-4,
                 max_grad_norm: Optional[float] = None, decay_filter: Callable[[str, torch.Tensor], bool] = _default_decay):
        arena = module_or_arena if isinstance(module_or_arena, ParamArena) else ParamArena(module_or_arena, decay_filter)
        self.arena = arena
        n_dec = sum(1 for o in arena.offsets if o < arena.n_decay)
        groups = [{"params": arena.params[:n_dec], "weight_decay": weight_decay},
                  {"params": arena.params[n_dec:], "weight_decay": 0.0}]
        groups = [g for g in groups if g["params"]]
        super().__init__(groups, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay))
        self.max_grad_norm = max_grad_norm
        dev = arena.flat.device
        lib = _lib.load()
        self.exp_avg = torch.zeros_like(arena.flat)
        self.exp_avg_sq = torch.zeros_like(arena.flat)
        self._hyper = torch.zeros(int(lib.mvit_adamw_hyper_floats()), dtype=torch.float32, device=dev)
        self._hyper_host = torch.zeros(6, dtype=torch.float32).pin_memory()
        self._ws = torch.zeros(int(lib.mvit_adamw_workspace_floats()), dtype=torch.float32, device=dev)
        self._last = None

    # -- hyper-parameters live on the device so that a captured graph replays with new values
    def _push_hyper(self):
        g0 = self.param_groups[0]
        lrs = {g["lr"] for g in self.param_groups}
        if len(lrs) != 1:
            raise ValueError("FusedAdamW: all parameter groups must share one learning rate")
        wd = g0["weight_decay"] if self.arena.n_decay > 0 else 0.0
        vals = (g0["lr"], g0["betas"][0], g0["betas"][1], g0["eps"], wd, self.max_grad_norm or 0.0)
        if vals != self._last:
            self._hyper_host.copy_(torch.tensor(vals, dtype=torch.float32))
            self._hyper[:6].copy_(self._hyper_host, non_blocking=True)       # the int step counter at [6] is device-owned
            self._last = vals

    def zero_grad(self, set_to_none: bool = True):
        self.arena.zero_grad()

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        if not torch.cuda.is_current_stream_capturing():
            self._push_hyper()
        a = self.arena
        _lib.check(_lib.load().mvit_adamw_clip_step(a.flat.data_ptr(), a.grad.data_ptr(), self.exp_avg.data_ptr(),
                                                    self.exp_avg_sq.data_ptr(), a.shadow.data_ptr(), a.n_decay, a.numel,
                                                    self._hyper.data_ptr(), self._ws.data_ptr(),
                                                    torch.cuda.current_stream().cuda_stream), "mvit_adamw_clip_step")
        from . import ops
        ops.launch_count += 2
        weights.bump_generation()              # transposed / folded operand caches rebuild from the new master weights
        return loss

    def set_lr(self, lr: float):
        """Per-iteration learning rate (the reference's optim.set_lr); safe between replays of a captured step."""
        for g in self.param_groups:
            g["lr"] = lr
        self._push_hyper()

    def last_grad_norm(self) -> float:
        """Pre-clip global gradient norm of the last step (synchronises)."""
        return float(self._hyper[7].item())

    def steps_done(self) -> int:
        return int(self._hyper[6:7].view(torch.int32).item())

    def state_dict(self):
        sd = super().state_dict()
        sd["fused"] = {"exp_avg": self.exp_avg.clone(), "exp_avg_sq": self.exp_avg_sq.clone(), "step": self.steps_done(),
                       "names": list(self.arena.names), "offsets": list(self.arena.offsets)}
        return sd

    def load_state_dict(self, state_dict):
        fused = state_dict.get("fused")
        super().load_state_dict({k: v for k, v in state_dict.items() if k != "fused"})
        if fused is not None:
            self.exp_avg.copy_(fused["exp_avg"])
            self.exp_avg_sq.copy_(fused["exp_avg_sq"])
            self._hyper[6:7].view(torch.int32).fill_(int(fused["step"]))
        self._last = None


class ArenaDataParallel(torch.nn.Module):
    """Data-parallel wrapper for arena-backed models: parameters are broadcast from rank 0 once, and after backward ONE
    all-reduce (NCCL over NVLink; gloo in CPU tests) averages the whole flat gradient buffer — 141 MB fp32 for MViTv2-B,
    ~0.35 ms at the measured 725 GB/s bus bandwidth — instead of DistributedDataParallel's bucketed reduction whose last
    buckets were exposed after block 0's backward.  `bf16_comm=True` halves the bytes (sum in bf16, as
    torch's bf16 compress hook).  Call `reduce_gradients()` between `backward()` and `optimizer.step()`;
    `module` attribute / `state_dict()` keep DistributedDataParallel's conventions."""

    def __init__(self, module: torch.nn.Module, arena: ParamArena, process_group=None, bf16_comm: bool = False):
        super().__init__()
        import torch.distributed as dist
        self.module, self.arena, self.group, self.bf16_comm = module, arena, process_group, bf16_comm
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        if self.world > 1:
            dist.broadcast(arena.flat, src=dist.get_global_rank(process_group, 0) if process_group is not None else 0,
                           group=process_group)
            arena.sync_shadow()
        self._comm = torch.empty_like(arena.grad, dtype=torch.bfloat16) if bf16_comm else None

    def forward(self, *a, **k):
        return self.module(*a, **k)

    def reduce_gradients(self):
        if self.world == 1:
            return
        import torch.distributed as dist
        if self.bf16_comm:
            self._comm.copy_(self.arena.grad)
            dist.all_reduce(self._comm, group=self.group)
            self.arena.grad.copy_(self._comm)
            self.arena.grad.mul_(1.0 / self.world)
        else:
            self.arena.grad.mul_(1.0 / self.world)
            dist.all_reduce(self.arena.grad, group=self.group)


class GraphedTrainStep:
    """One optimisation step — zero_grad, forward, loss, backward, [gradient all-reduce], fused AdamW + clip — captured as
    a single CUDA graph over static input buffers and replayed with one host call.  Requires MODEL.ACT_CHECKPOINT off or
    MVIT_B200_ACT_CHECKPOINT=never (torch.utils.checkpoint stashes RNG state through a host call that cannot be captured)."""

    def __init__(self, model: torch.nn.Module, optimizer: FusedAdamW, loss_fn, x: torch.Tensor, y: torch.Tensor,
                 net: Optional[torch.nn.Module] = None, warmup: int = 3):
        self.x, self.y, self.opt = x, y, optimizer
        net = net if net is not None else model
        reduce = getattr(net, "reduce_gradients", None)

        def step():
            optimizer.zero_grad()
            loss = loss_fn(net([self.x]), self.y)
            loss.backward()
            if reduce is not None:
                reduce()
            optimizer.step()
            return loss

        side = torch.cuda.Stream(device=x.device)
        side.wait_stream(torch.cuda.current_stream(x.device))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                step()
        torch.cuda.current_stream(x.device).wait_stream(side)
        optimizer._push_hyper()
        self.graph = torch.cuda.CUDAGraph()
        from . import ops
        n0 = ops.launch_count
        with torch.cuda.graph(self.graph):
            self.loss = step()
        self.launches = ops.launch_count - n0

    def __call__(self, x: Optional[torch.Tensor] = None, y: Optional[torch.Tensor] = None, lr: Optional[float] = None):
        if x is not None and x.data_ptr() != self.x.data_ptr():
            self.x.copy_(x, non_blocking=True)
        if y is not None and y.data_ptr() != self.y.data_ptr():
            self.y.copy_(y, non_blocking=True)
        if lr is not None:
            self.opt.set_lr(lr)
        self.graph.replay()
        from . import ops
        ops.launch_count += self.launches
        return self.loss
