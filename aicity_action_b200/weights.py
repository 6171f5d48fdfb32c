"""Per-dtype cache of Linear weights (fp32 master parameters -> bf16 operands for tcgen05 GEMMs).

Staleness contract.  A cached operand is keyed on (dtype, `param._version`, device, `data_ptr`): optimizer steps,
`load_state_dict` and every other in-place update through the parameter bump `_version` and refresh the operand on next use.
The refresh is done IN PLACE (`copy_` into the same storage) whenever dtype / shape / device still match, so a captured
CUDA graph (`graphed.GraphedForward`), which reads the operand by address, keeps reading valid, current memory.  Updates
that bypass version counting (`p.data.copy_()`, `p.data.mul_()` — the momentum / EMA encoder idiom) are invisible to the
key: call `invalidate_caches(model)` after them (it marks every cached operand stale; the next forward refreshes in place).
"""
from __future__ import annotations

import torch

_ATTR = "_b200_wcache"
_STALE = "_b200_wstale"
_SHADOW = "_b200_shadow"
_GENERATION = 0


def generation() -> int:
    """Bumped by optimizers that update master weights without touching autograd's version counter (optim.FusedAdamW):
    part of every cache key, so derived operands (transposed / folded weights, pos-embed tables) rebuild after a step."""
    return _GENERATION


def bump_generation() -> None:
    global _GENERATION
    _GENERATION += 1


def register_shadow(param: torch.Tensor, view: torch.Tensor) -> None:
    """`view` (bf16, same shape) is maintained by the fused optimizer kernel as the tensor-core operand of `param`:
    `cached_weight(param, bf16)` returns it without any per-step cast."""
    setattr(param, _SHADOW, [view, param._version])


def mark_shadow_current(param: torch.Tensor) -> None:
    slot = getattr(param, _SHADOW, None)
    if slot is not None:
        slot[1] = param._version


def _refresh(param, attr, key, source):
    """`source()` -> an (un-materialised) view of the master weights in the operand's layout."""
    slot = getattr(param, attr, None)
    stale = getattr(param, _STALE, 0)
    if slot is not None and slot[0] == key and not (stale and slot[2] != stale):
        return slot[1]
    dtype = key[0]
    src = source()
    if slot is not None and slot[0][0] == dtype and slot[0][2] == key[2] and slot[1].shape == src.shape:
        fresh = slot[1]
        fresh.copy_(src)                         # cast (+ transpose) in one kernel, same storage: captured graphs that read
    else:                                        # the operand by address stay valid
        fresh = src.to(dtype).contiguous()
    try:
        setattr(param, attr, (key, fresh, stale))
    except AttributeError:                       # plain tensors without __dict__ are converted every call
        pass
    return fresh


def invalidate_caches(module: torch.nn.Module) -> int:
    """Mark every cached low-precision operand under `module` stale (needed only after updates that bypass autograd's
    version counter, e.g. `p.data.copy_()`).  Returns the number of parameters touched."""
    n = 0
    for p in module.parameters():
        if hasattr(p, _ATTR) or hasattr(p, _ATTR + "_t") or hasattr(p, _ATTR + "_ln"):
            setattr(p, _STALE, getattr(p, _STALE, 0) + 1)
            n += 1
    for m in module.modules():                   # module-level caches (patch-embed GEMM weights, pos-embed table)
        for a in ("_b200_w", "_b200_wf", "_b200_pos"):
            if getattr(m, a, None) is not None:
                setattr(m, a, None)
                n += 1
    return n


def cached_weight(param: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """`param` converted to `dtype`, contiguous.  The converted copy is stored on the parameter object
    and rebuilt when the parameter is updated in place (optimizer step / load_state_dict bump
    `_version`), re-allocated, or moved to another device."""
    if param.dtype == dtype and param.is_contiguous():
        return param.detach()
    shadow = getattr(param, _SHADOW, None)
    if shadow is not None and shadow[0].dtype == dtype and shadow[0].device == param.device:
        if shadow[1] != param._version:          # updated through torch (load_state_dict, manual in-place op): re-sync
            shadow[0].copy_(param.detach())
            shadow[1] = param._version
        return shadow[0]
    key = (dtype, param._version, param.device, param.data_ptr(), _GENERATION)
    return _refresh(param, _ATTR, key, lambda: param.detach())


def cached_weight_t(param: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """`param` [N, K] as a contiguous [K, N] `dtype` tensor — the operand of the input-gradient GEMM
    dx = dy · W (a Linear whose weight is Wᵀ).  Cached like `cached_weight`."""
    key = (dtype, param._version, param.device, param.data_ptr(), _GENERATION)
    return _refresh(param, _ATTR + "_t", key, lambda: param.detach().t())


def folded_ln_linear(weight: torch.Tensor, bias, gamma: torch.Tensor, beta: torch.Tensor):
    """Operands of the LayerNorm-folded Linear (ops.linear_ln):  LN(x)·Wᵀ + b = rstd·(x·W'ᵀ − mean·colsum) + b'  with
    W' = W·diag(γ) rounded to bf16, colsum[n] = Σ_k W'[n,k] (of the ROUNDED values, so the mean correction cancels exactly
    what the tensor core summed) and b' = b + W·β in fp32.  Cached on `weight`, keyed on the versions of all four
    parameters; refreshed in place so captured graphs keep reading valid memory."""
    ps = [weight, bias, gamma, beta]
    key = tuple((p._version, p.data_ptr()) if p is not None else None for p in ps) + (weight.device, _GENERATION)
    slot = getattr(weight, _ATTR + "_ln", None)
    stale = getattr(weight, _STALE, 0)
    if slot is not None and slot[0] == key and slot[2] == stale:
        return slot[1]
    with torch.no_grad():
        w32 = weight.detach().float()
        wf = (w32 * gamma.detach().float()[None, :]).to(torch.bfloat16).contiguous()
        colsum = wf.float().sum(dim=1).contiguous()
        bf = w32 @ beta.detach().float()
        if bias is not None:
            bf = bf + bias.detach().float()
        bf = bf.contiguous()
    if slot is not None and slot[1][0].shape == wf.shape and slot[1][0].device == wf.device:
        for dst, src in zip(slot[1], (wf, bf, colsum)):
            dst.copy_(src)
        fresh = slot[1]
    else:
        fresh = (wf, bf, colsum)
    setattr(weight, _ATTR + "_ln", (key, fresh, stale))
    return fresh
