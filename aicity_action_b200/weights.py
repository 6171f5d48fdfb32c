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


def _refresh(param, attr, key, make):
    slot = getattr(param, attr, None)
    stale = getattr(param, _STALE, 0)
    if slot is not None and slot[0] == key and not (stale and slot[2] != stale):
        return slot[1]
    fresh = None
    if slot is not None and slot[0][0] == key[0] and slot[0][2] == key[2]:
        old = slot[1]
        src = make()
        if old.shape == src.shape:
            old.copy_(src)                       # same storage: captured graphs that read it by address stay valid
            fresh = old
        else:
            fresh = src
    if fresh is None:
        fresh = make()
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
        if hasattr(p, _ATTR) or hasattr(p, _ATTR + "_t"):
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
    key = (dtype, param._version, param.device, param.data_ptr())
    return _refresh(param, _ATTR, key, lambda: param.detach().to(dtype).contiguous())


def cached_weight_t(param: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """`param` [N, K] as a contiguous [K, N] `dtype` tensor — the operand of the input-gradient GEMM
    dx = dy · W (a Linear whose weight is Wᵀ).  Cached like `cached_weight`."""
    key = (dtype, param._version, param.device, param.data_ptr())
    return _refresh(param, _ATTR + "_t", key, lambda: param.detach().to(dtype).t().contiguous())
