"""Per-dtype cache of Linear weights (fp32 master parameters -> bf16 operands for tcgen05 GEMMs)."""
from __future__ import annotations

import torch

_ATTR = "_b200_wcache"


def cached_weight(param: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """`param` converted to `dtype`, contiguous.  The converted copy is stored on the parameter object
    and rebuilt when the parameter is updated in place (optimizer step / load_state_dict bump
    `_version`), re-allocated, or moved to another device."""
    if param.dtype == dtype and param.is_contiguous():
        return param.detach()
    key = (dtype, param._version, param.device, param.data_ptr())
    slot = getattr(param, _ATTR, None)
    if slot is None or slot[0] != key:
        slot = (key, param.detach().to(dtype).contiguous())
        try:
            setattr(param, _ATTR, slot)
        except AttributeError:      # plain tensors without __dict__ are converted every call
            pass
    return slot[1]


def cached_weight_t(param: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """`param` [N, K] as a contiguous [K, N] `dtype` tensor — the operand of the input-gradient GEMM
    dx = dy · W (a Linear whose weight is Wᵀ).  Cached like `cached_weight`."""
    key = (dtype, param._version, param.device, param.data_ptr())
    slot = getattr(param, _ATTR + "_t", None)
    if slot is None or slot[0] != key:
        slot = (key, param.detach().to(dtype).t().contiguous())
        try:
            setattr(param, _ATTR + "_t", slot)
        except AttributeError:
            pass
    return slot[1]
