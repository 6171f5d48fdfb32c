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
