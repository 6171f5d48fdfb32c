"""Mlp / DropPath with the parameter layout of slowfast/models/common.py, computed by the B200 kernels."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import autograd as AG
from . import ops
from .weights import cached_weight, folded_ln_linear


def drop_path_scale(batch: int, drop_prob: float, training: bool, device, dtype=torch.float32):
    """Per-sample multiplier of stochastic depth (common.py:46-59): floor(keep + U[0,1)) / keep.

    Returns None when the path is kept deterministically (eval or p == 0).  The random draw uses
    torch's generator like the reference (`torch.rand(shape, dtype, device)`).  MultiScaleBlock draws both masks of a
    block up front in fp32 (attention branch, then MLP branch): with MVIT.DROPOUT_RATE == 0 (every shipped config) and
    fp32 activations that is exactly the reference's consumption of the RNG stream; with dropout layers active or bf16
    activations the reference interleaves its draws differently (dropout masks sit between the two DropPath draws, and
    it draws in the activation dtype), so seeded runs are then statistically, not bitwise, equivalent.  The multiply
    itself is fused into the GEMM epilogue."""
    if drop_prob == 0.0 or not training:
        return None
    keep = 1.0 - drop_prob
    mask = keep + torch.rand((batch, 1, 1), dtype=dtype, device=device)
    mask.floor_()
    return (mask / keep).reshape(batch).float()


class DropPath(nn.Module):
    """State-less marker module; the block reads `drop_prob` and fuses the scaling (common.py:62-89)."""

    def __init__(self, drop_prob=None):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        s = drop_path_scale(x.shape[0], self.drop_prob or 0.0, self.training, x.device, x.dtype)
        if s is None:
            return x
        return x * s.to(x.dtype).reshape((x.shape[0],) + (1,) * (x.ndim - 1))


class Mlp(nn.Module):
    """fc1 -> GELU(erf) -> fc2 (common.py:7-34).  `fc1`/`fc2` are nn.Linear parameter holders so the
    state_dict keys (`mlp.fc1.weight` ...) match; the math is two fused-epilogue GEMMs."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop_rate=0.0):
        super().__init__()
        self.drop_rate = drop_rate
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features)
        if self.drop_rate > 0.0:
            self.drop = nn.Dropout(drop_rate)
        if not isinstance(self.act, nn.GELU) or getattr(self.act, "approximate", "none") != "none":
            raise NotImplementedError("the B200 MLP kernel fuses exact (erf) GELU only")

    def forward(self, x, residual=None, row_scale=None, ln_in=None, want_stats=False):
        """Returns fc2(gelu(fc1(x))) * row_scale + residual.
        ln_in = (norm2, row statistics of x): `x` is the raw block stream and norm2 is folded into fc1; want_stats: the
        fc2 epilogue emits the row statistics of the result (eval / bf16, see MultiScaleBlock.forward)."""
        if ln_in is not None or want_stats:
            if ln_in is not None:
                norm, stats = ln_in
                wf, bf, colsum = folded_ln_linear(self.fc1.weight, self.fc1.bias, norm.weight, norm.bias)
                if (residual is x and row_scale is None and ops.mlp_fused_enabled()
                        and ops.mlp_fused_supported(self.fc1.in_features, self.fc1.out_features, self.fc2.out_features)):
                    # front stages: fc1 -> GELU -> fc2 -> +x in one kernel, the 4C-wide hidden tile never reaches HBM
                    return ops.mlp_fused(x, stats, wf, bf, colsum, cached_weight(self.fc2.weight, x.dtype), self.fc2.bias,
                                         norm.eps)
                h = ops.linear_ln(x, stats, wf, bf, colsum, norm.eps, gelu=True)
            else:
                h = AG.linear(x, self.fc1.weight, self.fc1.bias, gelu=True)
            return ops.linear_stats(h, cached_weight(self.fc2.weight, h.dtype), self.fc2.bias, residual=residual,
                                    row_scale=row_scale)
        if self.drop_rate > 0.0 and self.training:
            # dropout after the activation and after fc2 (common.py:28-33): un-fused elementwise tail, see attention.py
            h = self.drop(AG.linear(x, self.fc1.weight, self.fc1.bias, gelu=True))
            y = self.drop(AG.linear(h, self.fc2.weight, self.fc2.bias))
            return AG.scale_add(y, residual, row_scale)
        h = AG.linear(x, self.fc1.weight, self.fc1.bias, gelu=True)
        return AG.linear(h, self.fc2.weight, self.fc2.bias, residual=residual, row_scale=row_scale)
