"""Deterministic synthetic weights / clips shared by the golden generator and the tests.

Weights are a pure function of (seed, parameter name, shape) so that the reference
model (in `oracle/make_golden.py`), the CPU oracle and the CUDA path can all be fed the
*same* state_dict without shipping tens of MB of tensors: only names+shapes are needed.

Unlike the reference's init (zero biases, LayerNorm (1,0)) every affine term is made
non-trivial so that bias / gamma / beta paths are actually exercised by parity checks.
"""
from __future__ import annotations

import zlib
from typing import Dict, Sequence

import torch


def _gen(seed: int, name: str) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((seed * 1000003 + zlib.crc32(name.encode())) % (2 ** 31 - 1))
    return g


def synth_tensor(seed: int, name: str, shape: Sequence[int]) -> torch.Tensor:
    g = _gen(seed, name)
    shape = tuple(shape)
    leaf = name.rsplit(".", 1)[-1]
    owner = name.rsplit(".", 2)[-2] if name.count(".") >= 1 else ""
    is_norm = owner.startswith("norm") or owner == "norm"
    if is_norm and leaf == "weight":
        return 1.0 + 0.1 * torch.randn(shape, generator=g)
    if is_norm and leaf == "bias":
        return 0.1 * torch.randn(shape, generator=g)
    if leaf == "bias":
        return 0.05 * torch.randn(shape, generator=g)
    if owner.startswith("pool_"):            # depthwise conv [d,1,kt,kh,kw]
        return 0.25 * torch.randn(shape, generator=g)
    if name.startswith("pos_embed") or name in ("cls_token",):
        return 0.1 * torch.randn(shape, generator=g)
    if name == "patch_embed.proj.weight":
        fan_in = shape[1] * shape[2] * shape[3] * shape[4]
        return torch.randn(shape, generator=g) / fan_in ** 0.5
    if name == "head.projection.weight":
        return 0.2 * torch.randn(shape, generator=g)
    if len(shape) == 2:                      # Linear [out,in]: keep activations O(1)
        return torch.randn(shape, generator=g) / shape[1] ** 0.5
    return 0.02 * torch.randn(shape, generator=g)


def synth_state_dict(shapes: Dict[str, Sequence[int]], seed: int = 0) -> Dict[str, torch.Tensor]:
    return {k: synth_tensor(seed, k, s).contiguous() for k, s in shapes.items()}


def synth_clip(seed: int, batch: int, frames: int, size: int) -> torch.Tensor:
    g = _gen(seed, f"clip{batch}x{frames}x{size}")
    return torch.randn((batch, 3, frames, size, size), generator=g)


def synth_input(seed: int, name: str, shape: Sequence[int]) -> torch.Tensor:
    return torch.randn(tuple(shape), generator=_gen(seed, "in:" + name))
