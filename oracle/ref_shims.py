"""Import shims that let the UNMODIFIED reference (`/root/reference`) be imported on CPU.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package imports this file.

The reference (a PySlowFast fork) imports a dozen third-party packages at module
top that are not installed in this image: fvcore,
iopath, simplejson, matplotlib, detectron2, pytorchvideo, fairscale, decord, av...
None of them carries arithmetic used by the MViT path; they are configuration /
registry / IO helpers.  We stub the *dependencies* (never the reference code)
so that `slowfast.models.build_model(cfg)` executes the reference's own
`MViT`, `MultiScaleBlock`, `MultiScaleAttention` and `attention_pool`.

Used by `oracle/make_golden.py` (to pin the oracle and to generate the fixtures
under tests/golden/) and by tests that are skipped when /root/reference is absent
(it does not exist on the GPU box).
"""
from __future__ import annotations

import ast
import copy
import importlib.abc
import importlib.machinery
import json
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
if os.path.dirname(_HERE) not in sys.path:
    sys.path.insert(0, os.path.dirname(_HERE))


def _find_reference() -> str:
    """$AICITY_REF, then the read-only checkout of the build container, then the git-ignored copy that
    `__graft_entry__.build()` vendors under oracle/_ref/ (the only one that exists on the GPU box)."""
    cands = [os.environ.get("AICITY_REF"), "/root/reference", os.path.join(_HERE, "_ref")]
    for c in cands:
        if c and os.path.isdir(os.path.join(c, "slowfast", "models")):
            return c
    return cands[1]


REFERENCE_ROOT = _find_reference()


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "slowfast", "models"))


from aicity_action_b200 import depshims  # noqa: E402  (dependency stand-ins are product code: the launcher needs them)
from aicity_action_b200.depshims import CfgNode, Registry  # noqa: E402,F401

_installed = False


def install():
    """Make `import slowfast` resolve to the reference with its missing deps stubbed."""
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    depshims.install()                  # only stubs what is genuinely missing
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    scripts = os.path.join(REFERENCE_ROOT, "scripts")
    if scripts not in sys.path:
        sys.path.append(scripts)
    _installed = True


def ref_cfg(yaml_name: str, overrides=None):
    """Reference cfg for configs/Aicity/<yaml_name> with NUM_GPUS=0 (CPU)."""
    install()
    from slowfast.config.defaults import get_cfg

    cfg = get_cfg()
    cfg.merge_from_file(os.path.join(REFERENCE_ROOT, "configs", "Aicity", yaml_name))
    cfg.NUM_GPUS = 0
    if overrides:
        cfg.merge_from_list(list(overrides))
    return cfg


def ref_build_model(cfg, seed=0):
    install()
    import torch
    from slowfast.models import build_model

    torch.manual_seed(seed)
    return build_model(cfg)
