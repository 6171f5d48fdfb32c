"""Bind a `JunweiLiang/aicity_action` (PySlowFast fork) checkout to the B200 path without editing it.

    import aicity_action_b200.patch as b200
    b200.install()                      # before build_model(cfg)

* `MODEL_REGISTRY["MViT"]` (slowfast/models/build.py:35-36, registered at video_model_builder.py:794) is rebound to
  `aicity_action_b200.mvit.MViT`, so `build_model(cfg)` constructs, `.cuda()`s and (NUM_GPUS > 1) DDP-wraps the drop-in;
* `slowfast.models.attention.{attention_pool, MultiScaleAttention, MultiScaleBlock}` are rebound as well, so reference
  code that builds blocks directly (and the reference `MViT` class, if someone keeps it) gets the B200 modules.
Everything else of the reference (configs, data loading, checkpoints, train/test loops) runs unchanged;
`python -m aicity_action_b200.launch <script> ...` (launch.py) does this for the reference's entry scripts.
"""
from __future__ import annotations


def install(replace_model: bool = True, replace_blocks: bool = True, compute_dtype=None) -> dict:
    """Returns {'registry': bool, 'attention': bool}: what was rebound.  Raises ImportError if `slowfast` is not importable.
    `compute_dtype` ("auto" | "bf16" | "fp32" | None = leave $MVIT_B200_COMPUTE in charge): see attention.set_compute_dtype."""
    import importlib

    from . import attention as b200_attn
    from . import mvit as b200_mvit

    if compute_dtype is not None:
        b200_attn.set_compute_dtype(compute_dtype)
    done = {"registry": False, "attention": False}
    if replace_blocks:
        ref_attn = importlib.import_module("slowfast.models.attention")
        for name in ("attention_pool", "MultiScaleAttention", "MultiScaleBlock"):
            setattr(ref_attn, name, getattr(b200_attn, name))
        try:                                   # modules that did `from .attention import MultiScaleBlock` at import time
            vmb = importlib.import_module("slowfast.models.video_model_builder")
            vmb.MultiScaleBlock = b200_attn.MultiScaleBlock
        except ImportError:
            pass
        done["attention"] = True
    if replace_model:
        build = importlib.import_module("slowfast.models.build")
        reg = build.MODEL_REGISTRY
        obj_map = getattr(reg, "_obj_map", None)
        if obj_map is None:
            raise ImportError("slowfast.models.build.MODEL_REGISTRY has no _obj_map (unknown fvcore Registry version)")
        importlib.import_module("slowfast.models")      # make sure the reference registered its own models first
        obj_map["MViT"] = b200_mvit.MViT
        done["registry"] = True
    return done
