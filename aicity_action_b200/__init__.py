"""aicity_action_b200 — B200-native MViTv2 multiscale-attention path (drop-in for the PySlowFast fork
JunweiLiang/aicity_action: slowfast/models/attention.py + the MViT of video_model_builder.py)."""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
__version__ = "0.1.0"
