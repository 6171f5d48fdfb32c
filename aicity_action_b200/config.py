"""Minimal cfg tree for the MViT path: the `MVIT.*`, `MODEL.*`, `DATA.*` keys the model reads
 with the reference's default values (slowfast/config/defaults.py:291-498,
726-754), plus the six `configs/Aicity/*.yaml` presets expressed as overrides.

`MViT(cfg)` accepts any object with the same attribute tree — the reference's fvcore `CfgNode`
included — so `tools/run_net.py` configs work unchanged; this class only exists so the package is
usable (bench, tests, sliding-window driver) without fvcore/yacs installed.
"""
from __future__ import annotations

import ast
import copy


class CfgNode(dict):
    """Attribute-access dict with yacs-style merge helpers."""

    def __init__(self, init=None):
        super().__init__()
        for k, v in (init or {}).items():
            self[k] = CfgNode(v) if isinstance(v, dict) else v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def clone(self):
        return copy.deepcopy(self)

    def __deepcopy__(self, memo):
        return CfgNode({k: copy.deepcopy(v, memo) for k, v in self.items()})

    @staticmethod
    def _lit(v):
        if isinstance(v, str):
            try:
                v = ast.literal_eval(v)
            except (ValueError, SyntaxError):
                return v
        return list(v) if isinstance(v, tuple) else v

    def merge_from_dict(self, other):
        for k, v in other.items():
            if isinstance(v, dict):
                self.setdefault(k, CfgNode())
                self[k].merge_from_dict(v)
            else:
                self[k] = self._lit(v)

    def merge_from_file(self, path):
        import yaml
        with open(path) as f:
            self.merge_from_dict(yaml.safe_load(f) or {})

    def merge_from_list(self, lst):
        assert len(lst) % 2 == 0, "override list must be KEY VALUE pairs"
        for key, val in zip(lst[0::2], lst[1::2]):
            node = self
            *parents, leaf = key.split(".")
            for p in parents:
                node = node[p]
            node[leaf] = self._lit(val)


def get_cfg() -> CfgNode:
    return CfgNode({
        "NUM_GPUS": 1, "NUM_SHARDS": 1, "SHARD_ID": 0, "RNG_SEED": 0, "DIST_BACKEND": "nccl",
        "DATA": {"NUM_FRAMES": 16, "SAMPLING_RATE": 4, "TRAIN_CROP_SIZE": 224, "TEST_CROP_SIZE": 224,
                 "INPUT_CHANNEL_NUM": [3]},
        "MODEL": {"ARCH": "mvit", "MODEL_NAME": "MViT", "NUM_CLASSES": 400, "DROPOUT_RATE": 0.5,
                  "HEAD_ACT": "softmax", "USE_HEAD_ACT_IN_TRAIN": False, "ACT_CHECKPOINT": False,
                  "USE_MULTI_HEAD": False, "MULTI_USE_MOCO": False, "USE_VICREG_LOSS": False,
                  "LOSS_FUNC": "cross_entropy"},
        "MVIT": {"MODE": "conv", "POOL_FIRST": False, "CLS_EMBED_ON": True, "PATCH_KERNEL": [3, 7, 7],
                 "PATCH_STRIDE": [2, 4, 4], "PATCH_PADDING": [2, 4, 4], "PATCH_2D": False, "EMBED_DIM": 96,
                 "NUM_HEADS": 1, "MLP_RATIO": 4.0, "QKV_BIAS": True, "DROPPATH_RATE": 0.1, "DEPTH": 16,
                 "NORM": "layernorm", "DIM_MUL": [], "HEAD_MUL": [], "POOL_KV_STRIDE": None,
                 "POOL_KV_STRIDE_ADAPTIVE": None, "POOL_Q_STRIDE": [], "POOL_KVQ_KERNEL": None,
                 "ZERO_DECAY_POS_CLS": True, "NORM_STEM": False, "SEP_POS_EMBED": False, "DROPOUT_RATE": 0.0,
                 "DIRECT_INPUT": False, "Q_POOL_RESIDUAL": False, "Q_POOL_ALL": False,
                 # default-off extension, NOT in the reference (upstream PySlowFast names)
                 "REL_POS_SPATIAL": False, "REL_POS_TEMPORAL": False, "REL_POS_ZERO_INIT": False,
                 "CHANNEL_EXPAND_FRONT": False, "POOL_SKIP_USE_CONV": False, "NO_NORM_BEFORE_AVG": False},
        "DETECTION": {"ENABLE": False, "USE_CUBE_PROP": False, "USE_SPATIAL_MAXPOOL_BEFORE_PROJ": False,
                      "ROI_XFORM_RESOLUTION": 7, "SPATIAL_SCALE_FACTOR": 16, "ALIGNED": True},
        "CONTRA": {"ENABLE": False, "embed_dim": 512, "use_MLP": False},
    })


def _aicity(depth16: bool, size: int, full: bool, frames: int, rate: int) -> dict:
    stages = [1, 3, 14] if depth16 else [2, 5, 21]
    return {
        "DATA": {"NUM_FRAMES": frames, "SAMPLING_RATE": rate, "TRAIN_CROP_SIZE": size, "TEST_CROP_SIZE": size,
                 "INPUT_CHANNEL_NUM": [3]},
        "MVIT": {"ZERO_DECAY_POS_CLS": False, "SEP_POS_EMBED": True, "DEPTH": 16 if depth16 else 24,
                 "NUM_HEADS": 1, "EMBED_DIM": 96, "PATCH_KERNEL": [3, 7, 7], "PATCH_STRIDE": [2, 4, 4],
                 "PATCH_PADDING": [1, 3, 3], "MLP_RATIO": 4.0, "QKV_BIAS": True,
                 "DROPPATH_RATE": 0.4 if depth16 else 0.3, "NORM": "layernorm", "MODE": "conv",
                 "CLS_EMBED_ON": False, "DIM_MUL": [[s, 2.0] for s in stages],
                 "HEAD_MUL": [[s, 2.0] for s in stages], "POOL_KVQ_KERNEL": [3, 3, 3],
                 "POOL_KV_STRIDE_ADAPTIVE": [1, 8, 8], "POOL_Q_STRIDE": [[s, 1, 2, 2] for s in stages],
                 "DROPOUT_RATE": 0.0, "CHANNEL_EXPAND_FRONT": True, "Q_POOL_ALL": full, "Q_POOL_RESIDUAL": full},
        "MODEL": {"NUM_CLASSES": 18, "ARCH": "mvit", "MODEL_NAME": "MViT", "LOSS_FUNC": "soft_cross_entropy",
                  "DROPOUT_RATE": 0.5},
    }


# name of the reference YAML (configs/Aicity/<name>.yaml) -> overrides
AICITY_PRESETS = {
    "MVITV2_B_16x4_CONV": _aicity(True, 224, False, 16, 4),
    "MVITV2_FULL_B_16x4_CONV": _aicity(True, 224, True, 16, 4),
    "MVITV2_FULL_B_16x4_CONV_448": _aicity(True, 448, True, 16, 4),
    "MVITV2_FULL_B_16x2_CONV_448": _aicity(True, 448, True, 16, 2),
    "MVITV2_FULL_B_32x3_CONV": _aicity(False, 224, True, 32, 3),
    "MVITV2_FULL_B_32x3_CONV_448": _aicity(False, 448, True, 32, 3),
}


def aicity_cfg(name: str, overrides=None) -> CfgNode:
    cfg = get_cfg()
    cfg.merge_from_dict(AICITY_PRESETS[name.replace(".yaml", "")])
    if overrides:
        cfg.merge_from_list(list(overrides))
    return cfg
