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

REFERENCE_ROOT = os.environ.get("AICITY_REF", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "slowfast", "models"))


# ----------------------------------------------------------------------------
# fvcore.common.config.CfgNode  (yacs-like attribute dict)
# ----------------------------------------------------------------------------
class CfgNode(dict):
    def __init__(self, init=None, key_list=None, new_allowed=False):
        super().__init__()
        for k, v in (init or {}).items():
            self[k] = CfgNode(v) if isinstance(v, dict) and not isinstance(v, CfgNode) else v

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError as e:
            raise AttributeError(name) from e

    def __setattr__(self, name, value):
        self[name] = value

    def clone(self):
        return copy.deepcopy(self)

    def __deepcopy__(self, memo):
        out = CfgNode()
        for k, v in self.items():
            out[k] = copy.deepcopy(v, memo)
        return out

    @staticmethod
    def _coerce(v):
        if isinstance(v, str):
            try:
                v = ast.literal_eval(v)
            except (ValueError, SyntaxError):
                return v
        if isinstance(v, tuple):
            v = list(v)
        return v

    def _merge(self, other):
        for k, v in other.items():
            if isinstance(v, dict):
                if k not in self or not isinstance(self[k], CfgNode):
                    self[k] = CfgNode()
                self[k]._merge(v)
            else:
                self[k] = self._coerce(v)

    def merge_from_file(self, path, allow_unsafe=False):
        import yaml

        with open(path) as f:
            self._merge(yaml.safe_load(f) or {})

    def merge_from_other_cfg(self, other):
        self._merge(other)

    def merge_from_list(self, lst):
        assert len(lst) % 2 == 0
        for key, val in zip(lst[0::2], lst[1::2]):
            node = self
            parts = key.split(".")
            for p in parts[:-1]:
                node = node[p]
            node[parts[-1]] = self._coerce(val)

    def dump(self, **kw):
        def plain(n):
            return {k: plain(v) if isinstance(v, dict) else v for k, v in n.items()}

        return json.dumps(plain(self), indent=1, default=str)

    def freeze(self):
        pass

    def defrost(self):
        pass


class Registry:
    def __init__(self, name):
        self._name = name
        self._obj_map = {}

    def register(self, obj=None):
        if obj is None:
            def deco(o):
                self._obj_map[o.__name__] = o
                return o
            return deco
        self._obj_map[obj.__name__] = obj
        return obj

    def get(self, name):
        if name not in self._obj_map:
            raise KeyError(f"No object named '{name}' found in '{self._name}' registry!")
        return self._obj_map[name]

    def __contains__(self, name):
        return name in self._obj_map


class _PathManager:
    def open(self, path, mode="r", **kw):
        return open(path, mode)

    def exists(self, p):
        return os.path.exists(p)

    def isfile(self, p):
        return os.path.isfile(p)

    def isdir(self, p):
        return os.path.isdir(p)

    def mkdirs(self, p):
        os.makedirs(p, exist_ok=True)

    def ls(self, p):
        return os.listdir(p)

    def get_local_path(self, p, **kw):
        return p

    def register_handler(self, *a, **k):
        pass


class _PathManagerFactory:
    _pm = {}

    @classmethod
    def get(cls, key="default", **kw):
        return cls._pm.setdefault(key, _PathManager())


class _Timer:
    def __init__(self):
        import time

        self._t = time.perf_counter
        self.reset()

    def reset(self):
        self._start = self._t()
        self._paused = None
        self._total_paused = 0.0

    def pause(self):
        self._paused = self._t()

    def is_paused(self):
        return self._paused is not None

    def resume(self):
        if self._paused is not None:
            self._total_paused += self._t() - self._paused
            self._paused = None

    def seconds(self):
        end = self._paused if self._paused is not None else self._t()
        return end - self._start - self._total_paused


def _checkpoint_wrapper(module, *a, **kw):
    """fairscale semantics: same module object back, forward recomputed in backward,
    no state_dict prefix (SURVEY.md §7 'fairscale semantics')."""
    import torch.utils.checkpoint as cp

    inner = module.forward

    def fwd(*args, **kwargs):
        return cp.checkpoint(inner, *args, use_reentrant=False, **kwargs)

    module.forward = fwd
    return module


class _Anything:
    """Permissive stand-in class for symbols that are imported but never used on the MViT path."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return None

    def __getattr__(self, n):
        if n.startswith("__"):
            raise AttributeError(n)
        return _Anything()


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        val = type(name, (_Anything,), {})
        setattr(self, name, val)
        return val


_STUB_ROOTS = (
    "fvcore", "iopath", "simplejson", "matplotlib", "detectron2", "pytorchvideo",
    "fairscale", "decord", "av", "bitsandbytes", "ftfy", "onnxruntime", "seaborn",
    "moviepy", "sklearn_stub",
)

_SPECIAL = {
    "fvcore.common.config": {"CfgNode": CfgNode},
    "fvcore.common.registry": {"Registry": Registry},
    "fvcore.common.timer": {"Timer": _Timer},
    "iopath.common.file_io": {"PathManagerFactory": _PathManagerFactory, "g_pathmgr": _PathManager()},
    "fairscale.nn.checkpoint": {"checkpoint_wrapper": _checkpoint_wrapper},
    "simplejson": {"dumps": lambda o, **k: json.dumps(o, **{kk: v for kk, v in k.items() if kk != "use_decimal"}),
                   "loads": json.loads},
}


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path, target=None):
        root = fullname.split(".")[0]
        if root in _STUB_ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        if fullname == "slowfast.visualization" or fullname.startswith("slowfast.visualization."):
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        for k, v in _SPECIAL.get(spec.name, {}).items():
            setattr(m, k, v)
        return m

    def exec_module(self, module):
        pass


_installed = False


def install():
    """Make `import slowfast` resolve to the reference with its missing deps stubbed."""
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    # only stub what is genuinely missing
    sys.meta_path.append(_StubFinder())
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
