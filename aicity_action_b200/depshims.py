"""Stand-ins for third-party packages the reference imports at module top but this image lacks.

The reference (a PySlowFast fork) imports fvcore, iopath, simplejson, matplotlib, detectron2,
pytorchvideo, fairscale, decord, av ....  None of them carries arithmetic
of the MViT path: they are configuration / registry / IO helpers.  `install()` registers a meta-path
finder that serves a stub ONLY for the roots that are genuinely not importable here, so that the
reference's entry scripts (`tools/run_net.py`, `scripts/run_action_classification_temporal_inf.py`)
and `slowfast.models.build_model` run unmodified next to `aicity_action_b200.patch.install()`; see
`aicity_action_b200/launch.py`.  With the real packages installed this module does nothing.

No reference code lives here: the stubs restate the public behaviour of the third-party APIs
(yacs-style CfgNode, fvcore Registry / Timer, iopath PathManager, fairscale checkpoint_wrapper,
decord VideoReader).  `slowfast.visualization.*` is also served as a stub: the reference imports
`slowfast.visualization.tensorboard_vis` from `tools/train_net.py:22` but does not ship that package
(SURVEY.md D3).
"""
from __future__ import annotations

import ast
import copy
import importlib.abc
import importlib.machinery
import importlib.util
import json
import os
import sys
import types


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
    def open(self, path, mode="r", buffering=-1, **kw):
        if any(c in mode for c in "wa") and os.path.dirname(path):
            os.makedirs(os.path.dirname(path), exist_ok=True)
        return open(path, mode, buffering=buffering)

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

_ACTIVE_ROOTS = ()


def _simplejson_dumps(obj, use_decimal=False, **kw):
    """simplejson.dumps: Decimal values are written as plain numbers (the reference logs them, slowfast/utils/logging.py:97)."""
    import decimal

    def default(o):
        if isinstance(o, decimal.Decimal):
            return float(o)
        raise TypeError(f"Object of type {type(o).__name__} is not JSON serializable")

    return json.dumps(obj, default=default, **kw)

_SPECIAL = {
    "fvcore.common.config": {"CfgNode": CfgNode},
    "fvcore.common.registry": {"Registry": Registry},
    "fvcore.common.timer": {"Timer": _Timer},
    "iopath.common.file_io": {"PathManagerFactory": _PathManagerFactory, "g_pathmgr": _PathManager()},
    "fairscale.nn.checkpoint": {"checkpoint_wrapper": _checkpoint_wrapper},
    "simplejson": {"dumps": lambda o, **k: _simplejson_dumps(o, **k), "loads": json.loads},
}


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path, target=None):
        root = fullname.split(".")[0]
        if root in _ACTIVE_ROOTS:
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


class _NDArrayBatch:
    """decord's NDArray surface the reference touches: `.asnumpy()` and `.shape`."""

    def __init__(self, arr):
        self._a = arr
        self.shape = arr.shape

    def asnumpy(self):
        return self._a


class VideoReader:
    """decord.VideoReader(path, num_threads=0): `len`, `[i]`, `get_batch(idxs)`, `get_avg_fps()` — the four members
    scripts/module_wrapper.py:179-300 uses.  `*.npy` files hold decoded RGB frames [N, H, W, 3] uint8 (memory-mapped;
    optional `<file>.fps` text file, default 30); anything else is decoded with OpenCV."""

    def __init__(self, path, ctx=None, num_threads=0, **kw):
        import numpy as np

        self._path = path
        if str(path).endswith(".npy"):
            self._frames = np.load(path, mmap_mode="r")
            fps_file = str(path) + ".fps"
            self._fps = float(open(fps_file).read()) if os.path.exists(fps_file) else 30.0
        else:
            import cv2

            cap = cv2.VideoCapture(str(path))
            if not cap.isOpened():
                raise RuntimeError(f"cannot open video {path}")
            self._fps = float(cap.get(cv2.CAP_PROP_FPS)) or 30.0
            frames = []
            while True:
                ok, f = cap.read()
                if not ok:
                    break
                frames.append(cv2.cvtColor(f, cv2.COLOR_BGR2RGB))
            cap.release()
            self._frames = np.stack(frames)

    def __len__(self):
        return int(self._frames.shape[0])

    def __getitem__(self, i):
        import numpy as np

        return _NDArrayBatch(np.asarray(self._frames[int(i)]))

    def get_batch(self, idxs):
        import numpy as np

        return _NDArrayBatch(np.stack([np.asarray(self._frames[int(i)]) for i in idxs]))

    def get_avg_fps(self):
        return self._fps


_SPECIAL["decord"] = {"VideoReader": VideoReader}

_installed = False


def _missing(root: str) -> bool:
    if root in sys.modules:
        return False
    try:
        return importlib.util.find_spec(root) is None
    except (ImportError, ValueError):
        return True


def install() -> list:
    """Register stubs for the third-party roots that cannot be imported here; returns the stubbed roots."""
    global _installed, _ACTIVE_ROOTS
    if _installed:
        return list(_ACTIVE_ROOTS)
    _ACTIVE_ROOTS = tuple(r for r in _STUB_ROOTS if _missing(r))
    sys.meta_path.append(_StubFinder())
    _installed = True
    return list(_ACTIVE_ROOTS)
