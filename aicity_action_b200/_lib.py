"""ctypes binding of libmvit_b200.so (the C ABI declared in include/mvit_b200.h).

The library is built in-tree by `make` / `__graft_entry__.build()` into
aicity_action_b200/lib/.  There is no fallback of any kind: if the shared object is
missing or a call is rejected, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libmvit_b200.so")

F32, BF16 = 0, 1
POOL_CONV, POOL_MAX, POOL_AVG = 0, 1, 2
EPI_NONE, EPI_GELU, EPI_GELU_GRAD = 0, 1, 2
IMPL_AUTO, IMPL_SIMT, IMPL_TCGEN05 = 0, 1, 2

_p, _i, _i64, _f = C.c_void_p, C.c_int, C.c_int64, C.c_float

# name -> (restype, argtypes); mirrors include/mvit_b200.h one to one
SIGNATURES = {
    "mvit_abi_version": (_i, []),
    "mvit_last_error": (C.c_char_p, []),
    "mvit_device_supported": (_i, []),
    "mvit_device_fault": (_i, []),
    "mvit_layernorm_fwd": (_i, [_p, _p, _p, _p, _i64, _i, _f, _i, _p]),
    "mvit_linear_fwd": (_i, [_p, _p, _p, _p, _p, _i64, _p, _i64, _i, _i, _i64, _i64, _i64, _i, _i, _i, _p]),
    "mvit_linear_stat_parts": (_i, [_i64, _i, _i]),
    "mvit_linear_ln_fwd": (_i, [_p, _p, _p, _p, _p, _i, _f, _p, _p, _i64, _p, _p, _i64, _i, _i, _i64, _i64, _i, _p]),
    "mvit_patch_conv_stats_fwd": (_i, [_p, _p, _p, _p, _p, _p] + [_i] * 12 + [_p]),
    "mvit_mlp_fused_supported": (_i, [_i, _i, _i]),
    "mvit_mlp_fused_fwd": (_i, [_p, _p, _i, _f, _p, _p, _p, _p, _p, _p, _p, _i64, _i, _i, _p]),
    "mvit_im2col3d_fwd": (_i, [_p, _p] + [_i] * 16 + [_p]),
    "mvit_attention_pool_fwd": (_i, [_p, _i64, _i64, _i64, _p, _p, _p, _p, _i64, _i64, _i64] + [_i] * 14
                                + [_f, _i, _p]),
    "mvit_attention_pool_fwd_save": (_i, [_p, _i64, _i64, _i64, _p, _p, _p, _p, _i64, _i64, _i64, _p] + [_i] * 12
                                     + [_f, _i, _p]),
    "mvit_attention_pool_qkv_fwd": (_i, [_p, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _f, _i, _p]),
    "mvit_attention_fwd": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _f, _i, _i, _i, _p]),
    "mvit_attention_rel_fwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _f, _i, _i, _i, _p]),
    "mvit_relpos_operands_fwd": (_i, [_p, _p, _p, _p, _p, _p] + [_i] * 7 + [_f, _i, _p]),
    "mvit_pos_embed_add": (_i, [_p, _i, _p, _p, _p, _i, _i, _i, _i, _i, _p]),
    "mvit_mean_head_workspace_floats": (C.c_size_t, [_i, _i, _i]),
    "mvit_mean_head_fwd": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p]),
    "mvit_fold_clip_fwd": (_i, [_p, _i, _p] + [_i] * 9 + [_f, _f, _p]),
    "mvit_patch_conv_fwd": (_i, [_p, _p, _p, _p, _p] + [_i] * 12 + [_p]),
    "mvit_preprocess_u8_fwd": (_i, [_p, _p, _i, _i, _i, _i, _f, _f, _i, _p]),
    "mvit_resize_gather_u8": (_i, [_p, _i, _i, _i, _p, _i, _p, _i, _i, _i, _p]),
    "mvit_layernorm_bwd": (_i, [_p, _p, _p, _p, _p, _p, _i64, _i, _f, _i, _p]),
    "mvit_gelu_bwd": (_i, [_p, _p, _p, _i64, _i, _p]),
    "mvit_linear_wgrad": (_i, [_p, _p, _p, _p, _i64, _i, _i, _i, _i, _p]),
    "mvit_attention_bwd_workspace_floats": (C.c_size_t, [_i, _i, _i]),
    "mvit_attention_bwd": (_i, [_p] * 10 + [_i] * 5 + [_f, _i, _i, _i, _p]),
    "mvit_adamw_workspace_floats": (C.c_size_t, []),
    "mvit_adamw_hyper_floats": (C.c_size_t, []),
    "mvit_adamw_clip_step": (_i, [_p, _p, _p, _p, _p, _i64, _i64, _p, _p, _p]),
    "mvit_attention_pool_bwd": (_i, [_i, _p, _i64, _i64, _i64, _p, _p, _p, _p] + [_i] * 13 + [_p]),
}

_lock = threading.Lock()
_lib = None


class MvitLibraryError(RuntimeError):
    pass


def load():
    """Load (once) and return the ctypes handle; raises if the extension was not built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise MvitLibraryError(
                f"{LIB_PATH} not found: build the CUDA extension first (`make` at the repo root or "
                f"`python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError if the .so lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def device_fault_check():
    """Synchronises the device and raises if a tensor-core kernel abandoned an mbarrier wait since the last call (its
    results are invalid).  The kernels raise a device flag instead of trapping, so the CUDA context stays usable."""
    n = load().mvit_device_fault()
    if n != 0:
        msg = load().mvit_last_error()
        raise MvitLibraryError(f"device fault reported by libmvit_b200.so: {msg.decode() if msg else '?'}")


def check(rc: int, what: str):
    if rc != 0:
        msg = load().mvit_last_error()
        raise MvitLibraryError(f"{what} failed (rc={rc}): {msg.decode() if msg else '?'}")
