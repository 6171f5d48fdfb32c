"""Thin torch-tensor wrappers over the C ABI (include/mvit_b200.h).

PyTorch is used here only for device memory (`torch.empty`) and the current CUDA stream; every
computation below is one call into libmvit_b200.so.  Tensors must live on a CUDA device: there
is no CPU implementation in the product (the CPU oracle lives in oracle/ and is test-only).
"""
from __future__ import annotations

import os

from typing import Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import (BF16, EPI_GELU, EPI_GELU_GRAD, EPI_NONE, F32, IMPL_AUTO, IMPL_SIMT, IMPL_TCGEN05, POOL_AVG,
                   POOL_CONV, POOL_MAX, check)

_DT = {torch.float32: F32, torch.bfloat16: BF16}
POOL_MODES = {"conv": POOL_CONV, "max": POOL_MAX, "avg": POOL_AVG}

# number of C-ABI kernel launches issued through this module (bench.py reports it as gpu_launches)
launch_count = 0

# Optional per-kernel timing: when `event_log` is a dict, every wrapper brackets its launch with CUDA
# events recorded on the launching stream and appends (category, start, end, work) to event_log[cat].
event_log = None


class _Timed:
    __slots__ = ("cat", "work", "ev0")

    def __init__(self, cat, work=0.0):
        self.cat, self.work, self.ev0 = cat, work, None

    def __enter__(self):
        if event_log is not None:
            self.ev0 = torch.cuda.Event(enable_timing=True)
            self.ev0.record()
        return self

    def __exit__(self, *exc):
        if self.ev0 is not None:
            ev1 = torch.cuda.Event(enable_timing=True)
            ev1.record()
            event_log.setdefault(self.cat, []).append((self.ev0, ev1, self.work))
        return False


def device_fault_check():
    """Raise if a tensor-core kernel abandoned an mbarrier wait since the last call (see mvit_device_fault).  Synchronises:
    call it where the host synchronises anyway."""
    _lib.device_fault_check()


def _dt(t: torch.Tensor) -> int:
    try:
        return _DT[t.dtype]
    except KeyError:
        raise TypeError(f"unsupported dtype {t.dtype}: the B200 path computes in float32 or bfloat16") from None


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.MvitLibraryError(
                "aicity_action_b200 kernels run on CUDA tensors only (no CPU fallback); got a CPU tensor")


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _f32c(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    if t is None:
        return None
    if t.dtype != torch.float32 or not t.is_contiguous():
        t = t.detach().float().contiguous()
    return t


def pooled_thw(thw: Sequence[int], kernel: Sequence[int], stride: Sequence[int]):
    """Output grid of Conv3d/MaxPool3d with padding k//2 and ceil_mode=False."""
    return [(n + 2 * (k // 2) - k) // s + 1 for n, k, s in zip(thw, kernel, stride)]


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float,
              out: Optional[torch.Tensor] = None) -> torch.Tensor:
    global launch_count
    _need_cuda(x, gamma, beta)
    x = x.contiguous()
    C = x.shape[-1]
    rows = x.numel() // C
    y = torch.empty_like(x) if out is None else out
    gamma, beta = _f32c(gamma), _f32c(beta)
    with _Timed("layernorm", 2.0 * x.numel() * x.element_size()):
        check(_lib.load().mvit_layernorm_fwd(_ptr(x), _ptr(gamma), _ptr(beta), _ptr(y), rows, C, float(eps),
                                             _dt(x), _stream()), "mvit_layernorm_fwd")
    launch_count += 1
    return y


def linear(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, *,
           residual: Optional[torch.Tensor] = None, row_scale: Optional[torch.Tensor] = None,
           gelu: bool = False, out: Optional[torch.Tensor] = None, impl: int = IMPL_AUTO,
           residual_row_period: int = 0, gelu_grad: bool = False) -> torch.Tensor:
    """y = epi(x·wᵀ + bias)·row_scale + residual ; x [..., K], w [N, K] (same dtype as x).
    residual_row_period = P > 0: `residual` is a [P, N] table and row m adds row m % P.
    gelu_grad: backward helper, y = residual · gelu'(x·wᵀ + bias) (`residual` = the upstream gradient)."""
    global launch_count
    _need_cuda(x, w, bias, residual, row_scale)
    x = x.contiguous()
    K = x.shape[-1]
    N = w.shape[0]
    assert w.shape[1] == K and w.dtype == x.dtype and w.is_contiguous(), "weight must be [N,K], contiguous, x.dtype"
    M = x.numel() // K
    y = torch.empty(x.shape[:-1] + (N,), dtype=x.dtype, device=x.device) if out is None else out
    assert y.is_contiguous() and y.numel() == M * N
    if residual is not None:
        residual = residual.contiguous()
        assert residual.numel() == (residual_row_period or M) * N and residual.dtype == x.dtype
    rows_per_sample = 0
    if row_scale is not None:
        row_scale = _f32c(row_scale).reshape(-1)
        assert M % row_scale.numel() == 0
        rows_per_sample = M // row_scale.numel()
    bias = _f32c(bias)
    with _Timed("linear", 2.0 * M * N * K):
        check(_lib.load().mvit_linear_fwd(_ptr(x), _ptr(w), _ptr(bias), _ptr(residual), _ptr(row_scale),
                                          rows_per_sample, _ptr(y), M, N, K, N, N, int(residual_row_period),
                                          EPI_GELU_GRAD if gelu_grad else (EPI_GELU if gelu else EPI_NONE), _dt(x), impl,
                                          _stream()),
              "mvit_linear_fwd")
    launch_count += 1
    return y


STATS_ATTR = "_b200_row_stats"      # tensor attribute: [parts, M, 2] fp32 (sum, sum of squares) of the tensor's rows


def ln_fold_enabled() -> bool:
    """LayerNorm folding of the eval path ($MVIT_B200_LN_FOLD, default on)."""
    return os.environ.get("MVIT_B200_LN_FOLD", "1") not in ("0", "false", "off")


def row_stats_of(x: torch.Tensor) -> Optional[torch.Tensor]:
    """Row statistics a producing GEMM attached to `x` (None when `x` did not come straight out of one)."""
    st = getattr(x, STATS_ATTR, None)
    if st is None or st.shape[1] != x.numel() // x.shape[-1] or st.device != x.device:
        return None
    return st


def linear_stats(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, *,
                 residual: Optional[torch.Tensor] = None, row_scale: Optional[torch.Tensor] = None) -> torch.Tensor:
    """ops.linear (no activation) that also emits the per-row (sum, sum of squares) of its bf16 output and attaches them
    to the result (`row_stats_of`): the producer half of the folded LayerNorm.  bf16 / tcgen05 only."""
    global launch_count
    _need_cuda(x, w, bias, residual, row_scale)
    x = x.contiguous()
    K, N = x.shape[-1], w.shape[0]
    assert x.dtype == torch.bfloat16 and w.dtype == x.dtype and w.shape[1] == K and w.is_contiguous()
    M = x.numel() // K
    y = torch.empty(x.shape[:-1] + (N,), dtype=x.dtype, device=x.device)
    if residual is not None:
        residual = residual.contiguous()
        assert residual.numel() == M * N and residual.dtype == x.dtype
    rows_per_sample = 0
    if row_scale is not None:
        row_scale = _f32c(row_scale).reshape(-1)
        assert M % row_scale.numel() == 0
        rows_per_sample = M // row_scale.numel()
    bias = _f32c(bias)
    lib = _lib.load()
    parts = lib.mvit_linear_stat_parts(M, N, K)
    stats = torch.empty((parts, M, 2), dtype=torch.float32, device=x.device)
    with _Timed("linear", 2.0 * M * N * K):
        check(lib.mvit_linear_ln_fwd(_ptr(x), _ptr(w), _ptr(bias), None, None, 0, 0.0, _ptr(residual), _ptr(row_scale),
                                     rows_per_sample, _ptr(y), _ptr(stats), M, N, K, N, N, EPI_NONE, _stream()),
              "mvit_linear_ln_fwd")
    launch_count += 1
    if parts > 4:        # the consumer stages at most 4 slices per row (small M on 128x96 tiles: wide rows span more N tiles)
        stats = stats.sum(dim=0, keepdim=True)
    setattr(y, STATS_ATTR, stats)
    return y


def linear_ln(x: torch.Tensor, stats: torch.Tensor, w_folded: torch.Tensor, bias_folded: torch.Tensor,
              colsum: torch.Tensor, eps: float, *, gelu: bool = False) -> torch.Tensor:
    """LayerNorm(x)·Wᵀ + b from the RAW rows `x` and their statistics: the consumer half of the folded LayerNorm
    (`w_folded`, `bias_folded`, `colsum` from weights.folded_ln_linear)."""
    global launch_count
    _need_cuda(x, stats, w_folded, bias_folded, colsum)
    x = x.contiguous()
    K, N = x.shape[-1], w_folded.shape[0]
    M = x.numel() // K
    assert x.dtype == torch.bfloat16 and w_folded.dtype == x.dtype and w_folded.shape[1] == K and w_folded.is_contiguous()
    assert stats.dtype == torch.float32 and stats.is_contiguous() and tuple(stats.shape[1:]) == (M, 2)
    assert bias_folded.dtype == torch.float32 and colsum.dtype == torch.float32
    y = torch.empty(x.shape[:-1] + (N,), dtype=x.dtype, device=x.device)
    with _Timed("linear", 2.0 * M * N * K):
        check(_lib.load().mvit_linear_ln_fwd(_ptr(x), _ptr(w_folded), _ptr(bias_folded), _ptr(colsum), _ptr(stats),
                                             stats.shape[0], float(eps), None, None, 0, _ptr(y), None, M, N, K, N, N,
                                             EPI_GELU if gelu else EPI_NONE, _stream()), "mvit_linear_ln_fwd")
    launch_count += 1
    return y


def mlp_fused_enabled() -> bool:
    """Single-kernel MLP for the C <= 192 blocks on the eval path ($MVIT_B200_MLP_FUSED, default on)."""
    return os.environ.get("MVIT_B200_MLP_FUSED", "1") not in ("0", "false", "off")


def mlp_fused_supported(C: int, H: int, C_out: int) -> bool:
    return bool(_lib.load().mvit_mlp_fused_supported(int(C), int(H), int(C_out)))


def mlp_fused(x: torch.Tensor, stats: torch.Tensor, w1_folded: torch.Tensor, b1_folded: torch.Tensor, colsum1: torch.Tensor,
              w2: torch.Tensor, b2: Optional[torch.Tensor], eps: float) -> torch.Tensor:
    """y = x + fc2(GELU(fc1(LayerNorm(x)))) in one kernel (the hidden activation never reaches HBM); `x` is the raw block
    stream with row statistics `stats`; the result carries its own row statistics (`row_stats_of`)."""
    global launch_count
    _need_cuda(x, stats, w1_folded, b1_folded, colsum1, w2, b2)
    x = x.contiguous()
    C = x.shape[-1]
    H = w1_folded.shape[0]
    M = x.numel() // C
    assert x.dtype == torch.bfloat16 and w1_folded.dtype == x.dtype and w2.dtype == x.dtype
    assert w1_folded.shape == (H, C) and w2.shape == (C, H) and w1_folded.is_contiguous() and w2.is_contiguous()
    assert stats.dtype == torch.float32 and stats.is_contiguous() and tuple(stats.shape[1:]) == (M, 2)
    y = torch.empty_like(x)
    out_stats = torch.empty((1, M, 2), dtype=torch.float32, device=x.device)
    b2 = _f32c(b2)
    with _Timed("linear", 4.0 * M * C * H):
        check(_lib.load().mvit_mlp_fused_fwd(_ptr(x), _ptr(stats), stats.shape[0], float(eps), _ptr(w1_folded), _ptr(b1_folded),
                                             _ptr(colsum1), _ptr(w2), _ptr(b2), _ptr(y), _ptr(out_stats), M, C, H, _stream()),
              "mvit_mlp_fused_fwd")
    launch_count += 1
    setattr(y, STATS_ATTR, out_stats)
    return y


def attention_pool_strided(src: torch.Tensor, src_offset: int, in_strides: Tuple[int, int, int], B: int,
                           heads: int, d: int, thw: Sequence[int], kernel: Sequence[int],
                           stride: Sequence[int], mode: str, weight: Optional[torch.Tensor],
                           gamma: Optional[torch.Tensor], beta: Optional[torch.Tensor], eps: float,
                           has_cls: bool, out: torch.Tensor, out_strides: Tuple[int, int, int]):
    """Raw form: reads in[b,l,head,:] at src_offset + b*bs + l*ls + head*hs (element strides)."""
    global launch_count
    _need_cuda(src, out, weight, gamma, beta)
    T, H, W = thw
    weight, gamma, beta = _f32c(weight), _f32c(gamma), _f32c(beta)
    if weight is not None:
        weight = weight.reshape(d, -1)
    base = src.data_ptr() + src_offset * src.element_size()
    # algorithmic bytes: every input element of this q/k/v (or skip) tensor read once + output written once
    work = float(B * heads * d * T * H * W + out.numel()) * src.element_size()
    with _Timed("pool_" + mode, work):
        check(_lib.load().mvit_attention_pool_fwd(
            base, in_strides[0], in_strides[1], in_strides[2], _ptr(weight), _ptr(gamma), _ptr(beta),
            _ptr(out), out_strides[0], out_strides[1], out_strides[2], B, heads, d, T, H, W,
            kernel[0], kernel[1], kernel[2], stride[0], stride[1], stride[2], POOL_MODES[mode],
            1 if has_cls else 0, float(eps), _dt(src), _stream()), "mvit_attention_pool_fwd")
    launch_count += 1


def attention_pool_heads(x: torch.Tensor, thw: Sequence[int], kernel: Sequence[int], stride: Sequence[int], *,
                         mode: str = "conv", weight: Optional[torch.Tensor] = None,
                         ln: Optional[Tuple[torch.Tensor, torch.Tensor, float]] = None,
                         has_cls: bool = False):
    """x: any strided view [B, heads, L, d] with unit channel stride -> contiguous [B, heads, L', d]."""
    B, heads, L, d = x.shape
    assert x.stride(3) == 1
    out_thw = pooled_thw(thw, kernel, stride)
    Lo = out_thw[0] * out_thw[1] * out_thw[2] + (1 if has_cls else 0)
    out = torch.empty((B, heads, Lo, d), dtype=x.dtype, device=x.device)
    g, b, eps = ln if ln is not None else (None, None, 0.0)
    attention_pool_strided(x, 0, (x.stride(0), x.stride(2), x.stride(1)), B, heads, d, thw, kernel, stride,
                           mode, weight, g, b, eps, has_cls, out, (heads * Lo * d, d, Lo * d))
    # note: `x` may be a view with a storage offset; data_ptr() already includes it
    return out, out_thw


def pool_qkv_supported(qkv: torch.Tensor, heads: int, strides) -> bool:
    """Whether mvit_attention_pool_qkv_fwd applies: bf16 contiguous qkv GEMM output, head_dim 96, 3x3x3 conv pools of
    stride (1, s, s) with s in {1, 2, 4, 8} on all three tensors (every block of the shipped FULL configurations)."""
    return (qkv.dtype == torch.bfloat16 and qkv.is_contiguous() and qkv.shape[-1] == 3 * heads * 96
            and qkv.data_ptr() % 16 == 0 and all(s is not None and s[0] == 1 and s[1] == s[2] and s[1] in (1, 2, 4, 8)
                                                 for s in strides))


def attention_pool_qkv(qkv: torch.Tensor, heads: int, thw: Sequence[int], weights, lns, strides, save_pre: bool = False,
                       only: Optional[Sequence[int]] = None):
    """qkv: [B, N, 3*heads*96] bf16 (the qkv GEMM output, read in place); weights[i]: [96,1,3,3,3] conv filters of q / k / v;
    lns[i]: (gamma, beta) or None; strides[i]: (1, s, s).  One C-ABI call; returns ([q, k, v] pooled + LayerNorm-ed, contiguous
    [B, heads, L', 96]), their [T, H', W'] grids, and (save_pre) the pre-LayerNorm conv outputs.  `only`: indices of the
    tensors to pool in this call (the others come back as None) — lets a caller put q and k/v on different streams."""
    global launch_count
    import ctypes as C
    _need_cuda(qkv, *weights)
    B, N, _ = qkv.shape
    T, H, W = thw
    assert N == T * H * W
    outs, pres, grids, keep = [], [], [], []
    sel = set(range(3)) if only is None else set(only)
    for i, s in enumerate(strides):
        g = pooled_thw(thw, [3, 3, 3], s)
        grids.append(g)
        outs.append(torch.empty((B, heads, g[0] * g[1] * g[2], 96), dtype=qkv.dtype, device=qkv.device) if i in sel else None)
        pres.append(torch.empty_like(outs[-1]) if (save_pre and i in sel) else None)
    w32 = [_f32c(w).reshape(96, 27) for w in weights]
    gam = [_f32c(ln[0]) if ln is not None else None for ln in lns]
    bet = [_f32c(ln[1]) if ln is not None else None for ln in lns]
    keep += w32 + gam + bet
    arr = lambda ts: (C.c_void_p * 3)(*[_ptr(t) for t in ts])
    sarr = (C.c_int * 3)(*[int(s[1]) if i in sel else 0 for i, s in enumerate(strides)])
    eps = next((float(ln[2]) for ln in lns if ln is not None), 0.0)
    n_tma = len({int(s[1]) for i, s in enumerate(strides) if s[1] <= 2 and i in sel})
    n_old = sum(1 for i, s in enumerate(strides) if s[1] > 2 and i in sel)
    work = float(qkv.numel() // 3 * len(sel) + sum(o.numel() for o in outs if o is not None) * (2 if save_pre else 1)) \
        * qkv.element_size()
    with _Timed("pool_conv", work):
        check(_lib.load().mvit_attention_pool_qkv_fwd(_ptr(qkv), B, heads, T, H, W, arr(w32), arr(gam), arr(bet), sarr,
                                                      arr(outs), arr(pres) if save_pre else None, eps, _dt(qkv), _stream()),
              "mvit_attention_pool_qkv_fwd")
    launch_count += n_tma + n_old
    return outs, grids, pres


def pool_save_supported(x: torch.Tensor, kernel: Sequence[int], stride: Sequence[int]) -> bool:
    """Whether mvit_attention_pool_fwd_save applies to this q/k/v view (the tuned kernel's conditions)."""
    es = x.element_size()
    return (list(kernel) == [3, 3, 3] and x.shape[3] == 96 and stride[0] == 1 and stride[1] == stride[2]
            and stride[1] in (1, 2, 4, 8) and x.data_ptr() % 16 == 0
            and all((x.stride(i) * es) % 16 == 0 for i in range(3)))


def attention_pool_heads_save(x: torch.Tensor, thw: Sequence[int], kernel: Sequence[int], stride: Sequence[int],
                              weight: torch.Tensor, ln: Tuple[torch.Tensor, torch.Tensor, float]):
    """Training forward of conv pooling + LayerNorm: -> (pooled [B, heads, L', d], pre-LayerNorm values, thw')."""
    global launch_count
    _need_cuda(x, weight, ln[0], ln[1])
    B, heads, L, d = x.shape
    out_thw = pooled_thw(thw, kernel, stride)
    Lo = out_thw[0] * out_thw[1] * out_thw[2]
    out = torch.empty((B, heads, Lo, d), dtype=x.dtype, device=x.device)
    pre = torch.empty_like(out)
    w = _f32c(weight).reshape(d, -1)
    with _Timed("pool_conv", float(B * heads * d * L + 2 * out.numel()) * x.element_size()):
        check(_lib.load().mvit_attention_pool_fwd_save(
            _ptr(x), x.stride(0), x.stride(2), x.stride(1), _ptr(w), _ptr(_f32c(ln[0])), _ptr(_f32c(ln[1])), _ptr(out),
            heads * Lo * d, d, Lo * d, _ptr(pre), B, heads, d, thw[0], thw[1], thw[2], *kernel, *stride, float(ln[2]),
            _dt(x), _stream()), "mvit_attention_pool_fwd_save")
    launch_count += 1
    return out, pre, out_thw


def attention_pool_tokens(x: torch.Tensor, thw: Sequence[int], kernel: Sequence[int], stride: Sequence[int], *,
                          mode: str = "max", has_cls: bool = False, d: int = 96):
    """x: [B, L, C] channels-last tokens -> [B, L', C] (skip-path pooling, attention.py:427-432)."""
    x = x.contiguous()
    B, L, Cc = x.shape
    if Cc % d != 0:
        d = 32 if Cc % 32 == 0 else None
        if d is None:
            raise _lib.MvitLibraryError(f"skip pooling needs channels % 32 == 0, got {Cc}")
    heads = Cc // d
    out_thw = pooled_thw(thw, kernel, stride)
    Lo = out_thw[0] * out_thw[1] * out_thw[2] + (1 if has_cls else 0)
    out = torch.empty((B, Lo, Cc), dtype=x.dtype, device=x.device)
    attention_pool_strided(x, 0, (L * Cc, Cc, d), B, heads, d, thw, kernel, stride, mode, None, None, None,
                           0.0, has_cls, out, (Lo * Cc, Cc, d))
    return out, out_thw


def relpos_operands(q: torch.Tensor, q_thw: Sequence[int], k_thw: Sequence[int], rel_h: torch.Tensor,
                    rel_w: torch.Tensor, rel_t: torch.Tensor, scale: float):
    """Default-off relative-position operands (NOT in the reference): from the pooled, unscaled
    q [B,h,Lq,96] and the tables rel_pos_h/w/t -> (q_ext [B,h,Lq,64], k_ext [B,h,Lk,64]) for `attention(..., rel=...)`."""
    global launch_count
    _need_cuda(q, rel_h, rel_w, rel_t)
    q = q.contiguous()
    B, h, Lq, d = q.shape
    qt, qh, qw = q_thw
    kt, kh, kw = k_thw
    assert d == 96 and Lq == qt * qh * qw, "relpos: q must be [B, heads, qt*qh*qw, 96] (no cls token)"
    assert rel_h.shape == (2 * max(qh, kh) - 1, d) and rel_w.shape == (2 * max(qw, kw) - 1, d) \
        and rel_t.shape == (2 * max(qt, kt) - 1, d), "relpos: table shapes are [2*max(q,k)-1, 96]"
    Lk = kt * kh * kw
    q_ext = torch.empty((B, h, Lq, 64), dtype=q.dtype, device=q.device)
    k_ext = torch.empty((B, h, Lk, 64), dtype=q.dtype, device=q.device)
    rh, rw, rt = _f32c(rel_h), _f32c(rel_w), _f32c(rel_t)
    with _Timed("relpos", 2.0 * B * h * Lq * (kt + kh + kw) * d):
        check(_lib.load().mvit_relpos_operands_fwd(_ptr(q), _ptr(rh), _ptr(rw), _ptr(rt), _ptr(q_ext), _ptr(k_ext), B * h,
                                                   qt, qh, qw, kt, kh, kw, float(scale), _dt(q), _stream()),
              "mvit_relpos_operands_fwd")
    launch_count += 2
    return q_ext, k_ext


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, scale: float, add_q: bool, *,
              want_lse: bool = False, impl: int = IMPL_AUTO, rel=None):
    """q [B,h,Lq,96], k/v [B,h,Lk,96] contiguous -> out [B, Lq, h*96] (+ lse [B,h,Lq] fp32).
    rel = (q_ext, k_ext) from `relpos_operands`: scores = scale*(q·k + q_ext·k_ext) (default-off relative-position bias)."""
    global launch_count
    _need_cuda(q, k, v)
    q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
    B, h, Lq, d = q.shape
    Lk = k.shape[2]
    assert k.shape == (B, h, Lk, d) and v.shape == (B, h, Lk, d) and q.dtype == k.dtype == v.dtype
    out = torch.empty((B, Lq, h * d), dtype=q.dtype, device=q.device)
    lse = torch.empty((B, h, Lq), dtype=torch.float32, device=q.device) if want_lse else None
    if rel is not None:
        qe, ke = rel
        assert qe.shape == (B, h, Lq, 64) and ke.shape == (B, h, Lk, 64) and qe.dtype == q.dtype == ke.dtype
        assert qe.is_contiguous() and ke.is_contiguous()
        with _Timed("attention", 4.0 * B * h * Lq * Lk * d + 2.0 * B * h * Lq * Lk * 64):
            check(_lib.load().mvit_attention_rel_fwd(_ptr(q), _ptr(k), _ptr(v), _ptr(qe), _ptr(ke), _ptr(out), _ptr(lse),
                                                     B, h, Lq, Lk, d, float(scale), 1 if add_q else 0, _dt(q), impl,
                                                     _stream()), "mvit_attention_rel_fwd")
        launch_count += 1
        return (out, lse) if want_lse else out
    with _Timed("attention", 4.0 * B * h * Lq * Lk * d):
        check(_lib.load().mvit_attention_fwd(_ptr(q), _ptr(k), _ptr(v), _ptr(out), _ptr(lse), B, h, Lq, Lk, d,
                                             float(scale), 1 if add_q else 0, _dt(q), impl, _stream()),
              "mvit_attention_fwd")
    launch_count += 1
    return (out, lse) if want_lse else out


def pos_embed_add(tokens: torch.Tensor, pos_spatial: torch.Tensor, pos_temporal: torch.Tensor, T: int,
                  out_dtype: torch.dtype) -> torch.Tensor:
    """tokens [B, T*HW, C] (fp32/bf16, contiguous) + separable pos-embed -> [B, T*HW, C] in out_dtype."""
    global launch_count
    _need_cuda(tokens, pos_spatial, pos_temporal)
    tokens = tokens.contiguous()
    B, N, Cc = tokens.shape
    HW = N // T
    out = torch.empty((B, N, Cc), dtype=out_dtype, device=tokens.device)
    ps, pt = _f32c(pos_spatial).reshape(HW, Cc), _f32c(pos_temporal).reshape(T, Cc)
    check(_lib.load().mvit_pos_embed_add(_ptr(tokens), _dt(tokens), _ptr(ps), _ptr(pt), _ptr(out), B, T, HW,
                                         Cc, _DT[out_dtype], _stream()), "mvit_pos_embed_add")
    launch_count += 1
    return out


def mean_head(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], softmax: bool,
              want_feat: bool = False):
    """x [B, L, C] -> fp32 [B, classes] (token mean -> Linear -> optional softmax)."""
    global launch_count
    _need_cuda(x, w, bias)
    x = x.contiguous()
    B, L, Cc = x.shape
    w, bias = _f32c(w), _f32c(bias)
    n = w.shape[0]
    out = torch.empty((B, n), dtype=torch.float32, device=x.device)
    feat = torch.empty((B, Cc), dtype=torch.float32, device=x.device) if want_feat else None
    lib = _lib.load()
    ws = torch.empty((lib.mvit_mean_head_workspace_floats(B, L, Cc),), dtype=torch.float32, device=x.device)
    check(lib.mvit_mean_head_fwd(_ptr(x), _ptr(w), _ptr(bias), _ptr(feat), _ptr(out), _ptr(ws), B, L, Cc, n,
                                 1 if softmax else 0, _dt(x), _stream()), "mvit_mean_head_fwd")
    launch_count += 2 if L > 32 else 1
    return (out, feat) if want_feat else out


def preprocess_u8(frames: torch.Tensor, dtype: torch.dtype, mean: float = 0.45, std: float = 0.225) -> torch.Tensor:
    """uint8 frames [B, T, H, W, 3] -> normalised clip [B, 3, T, H, W] (module_wrapper.py:326-346)."""
    global launch_count
    _need_cuda(frames)
    assert frames.dtype == torch.uint8 and frames.ndim == 5 and frames.shape[-1] == 3
    frames = frames.contiguous()
    B, T, H, W, _ = frames.shape
    out = torch.empty((B, 3, T, H, W), dtype=dtype, device=frames.device)
    check(_lib.load().mvit_preprocess_u8_fwd(_ptr(frames), _ptr(out), B, T, H, W, float(mean), float(std),
                                             _DT[dtype], _stream()), "mvit_preprocess_u8_fwd")
    launch_count += 1
    return out


def resize_gather_u8(frames: torch.Tensor, frame_idx: Optional[torch.Tensor], out_hw: Tuple[int, int],
                     out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """frames [n, H, W, C] uint8 (device) -> [len(frame_idx) or n, out_h, out_w, C] uint8: output frame i is source frame
    frame_idx[i] resized exactly as `cv2.resize(frame, (out_w, out_h), interpolation=cv2.INTER_LINEAR)` does
    (scripts/utils.py:207-211).  `frame_idx`: int32 device tensor or None (identity)."""
    global launch_count
    _need_cuda(frames, frame_idx, out)
    assert frames.dtype == torch.uint8 and frames.ndim == 4 and frames.is_contiguous()
    n, H, W, Cc = frames.shape
    n_out = n if frame_idx is None else frame_idx.numel()
    if frame_idx is not None:
        assert frame_idx.dtype == torch.int32 and frame_idx.is_contiguous()
    oh, ow = out_hw
    if out is None:
        out = torch.empty((n_out, oh, ow, Cc), dtype=torch.uint8, device=frames.device)
    assert out.dtype == torch.uint8 and out.is_contiguous() and out.numel() == n_out * oh * ow * Cc
    with _Timed("resize_u8", float(frames.numel() + out.numel())):
        check(_lib.load().mvit_resize_gather_u8(_ptr(frames), n, H, W, _ptr(frame_idx), n_out, _ptr(out), oh, ow, Cc,
                                                _stream()), "mvit_resize_gather_u8")
    launch_count += 1
    return out


def im2col3d(clip: torch.Tensor, kernel: Sequence[int], stride: Sequence[int], padding: Sequence[int], Kp: int):
    """clip [B, C, T, H, W] -> patch matrix [B*To*Ho*Wo, Kp] (zero padded), see mvit_im2col3d_fwd."""
    global launch_count
    _need_cuda(clip)
    clip = clip.contiguous()
    B, Cc, T, H, W = clip.shape
    To, Ho, Wo = [(n + 2 * p - k) // s + 1 for n, k, s, p in zip((T, H, W), kernel, stride, padding)]
    out = torch.empty((B * To * Ho * Wo, Kp), dtype=clip.dtype, device=clip.device)
    with _Timed("im2col", float(out.numel() + clip.numel()) * clip.element_size()):
        check(_lib.load().mvit_im2col3d_fwd(_ptr(clip), _ptr(out), B, Cc, T, H, W, *kernel, *stride, *padding, Kp,
                                            _dt(clip), _stream()), "mvit_im2col3d_fwd")
    launch_count += 1
    return out, [To, Ho, Wo]


def fold_clip(clip: torch.Tensor, stride: Sequence[int], Cf: int, mean: float = 0.45, std: float = 0.225) -> torch.Tensor:
    """Space-to-depth of a clip by the patch-embed stride (see mvit_fold_clip_fwd).  `clip` is a channels-first
    fp32/bf16 clip [B, C, T, H, W] or uint8 frames [B, T, H, W, C] (normalised on the fly); returns bf16
    [B, T/st, H/sh, W/sw, Cf]."""
    global launch_count
    _need_cuda(clip)
    clip = clip.contiguous()
    if clip.dtype == torch.uint8:
        B, T, H, W, Cc = clip.shape
        kind = 2
    else:
        B, Cc, T, H, W = clip.shape
        kind = {torch.float32: 0, torch.bfloat16: 1}[clip.dtype]
    st, sh, sw = stride
    out = torch.empty((B, T // st, H // sh, W // sw, Cf), dtype=torch.bfloat16, device=clip.device)
    with _Timed("fold_clip", float(clip.numel() * clip.element_size() + out.numel() * 2)):
        check(_lib.load().mvit_fold_clip_fwd(_ptr(clip), kind, _ptr(out), B, Cc, T, H, W, st, sh, sw, Cf, float(mean),
                                             float(std), _stream()), "mvit_fold_clip_fwd")
    launch_count += 1
    return out


def patch_conv(folded: torch.Tensor, wf: torch.Tensor, bias: Optional[torch.Tensor], pos: Optional[torch.Tensor],
               taps: Sequence[int], lows: Sequence[int], want_stats: bool = False) -> torch.Tensor:
    """Implicit-GEMM patch embedding on the folded clip: returns tokens [B, Tf*Hf*Wf, N] (bf16).  want_stats: also emit
    the tokens' row statistics (`row_stats_of`) for the folded LayerNorm of the first block."""
    global launch_count
    _need_cuda(folded, wf, bias, pos)
    B, Tf, Hf, Wf, Cf = folded.shape
    N = wf.shape[0]
    assert folded.dtype == torch.bfloat16 and wf.dtype == torch.bfloat16 and wf.is_contiguous()
    assert wf.shape[1] == taps[0] * taps[1] * taps[2] * Cf
    out = torch.empty((B, Tf * Hf * Wf, N), dtype=torch.bfloat16, device=folded.device)
    bias = _f32c(bias)
    if want_stats:
        stats = torch.empty(((N + 95) // 96, B * Tf * Hf * Wf, 2), dtype=torch.float32, device=folded.device)
        with _Timed("patch_conv", 2.0 * B * Tf * Hf * Wf * N * wf.shape[1]):
            check(_lib.load().mvit_patch_conv_stats_fwd(_ptr(folded), _ptr(wf), _ptr(bias), _ptr(pos), _ptr(out), _ptr(stats),
                                                        B, Tf, Hf, Wf, Cf, taps[0], taps[1], taps[2], lows[0], lows[1],
                                                        lows[2], N, _stream()), "mvit_patch_conv_stats_fwd")
        launch_count += 1
        setattr(out, STATS_ATTR, stats)
        return out
    with _Timed("patch_conv", 2.0 * B * Tf * Hf * Wf * N * wf.shape[1]):
        check(_lib.load().mvit_patch_conv_fwd(_ptr(folded), _ptr(wf), _ptr(bias), _ptr(pos), _ptr(out), B, Tf, Hf, Wf, Cf,
                                              taps[0], taps[1], taps[2], lows[0], lows[1], lows[2], N, _stream()),
              "mvit_patch_conv_fwd")
    launch_count += 1
    return out


# ------------------------------------------------------------------------------------------------ backward
def layernorm_bwd(x: torch.Tensor, gamma: torch.Tensor, dy: torch.Tensor, eps: float):
    """-> (dx like x, dgamma fp32 [C], dbeta fp32 [C])."""
    global launch_count
    _need_cuda(x, gamma, dy)
    x, dy = x.contiguous(), dy.contiguous()
    assert x.shape == dy.shape and x.dtype == dy.dtype
    C = x.shape[-1]
    rows = x.numel() // C
    dx = torch.empty_like(x)
    dgb = torch.zeros((2, C), dtype=torch.float32, device=x.device)
    with _Timed("layernorm_bwd", 3.0 * x.numel() * x.element_size()):
        check(_lib.load().mvit_layernorm_bwd(_ptr(x), _ptr(_f32c(gamma)), _ptr(dy), _ptr(dx), _ptr(dgb[0]), _ptr(dgb[1]),
                                             rows, C, float(eps), _dt(x), _stream()), "mvit_layernorm_bwd")
    launch_count += 1
    return dx, dgb[0], dgb[1]


def gelu_bwd(pre: torch.Tensor, dy: torch.Tensor) -> torch.Tensor:
    global launch_count
    _need_cuda(pre, dy)
    pre, dy = pre.contiguous(), dy.contiguous()
    assert pre.shape == dy.shape and pre.dtype == dy.dtype
    out = torch.empty_like(pre)
    with _Timed("gelu_bwd", 3.0 * pre.numel() * pre.element_size()):
        check(_lib.load().mvit_gelu_bwd(_ptr(pre), _ptr(dy), _ptr(out), pre.numel(), _dt(pre), _stream()), "mvit_gelu_bwd")
    launch_count += 1
    return out


def linear_wgrad(dy: torch.Tensor, x: torch.Tensor, want_bias: bool, impl: int = IMPL_AUTO):
    """dy [..., N], x [..., K] -> (dw fp32 [N, K], db fp32 [N] | None)."""
    global launch_count
    _need_cuda(dy, x)
    dy, x = dy.contiguous(), x.contiguous()
    N, K = dy.shape[-1], x.shape[-1]
    M = x.numel() // K
    assert dy.numel() == M * N and dy.dtype == x.dtype
    dw = torch.zeros((N, K), dtype=torch.float32, device=x.device)
    db = torch.zeros((N,), dtype=torch.float32, device=x.device) if want_bias else None
    with _Timed("linear_wgrad", 2.0 * M * N * K):
        check(_lib.load().mvit_linear_wgrad(_ptr(dy), _ptr(x), _ptr(dw), _ptr(db), M, N, K, _dt(x), impl, _stream()),
              "mvit_linear_wgrad")
    launch_count += 1
    return dw, db


def attention_bwd(q, k, v, out, dout, lse, scale: float, add_q: bool, impl: int = IMPL_AUTO):
    """-> (dq, dk, dv) in q.dtype, shapes of q / k / v."""
    global launch_count
    _need_cuda(q, k, v, out, dout, lse)
    dout = dout.contiguous()
    B, h, Lq, d = q.shape
    Lk = k.shape[2]
    dq = torch.empty_like(q)
    dkv = torch.zeros((2, B, h, Lk, d), dtype=torch.float32, device=q.device)
    lib = _lib.load()
    ws = None
    if q.dtype == torch.bfloat16 and impl != IMPL_SIMT:
        ws = torch.empty((lib.mvit_attention_bwd_workspace_floats(B, h, Lq),), dtype=torch.float32, device=q.device)
    with _Timed("attention_bwd", 10.0 * B * h * Lq * Lk * d):
        check(lib.mvit_attention_bwd(_ptr(q), _ptr(k), _ptr(v), _ptr(out), _ptr(dout), _ptr(lse), _ptr(dq), _ptr(dkv[0]),
                                     _ptr(dkv[1]), _ptr(ws), B, h, Lq, Lk, d, float(scale), 1 if add_q else 0, _dt(q),
                                     impl, _stream()), "mvit_attention_bwd")
    launch_count += 3 if ws is not None else 1
    return dq, dkv[0].to(q.dtype), dkv[1].to(q.dtype)


def attention_pool_bwd(what: int, x: Optional[torch.Tensor], strides: Tuple[int, int, int], dy: torch.Tensor,
                       weight: Optional[torch.Tensor], dx: Optional[torch.Tensor], dw: Optional[torch.Tensor], B: int,
                       heads: int, d: int, thw: Sequence[int], kernel: Sequence[int], stride: Sequence[int]):
    """Raw form of mvit_attention_pool_bwd; x / dx are addressed through `strides` = (batch, token, head)."""
    global launch_count
    _need_cuda(x, dy, weight, dx, dw)
    assert dy.is_contiguous()
    weight = _f32c(weight)
    with _Timed(("pool_bwd_dgrad", "pool_bwd_wgrad", "pool_bwd_max", "pool_bwd_max")[what] + "_s%d" % stride[1], 0.0):
        check(_lib.load().mvit_attention_pool_bwd(what, _ptr(x), strides[0], strides[1], strides[2], _ptr(dy), _ptr(weight),
                                                  _ptr(dx), _ptr(dw), B, heads, d, thw[0], thw[1], thw[2], *kernel, *stride,
                                                  _dt(dy), _stream()), "mvit_attention_pool_bwd")
    launch_count += 1
