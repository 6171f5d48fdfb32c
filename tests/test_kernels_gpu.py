"""GPU parity of each C-ABI kernel against the CPU oracle on identical seeded inputs.

Tolerances (north_star): fp32 within 1e-4, bf16 within 2e-2, both as ‖Δ‖∞/‖ref‖∞."""
import math

import pytest
import torch
import torch.nn.functional as F

import mvit_oracle as O
from aicity_action_b200 import ops
from aicity_action_b200._lib import IMPL_AUTO, IMPL_SIMT, IMPL_TCGEN05
from tests.conftest import rel_inf
from tests.golden.cases import POOL_CASES
from tests.golden.synth import synth_input, synth_tensor
from tests.test_oracle_golden import pool_oracle

pytestmark = pytest.mark.gpu
TOL = {torch.float32: 1e-4, torch.bfloat16: 2e-2}
DTYPES = [torch.float32, torch.bfloat16]


def dev(t, dtype=None):
    t = t.cuda()
    return t.to(dtype) if dtype is not None else t


def rounded(t, dtype):
    """The oracle sees exactly the values the kernel sees (bf16-rounded inputs for the bf16 path)."""
    return t.to(dtype).float()


@pytest.mark.parametrize("dtype", DTYPES, ids=["f32", "bf16"])
@pytest.mark.parametrize("C", [96, 192, 384, 768, 80])
def test_layernorm(C, dtype):
    x = rounded(synth_input(1, f"ln{C}", (3, 37, C)) * 2 + 0.5, dtype)
    g, b = synth_tensor(1, "norm.weight", (C,)), synth_tensor(1, "norm.bias", (C,))
    ref = F.layer_norm(x, (C,), g, b, 1e-6)
    got = ops.layernorm(dev(x, dtype), dev(g), dev(b), 1e-6)
    assert got.dtype == dtype and rel_inf(got, ref) < TOL[dtype]


@pytest.mark.parametrize("dtype", DTYPES, ids=["f32", "bf16"])
@pytest.mark.parametrize("c", POOL_CASES, ids=lambda c: c["name"])
def test_attention_pool_vs_golden(c, dtype, golden):
    x, w, ln, _ = pool_oracle(c)
    xr = rounded(x, dtype)
    ref, thw = O.attention_pool(xr, c["thw"], mode=c["mode"], kernel=c["kernel"], stride=c["stride"], weight=w,
                                has_cls=c["cls"], ln=ln)
    xd = dev(xr, dtype)
    lnd = None if ln is None else (dev(ln[0]), dev(ln[1]), ln[2])
    if c["ndim"] == 4:
        got, gthw = ops.attention_pool_heads(xd, c["thw"], c["kernel"], c["stride"], mode=c["mode"],
                                             weight=None if w is None else dev(w), ln=lnd, has_cls=c["cls"])
    else:
        got, gthw = ops.attention_pool_tokens(xd, c["thw"], c["kernel"], c["stride"], mode=c["mode"],
                                              has_cls=c["cls"])
    assert gthw == thw == list(golden[c["name"] + ".thw"])
    assert rel_inf(got, ref) < TOL[dtype]
    if dtype == torch.float32:      # and directly against the reference-generated fixture
        assert rel_inf(got, torch.from_numpy(golden[c["name"]])) < TOL[dtype]


@pytest.mark.parametrize("dtype", DTYPES, ids=["f32", "bf16"])
def test_attention_pool_reads_qkv_layout_in_place(dtype):
    """q/k/v are pooled straight out of the [B, N, 3, h, d] GEMM output (strided view, no copy)."""
    B, h, d, thw = 2, 2, 96, (4, 8, 8)
    N = math.prod(thw)
    qkv = rounded(synth_input(5, "qkv", (B, N, 3, h, d)), dtype)
    w = synth_tensor(5, "pool_k.weight", (d, 1, 3, 3, 3))
    g, b = synth_tensor(5, "norm_k.weight", (d,)), synth_tensor(5, "norm_k.bias", (d,))
    qd = dev(qkv, dtype)
    for which, stride in ((0, [1, 1, 1]), (1, [1, 2, 2]), (2, [1, 4, 4])):
        view = qkv[:, :, which].permute(0, 2, 1, 3)
        ref, _ = O.attention_pool(view, thw, mode="conv", kernel=[3, 3, 3], stride=stride, weight=w, ln=(g, b, 1e-5))
        got, _ = ops.attention_pool_heads(qd[:, :, which].permute(0, 2, 1, 3), list(thw), [3, 3, 3], stride,
                                          mode="conv", weight=dev(w), ln=(dev(g), dev(b), 1e-5))
        assert rel_inf(got, ref) < TOL[dtype]


@pytest.mark.parametrize("save", [False, True], ids=["infer", "save_pre"])
@pytest.mark.parametrize("B,h,thw,strides", [
    (2, 2, (4, 8, 8), (1, 2, 2)),            # q stride 1, k/v stride 2 (blocks 4-13): two TMA launches
    (1, 3, (3, 7, 5), (1, 1, 1)),            # all stride 1 (block 15), odd grid: ragged tiles, one launch for q, k and v
    (2, 1, (2, 16, 16), (1, 8, 8)),          # block 0: q on the TMA kernel, k/v (stride 8) on the tiled kernel
    (1, 2, (4, 16, 16), (2, 4, 4)),          # block 1
    (2, 4, (8, 12, 10), (2, 2, 2)),          # block 3: all stride 2, one launch
    (1, 2, (9, 9, 17), (2, 1, 1)),           # block 14, T not a multiple of 3, odd sizes
    (3, 2, (1, 4, 4), (1, 1, 1)),            # a single frame
], ids=str)
def test_attention_pool_qkv_fused(B, h, thw, strides, save):
    """mvit_attention_pool_qkv_fwd (persistent TMA-fed kernel; q/k/v of one block in one call) against the oracle's
    attention_pool on each of q, k, v: conv + LayerNorm(1e-5), and the saved pre-LayerNorm conv output."""
    d = 96
    N = math.prod(thw)
    qkv = rounded(synth_input(9, f"qkv{B}{h}{thw}", (B, N, 3, h, d)), torch.bfloat16)
    ws = [synth_tensor(9, f"pool_{n}.weight", (d, 1, 3, 3, 3)) for n in "qkv"]
    lns = [(synth_tensor(9, f"norm_{n}.weight", (d,)), synth_tensor(9, f"norm_{n}.bias", (d,)), 1e-5) for n in "qkv"]
    st = [(1, strides[0], strides[0]), (1, strides[1], strides[1]), (1, strides[2], strides[2])]
    qd = dev(qkv, torch.bfloat16).reshape(B, N, 3 * h * d)
    assert ops.pool_qkv_supported(qd, h, st)
    outs, grids, pres = ops.attention_pool_qkv(qd, h, list(thw), [dev(w) for w in ws],
                                               [(dev(g), dev(b), e) for g, b, e in lns], st, save_pre=save)
    for i in range(3):
        view = qkv[:, :, i].permute(0, 2, 1, 3)
        ref, rthw = O.attention_pool(view, thw, mode="conv", kernel=[3, 3, 3], stride=list(st[i]), weight=ws[i], ln=lns[i])
        assert grids[i] == rthw
        assert rel_inf(outs[i], ref) < TOL[torch.bfloat16], (i, rel_inf(outs[i], ref))
        if save:
            pre, _ = O.attention_pool(view, thw, mode="conv", kernel=[3, 3, 3], stride=list(st[i]), weight=ws[i], ln=None)
            assert rel_inf(pres[i], pre) < TOL[torch.bfloat16]
        # and against the per-tensor kernel it replaces (same fp32 convolution order; the LayerNorm sums differ in order)
        old, _ = ops.attention_pool_heads(qd.view(B, N, 3, h, d)[:, :, i].permute(0, 2, 1, 3), list(thw), [3, 3, 3],
                                          list(st[i]), mode="conv", weight=dev(ws[i]), ln=(dev(lns[i][0]), dev(lns[i][1]), 1e-5))
        assert rel_inf(outs[i], old) < 8e-3, i


def _impls(dtype):
    return [IMPL_SIMT] if dtype == torch.float32 else [IMPL_SIMT, IMPL_AUTO]


@pytest.mark.parametrize("dtype", DTYPES, ids=["f32", "bf16"])
@pytest.mark.parametrize("M,N,K", [(200, 288, 96), (130, 96, 384), (257, 768, 192), (64, 3072, 768), (50, 18, 768),
                                   (1000, 192, 96), (384, 1152, 384)])
def test_linear(M, N, K, dtype):
    x = rounded(synth_input(2, f"x{M}{N}{K}", (M, K)), dtype)
    w = rounded(synth_input(2, f"w{M}{N}{K}", (N, K)) / K ** 0.5, dtype)
    bias = synth_input(2, "b", (N,)) * 0.1
    res = rounded(synth_input(2, "r", (M, N)), dtype)
    scale = torch.tensor([0.0, 1.25])
    for impl in _impls(dtype):
        for gelu, use_res, use_scale in [(False, False, False), (True, False, False), (False, True, False),
                                         (False, True, True)]:
            if use_scale and M % 2:
                continue
            ref = F.linear(x, w, bias)
            if gelu:
                ref = F.gelu(ref)
            if use_scale:
                ref = ref * scale.repeat_interleave(M // 2)[:, None]
            if use_res:
                ref = ref + res
            got = ops.linear(dev(x, dtype), dev(w, dtype), dev(bias), residual=dev(res, dtype) if use_res else None,
                             row_scale=dev(scale) if use_scale else None, gelu=gelu, impl=impl)
            assert rel_inf(got, ref) < TOL[dtype], (impl, gelu, use_res, use_scale)


@pytest.mark.parametrize("dtype", DTYPES, ids=["f32", "bf16"])
@pytest.mark.parametrize("B,h,Lq,Lk", [(2, 2, 256, 64), (1, 4, 72, 72), (1, 1, 1024, 16), (2, 1, 130, 392),
                                       (1, 2, 392, 1568), (1, 8, 33, 257)])
def test_attention(B, h, Lq, Lk, dtype):
    d = 96
    q = rounded(synth_input(3, f"q{Lq}", (B, h, Lq, d)), dtype)
    k = rounded(synth_input(3, f"k{Lk}", (B, h, Lk, d)), dtype)
    v = rounded(synth_input(3, f"v{Lk}", (B, h, Lk, d)), dtype)
    scale = d ** -0.5
    s = (q @ k.transpose(-2, -1)) * scale
    ref_lse = torch.logsumexp(s, dim=-1)
    o = s.softmax(-1) @ v
    for impl in _impls(dtype):
        for add_q in (True, False):
            ref = (o + q if add_q else o).transpose(1, 2).reshape(B, Lq, h * d)
            got, lse = ops.attention(dev(q, dtype), dev(k, dtype), dev(v, dtype), scale, add_q, want_lse=True,
                                     impl=impl)
            assert rel_inf(got, ref) < TOL[dtype], (impl, add_q)
            assert rel_inf(lse, ref_lse) < 1e-3, (impl, add_q)


@pytest.mark.parametrize("B,T,H,W,C", [(2, 4, 14, 14, 192), (1, 2, 7, 9, 96), (3, 1, 28, 28, 768), (1, 3, 5, 1, 32)])
def test_skip_maxpool_is_exact(B, T, H, W, C):
    """MaxPool3d([1,3,3],[1,2,2],[0,1,1]) of the skip path (attention.py:427-432) on channels-last tokens: a maximum has no
    rounding, so the specialised kernel (clamped taps instead of -inf padding) must equal torch bit for bit, odd sizes and
    single-column images included."""
    g = torch.Generator().manual_seed(5)
    x = torch.randn(B, T * H * W, C, generator=g).bfloat16().cuda()
    got, thw = ops.attention_pool_tokens(x, [T, H, W], [1, 3, 3], [1, 2, 2], mode="max")
    ref = F.max_pool3d(x.view(B, T, H, W, C).permute(0, 4, 1, 2, 3).float(), (1, 3, 3), (1, 2, 2), (0, 1, 1))
    assert thw == list(ref.shape[2:])
    assert torch.equal(got, ref.permute(0, 2, 3, 4, 1).reshape(B, -1, C).bfloat16())


@pytest.mark.parametrize("sigma", [2.0, 3.0, 4.0])
@pytest.mark.parametrize("B,h,Lq,Lk", [(2, 4, 1568, 1568), (1, 2, 6272, 1568), (2, 1, 300, 257)])
def test_attention_large_score_spread(B, h, Lq, Lk, sigma):
    """Scores whose row maximum grows by more than 2^8 from one key tile to a later one take the lazy-rescale path of the
    softmax warps (O and the denominator are rescaled in tensor memory behind the P.V MMA).  Unit-variance inputs never
    reach it; trained weights and the stage-transition blocks do.  Regression for the o_done parity aliasing: checked
    against fp32 AND for run-to-run bit equality (the failure was a race)."""
    d = 96
    g = torch.Generator().manual_seed(11)
    q = (torch.randn(B, h, Lq, d, generator=g) * sigma).bfloat16()
    k = (torch.randn(B, h, Lk, d, generator=g) * sigma).bfloat16()
    v = torch.randn(B, h, Lk, d, generator=g).bfloat16()
    scale = d ** -0.5
    qd, kd, vd = dev(q, torch.bfloat16), dev(k, torch.bfloat16), dev(v, torch.bfloat16)
    s = (qd.float() @ kd.float().transpose(-2, -1)) * scale
    ref_lse = torch.logsumexp(s, dim=-1)
    ref = (s.softmax(-1) @ vd.float() + qd.float()).transpose(1, 2).reshape(B, Lq, h * d)
    first, lse = ops.attention(qd, kd, vd, scale, True, want_lse=True)
    first = first.clone()
    assert rel_inf(first, ref) < TOL[torch.bfloat16]
    assert rel_inf(lse, ref_lse) < 1e-3
    for _ in range(10):
        assert torch.equal(ops.attention(qd, kd, vd, scale, True), first)


@pytest.mark.parametrize("dtype", DTYPES, ids=["f32", "bf16"])
def test_pos_embed_and_head(dtype):
    B, T, HW, C = 2, 4, 36, 96
    tok = rounded(synth_input(4, "tok", (B, T * HW, C)), dtype)
    ps, pt = synth_input(4, "ps", (1, HW, C)), synth_input(4, "pt", (1, T, C))
    ref = tok + ps.repeat(1, T, 1) + torch.repeat_interleave(pt, HW, dim=1)
    got = ops.pos_embed_add(dev(tok, dtype), dev(ps), dev(pt), T, dtype)
    assert rel_inf(got, ref) < TOL[dtype]
    x = rounded(synth_input(4, "feat", (B, 50, 768)), dtype)
    w, b = synth_input(4, "hw", (18, 768)) * 0.05, synth_input(4, "hb", (18,))
    logits = F.linear(x.mean(1), w, b)
    assert rel_inf(ops.mean_head(dev(x, dtype), dev(w), dev(b), softmax=True), logits.softmax(1)) < TOL[dtype]
    assert rel_inf(ops.mean_head(dev(x, dtype), dev(w), dev(b), softmax=False), logits) < TOL[dtype]


def test_unsupported_requests_fail_loudly():
    from aicity_action_b200._lib import MvitLibraryError
    x = torch.randn(4, 96, device="cuda", dtype=torch.float16)
    with pytest.raises(TypeError):
        ops.layernorm(x, torch.ones(96, device="cuda"), torch.zeros(96, device="cuda"), 1e-6)
    q = torch.randn(1, 1, 8, 64, device="cuda")
    with pytest.raises(MvitLibraryError):
        ops.attention(q, q, q, 0.1, False)


def test_preprocess_u8_bit_exact_fp32():
    """module_wrapper.py:332-346 in NumPy float32 == the device kernel, bit for bit."""
    import numpy as np
    g = torch.Generator().manual_seed(9)
    fr = torch.randint(0, 256, (2, 4, 12, 10, 3), dtype=torch.uint8, generator=g)
    x = fr.numpy().astype("float32")
    x /= 255.0
    x = x.transpose([0, 4, 1, 2, 3])
    ref = (x - np.float32(0.45)) / np.float32(0.225)
    got = ops.preprocess_u8(fr.cuda(), torch.float32).cpu().numpy()
    assert np.array_equal(got, np.ascontiguousarray(ref))
    got16 = ops.preprocess_u8(fr.cuda(), torch.bfloat16).float().cpu()
    assert torch.equal(got16, torch.from_numpy(np.ascontiguousarray(ref)).bfloat16().float())


@pytest.mark.parametrize("dtype", DTYPES, ids=["f32", "bf16"])
def test_patch_embed_im2col_gemm(dtype):
    """im2col + GEMM (+bias, + positional table broadcast over the batch) == Conv3d + pos-embed add."""
    B, T, S = 2, 8, 32
    x = rounded(synth_input(6, "clip", (B, 3, T, S, S)), dtype)
    w = rounded(synth_input(6, "pw", (96, 3, 3, 7, 7)) / 21.0, dtype)
    bias = synth_input(6, "pb", (96,)) * 0.1
    ref = F.conv3d(x, w, bias, stride=(2, 4, 4), padding=(1, 3, 3)).flatten(2).transpose(1, 2)   # [B, N, 96]
    N = ref.shape[1]
    pos = rounded(synth_input(6, "pos", (N, 96)) * 0.1, dtype)
    ref = ref + pos
    patches, thw = ops.im2col3d(dev(x, dtype), [3, 7, 7], [2, 4, 4], [1, 3, 3], 448)
    assert thw == [4, 8, 8] and patches.shape == (B * N, 448)
    wp = torch.zeros(96, 448)
    wp[:, :441] = w.reshape(96, 441)
    for impl in _impls(dtype):
        got = ops.linear(patches, dev(wp, dtype), dev(bias), residual=dev(pos, dtype), residual_row_period=N, impl=impl)
        assert rel_inf(got.view(B, N, 96), ref) < TOL[dtype], impl


def test_patch_embed_implicit_gemm_conv():
    """fold + tcgen05 implicit-GEMM convolution (5-D TMA tap boxes) == Conv3d + pos-embed; uint8 frames == clip."""
    from aicity_action_b200.mvit import PatchEmbed
    B, T, S = 2, 8, 64
    pe = PatchEmbed(3, 96, (3, 7, 7), (2, 4, 4), (1, 3, 3))
    with torch.no_grad():
        pe.proj.weight.copy_(synth_input(7, "pw", (96, 3, 3, 7, 7)) / 21.0)
        pe.proj.bias.copy_(synth_input(7, "pb", (96,)) * 0.1)
    pe = pe.cuda()
    g = torch.Generator().manual_seed(11)
    frames = torch.randint(0, 256, (B, T, S, S, 3), dtype=torch.uint8, generator=g)
    clip = ((frames.float() / 255.0 - 0.45) / 0.225).permute(0, 4, 1, 2, 3).contiguous()
    xr = rounded(clip, torch.bfloat16)
    wr = rounded(pe.proj.weight.detach().cpu(), torch.bfloat16)
    ref = F.conv3d(xr, wr, pe.proj.bias.detach().cpu(), stride=(2, 4, 4), padding=(1, 3, 3)).flatten(2).transpose(1, 2)
    N = ref.shape[1]
    pos = rounded(synth_input(7, "pos", (N, 96)) * 0.1, torch.bfloat16)
    with torch.no_grad():
        got_clip = pe(xr.cuda().bfloat16(), torch.bfloat16, pos=pos.cuda().bfloat16(), pos_period=N)
        got_u8 = pe(frames.cuda(), torch.bfloat16, pos=pos.cuda().bfloat16(), pos_period=N)
        got_nopos = pe(xr.cuda().bfloat16(), torch.bfloat16)
    assert got_clip.shape == (B, N, 96)
    assert rel_inf(got_clip, ref + pos) < TOL[torch.bfloat16]
    assert rel_inf(got_u8, ref + pos) < TOL[torch.bfloat16]
    assert rel_inf(got_nopos, ref) < TOL[torch.bfloat16]


# ------------------------------------------------------------------------------------------------ folded LayerNorm
LN_FOLD_SHAPES = [
    # (M, K, N): block-stream widths of MViTv2-B and the three GEMM tile configurations (128x96, 128x128, CTA-pair 256x192)
    (300, 96, 288), (1000, 96, 384), (777, 192, 576), (20000, 384, 1152), (40000, 384, 1536), (19000, 768, 2304),
    (19000, 384, 1024), (50, 80, 40),
]


@pytest.mark.parametrize("M,K,N", LN_FOLD_SHAPES, ids=str)
@pytest.mark.parametrize("gelu", [False, True], ids=["plain", "gelu"])
def test_linear_ln_folded(M, K, N, gelu):
    """Producer + consumer of the folded LayerNorm (mvit_linear_ln_fwd) against the oracle's LayerNorm -> Linear (-> GELU):
    the stream x comes out of a stats-emitting GEMM (identity-ish weight + residual), so both halves are exercised."""
    from aicity_action_b200.weights import folded_ln_linear
    dt = torch.bfloat16
    base = rounded(synth_input(21, f"lnf{M}{K}", (M, K)) * 1.5 + 0.7, dt)          # non-zero mean rows
    eye = torch.eye(K) * 0.5
    res = rounded(synth_input(22, f"lnr{M}{K}", (M, K)), dt)
    x_dev = ops.linear_stats(dev(base, dt), dev(eye, dt), None, residual=dev(res, dt))
    x = x_dev.float().cpu()                                                        # exactly the bf16 values the consumer reads
    assert rel_inf(x, base @ eye.t() + res) < TOL[dt]
    stats = ops.row_stats_of(x_dev)
    assert stats is not None and stats.shape[1:] == (M, 2)
    tot = stats.sum(0).cpu()
    assert rel_inf(tot[:, 0], x.sum(1)) < 1e-4 and rel_inf(tot[:, 1], (x * x).sum(1)) < 1e-4
    w = torch.nn.Parameter(synth_tensor(23, f"w{N}{K}", (N, K)) * K ** -0.5)
    b = torch.nn.Parameter(synth_tensor(23, f"b{N}", (N,)) * 0.1)
    g = torch.nn.Parameter(1.0 + 0.3 * synth_tensor(23, f"g{K}", (K,)))
    be = torch.nn.Parameter(0.2 * synth_tensor(23, f"be{K}", (K,)))
    ref = F.linear(F.layer_norm(x, (K,), g, be, 1e-6), w, b)
    if gelu:
        ref = F.gelu(ref)
    wf, bf, cs = folded_ln_linear(w.cuda(), b.cuda(), g.cuda(), be.cuda())
    got = ops.linear_ln(x_dev, stats, wf, bf, cs, 1e-6, gelu=gelu)
    assert rel_inf(got, ref.detach()) < TOL[dt], rel_inf(got, ref.detach())
    # and against the unfolded kernels on the same device values
    xn = ops.layernorm(x_dev, g.detach().cuda(), be.detach().cuda(), 1e-6)
    unf = ops.linear(xn, w.detach().cuda().to(dt), b.detach().cuda(), gelu=gelu)
    assert rel_inf(got, unf) < TOL[dt]


@pytest.mark.parametrize("offset,outlier", [(8.0, 1.0), (30.0, 1.0), (2.0, 40.0)], ids=["dc8", "dc30", "outlier40"])
@pytest.mark.parametrize("M,K,N", [(3000, 96, 384), (3000, 384, 1152), (2000, 768, 2304)], ids=str)
def test_linear_ln_folded_hard_rows(M, K, N, offset, outlier):
    """The folded LayerNorm takes its variance in one pass (sum, sum of squares in fp32) and subtracts mean * colsum(W') after
    the GEMM: both cancel when |mean| >> std.  Rows with a DC offset of 8 and 30 standard deviations and rows with a few
    channels 40x larger than the rest (the 'massive activation' pattern of trained ViTs) must still meet the bf16 tolerance
    against LayerNorm -> Linear on the same bf16 rows.  (The fp32 one-pass variance loses ~ (mean/std)^2 * 1e-7 of relative
    accuracy: 1e-4 at 30 sigma; a stream with |mean|/std >~ 300 should run with MVIT_B200_LN_FOLD=0.)"""
    from aicity_action_b200.weights import folded_ln_linear
    dt = torch.bfloat16
    g0 = torch.Generator().manual_seed(31)
    base = torch.randn(M, K, generator=g0)
    base[:, ::37] *= outlier
    base = base + offset * torch.randn(M, 1, generator=g0).sign()
    base = rounded(base, dt)
    x_dev = ops.linear_stats(dev(base, dt), dev(torch.eye(K), dt), None)
    x = x_dev.float().cpu()
    stats = ops.row_stats_of(x_dev)
    w = torch.nn.Parameter(torch.randn(N, K, generator=g0) * K ** -0.5)
    b = torch.nn.Parameter(torch.randn(N, generator=g0) * 0.1)
    g = torch.nn.Parameter(1.0 + 0.3 * torch.randn(K, generator=g0))
    be = torch.nn.Parameter(0.2 * torch.randn(K, generator=g0))
    ref = F.linear(F.layer_norm(x.double(), (K,), g.double(), be.double(), 1e-6), w.double(), b.double()).float()
    wf, bf, cs = folded_ln_linear(w.cuda(), b.cuda(), g.cuda(), be.cuda())
    got = ops.linear_ln(x_dev, stats, wf, bf, cs, 1e-6)
    assert rel_inf(got, ref.detach()) < TOL[dt], rel_inf(got, ref.detach())


def test_patch_conv_row_stats():
    """mvit_patch_conv_stats_fwd: tokens identical to mvit_patch_conv_fwd, statistics indexed by token (not by tile row)."""
    from aicity_action_b200.mvit import PatchEmbed
    pe = PatchEmbed(3, 96, kernel=(3, 7, 7), stride=(2, 4, 4), padding=(1, 3, 3)).cuda()
    clip = dev(synth_input(5, "clip", (2, 3, 8, 64, 96)), torch.bfloat16)
    pos = dev(synth_input(5, "pos", (4 * 16 * 24, 96)), torch.bfloat16)
    a = pe(clip, torch.bfloat16, pos=pos, pos_period=pos.shape[0])
    b = pe(clip, torch.bfloat16, pos=pos, pos_period=pos.shape[0], want_stats=True)
    assert torch.equal(a, b)
    st = ops.row_stats_of(b)
    assert st is not None
    tot = st.sum(0)
    bf = b.float().reshape(-1, 96)
    assert rel_inf(tot[:, 0], bf.sum(1)) < 1e-4 and rel_inf(tot[:, 1], (bf * bf).sum(1)) < 1e-4


# ------------------------------------------------------------------------------------------------ relative-position bias
# DEFAULT-OFF extension that the reference does not have: parity here is against the in-repo restatement
# oracle/mvit_oracle.py::rel_pos_bias of upstream PySlowFast's formula (Appendix F) — unpinned.
REL_CASES = [
    # (B, heads, q_thw, k_thw)
    (1, 2, (2, 8, 8), (2, 8, 8)),            # equal grids
    (2, 1, (4, 8, 12), (4, 2, 3)),           # K/V pooled 4x (q finer than k)
    (1, 2, (2, 4, 4), (2, 8, 8)),            # q pooled (q coarser than k)
    (1, 1, (8, 14, 14), (8, 14, 14)),        # a stage-4 @448 shape: 36 bias columns, ragged tiles
    (1, 2, (8, 28, 28), (8, 28, 28)),        # 64 bias columns (the maximum)
]


@pytest.mark.parametrize("impl,dtype", [(IMPL_SIMT, torch.float32), (IMPL_SIMT, torch.bfloat16),
                                        (IMPL_TCGEN05, torch.bfloat16)], ids=["simt-f32", "simt-bf16", "tc-bf16"])
@pytest.mark.parametrize("B,h,q_thw,k_thw", REL_CASES, ids=str)
def test_attention_rel_pos(B, h, q_thw, k_thw, impl, dtype):
    d = 96
    Lq, Lk = math.prod(q_thw), math.prod(k_thw)
    q = rounded(synth_input(31, f"rq{q_thw}{k_thw}", (B, h, Lq, d)), dtype)
    k = rounded(synth_input(32, f"rk{q_thw}{k_thw}", (B, h, Lk, d)), dtype)
    v = rounded(synth_input(33, f"rv{q_thw}{k_thw}", (B, h, Lk, d)), dtype)
    tabs = [synth_tensor(34, f"rel{i}", (2 * max(a, b) - 1, d)) * 0.5
            for i, (a, b) in enumerate(((q_thw[1], k_thw[1]), (q_thw[2], k_thw[2]), (q_thw[0], k_thw[0])))]
    scale = d ** -0.5
    bias = O.rel_pos_bias(q, q_thw, k_thw, tabs[0], tabs[1], tabs[2])
    attn = ((q @ k.transpose(-2, -1)) * scale + bias).softmax(-1)
    ref = (attn @ v + q).transpose(1, 2).reshape(B, Lq, h * d)
    qd = dev(q, dtype)
    rel = ops.relpos_operands(qd, q_thw, k_thw, dev(tabs[0]), dev(tabs[1]), dev(tabs[2]), scale)
    # the operands themselves: q_ext . k_ext^T * scale == bias
    got_bias = (rel[0].float() @ rel[1].float().transpose(-2, -1)).cpu() * scale
    assert rel_inf(got_bias, bias) < TOL[dtype]
    got = ops.attention(qd, dev(k, dtype), dev(v, dtype), scale, True, impl=impl, rel=rel)
    assert rel_inf(got, ref) < TOL[dtype], rel_inf(got, ref)
    # and the bias matters in this test (guards against a silently ignored operand)
    plain = ops.attention(qd, dev(k, dtype), dev(v, dtype), scale, True, impl=impl)
    assert rel_inf(plain, ref) > 2 * TOL[dtype]


def test_gelu_epilogue_pointwise_accuracy():
    """The tcgen05 epilogue's GELU (tanh of a fitted odd polynomial, MUFU) against the exact erf GELU of common.py:20, point by
    point on [-9, 9]: within the bf16 rounding of the result plus 1e-3 (its documented bound is 3e-4 + 2.5e-4 |x|)."""
    K = 64
    x = torch.linspace(-9, 9, 512 * K).reshape(512, K).to(torch.bfloat16)
    eye = torch.eye(K, dtype=torch.bfloat16)
    got = ops.linear(x.cuda(), eye.cuda(), None, gelu=True, impl=IMPL_TCGEN05).float().cpu()
    ref = F.gelu(x.float())
    err = (got - ref).abs()
    bound = ref.abs() * 2.0 ** -8 + 1e-3
    assert bool((err <= bound).all()), float((err - bound).max())
    assert float(err.max()) < 2e-2 * float(ref.abs().max())


@pytest.mark.parametrize("M,C", [(128, 96), (300, 96), (5000, 96), (40000, 96), (777, 192), (30000, 192)], ids=str)
def test_mlp_fused(M, C):
    """mvit_mlp_fused_fwd: y = x + fc2(GELU(fc1(LN(x)))) in one kernel, against the oracle's LayerNorm -> Linear -> GELU ->
    Linear -> + x on the same bf16 stream, and against the two-GEMM path it replaces; row statistics of the result."""
    from aicity_action_b200.weights import folded_ln_linear
    dt = torch.bfloat16
    H = 4 * C
    base = rounded(synth_input(41, f"mf{M}{C}", (M, C)) * 1.5 + 0.4, dt)
    res = rounded(synth_input(42, f"mr{M}{C}", (M, C)), dt)
    x_dev = ops.linear_stats(dev(base, dt), dev(torch.eye(C) * 0.5, dt), None, residual=dev(res, dt))
    x = x_dev.float().cpu()
    stats = ops.row_stats_of(x_dev)
    w1 = torch.nn.Parameter(synth_tensor(43, f"w1{C}", (H, C)) * C ** -0.5)
    b1 = torch.nn.Parameter(synth_tensor(43, f"b1{C}", (H,)) * 0.1)
    w2 = torch.nn.Parameter(synth_tensor(43, f"w2{C}", (C, H)) * H ** -0.5)
    b2 = torch.nn.Parameter(synth_tensor(43, f"b2{C}", (C,)) * 0.1)
    g = torch.nn.Parameter(1.0 + 0.3 * synth_tensor(43, f"g{C}", (C,)))
    be = torch.nn.Parameter(0.2 * synth_tensor(43, f"be{C}", (C,)))
    ref = (x + F.linear(F.gelu(F.linear(F.layer_norm(x, (C,), g, be, 1e-6), w1, b1)), w2, b2)).detach()
    wf, bf, cs = folded_ln_linear(w1.cuda(), b1.cuda(), g.cuda(), be.cuda())
    w2d = w2.detach().cuda().to(dt)
    assert ops.mlp_fused_supported(C, H, C)
    got = ops.mlp_fused(x_dev, stats, wf, bf, cs, w2d, b2.detach().cuda(), 1e-6)
    assert rel_inf(got, ref) < TOL[dt], rel_inf(got, ref)
    h = ops.linear_ln(x_dev, stats, wf, bf, cs, 1e-6, gelu=True)
    two = ops.linear(h, w2d, b2.detach().cuda(), residual=x_dev)
    assert rel_inf(got, two) < TOL[dt]
    st = ops.row_stats_of(got)
    gf = got.float()
    assert st is not None and rel_inf(st[0, :, 0], gf.sum(1)) < 1e-4 and rel_inf(st[0, :, 1], (gf * gf).sum(1)) < 1e-4
