"""CPU: the C-ABI library loads and exports what include/mvit_b200.h declares; the Python drop-in has
the reference's state_dict layout; the product refuses to compute without CUDA (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

from aicity_action_b200 import _lib
from aicity_action_b200.attention import MultiScaleAttention, MultiScaleBlock, attention_pool
from aicity_action_b200.config import AICITY_PRESETS, aicity_cfg
from aicity_action_b200.mvit import MViT, round_width
from tests.golden.cases import MODEL_CASES, tiny_cfg_overrides

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "mvit_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mvit_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 9
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/mvit_b200.h but missing from the .so"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert set(_lib.SIGNATURES) == set(names)
    assert lib.mvit_abi_version() == 1


def test_library_has_no_torch_dependency():
    # a C ABI: the .so must not link against libtorch / libc10 / libpython
    import subprocess
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "torch" not in out and "c10" not in out and "python" not in out, out


def test_bad_arguments_are_rejected_before_launch():
    lib = _lib.load()
    rc = lib.mvit_layernorm_fwd(None, None, None, None, 4, 96, 1e-6, 0, None)
    assert rc != 0 and b"null" in lib.mvit_last_error()
    rc = lib.mvit_attention_fwd(1, 1, 1, 1, None, 1, 1, 8, 8, 64, 0.1, 0, 0, 0, None)
    assert rc != 0 and b"head_dim" in lib.mvit_last_error()


@pytest.mark.parametrize("c", MODEL_CASES, ids=lambda c: c["name"])
def test_state_dict_layout_matches_reference(c, golden_index):
    torch.manual_seed(0)
    m = MViT(aicity_cfg(c["yaml"], tiny_cfg_overrides(c)))
    ref_shapes = golden_index["model_shapes"][c["name"]]
    mine = {k: list(v.shape) for k, v in m.state_dict().items()}
    assert list(mine.keys()) == list(ref_shapes.keys())      # names AND registration order
    assert mine == ref_shapes


def test_presets_cover_the_six_reference_yamls():
    assert len(AICITY_PRESETS) == 6
    n = {k: sum(p.numel() for p in MViT(aicity_cfg(k)).parameters()) for k in
         ("MVITV2_FULL_B_16x4_CONV", "MVITV2_FULL_B_16x4_CONV_448", "MVITV2_B_16x4_CONV", "MVITV2_FULL_B_32x3_CONV")}
    # SURVEY.md Appendix A / §6 parameter counts measured on the reference
    assert n["MVITV2_FULL_B_16x4_CONV"] == 34415538
    assert n["MVITV2_FULL_B_16x4_CONV_448"] == 35318706
    assert n["MVITV2_B_16x4_CONV"] == 34379346
    assert n["MVITV2_FULL_B_32x3_CONV"] == 51000018


def test_round_width():
    assert round_width(96, 2.0, divisor=2) == 192 and round_width(1, 2.0) == 2 and round_width(96, 0) == 96


def test_no_cpu_fallback():
    blk = MultiScaleBlock(96, 96, 1, kernel_q=[3, 3, 3], kernel_kv=[3, 3, 3], stride_q=[1, 1, 1],
                          stride_kv=[1, 2, 2], has_cls_embed=False)
    with pytest.raises(_lib.MvitLibraryError):
        with torch.no_grad():
            blk(torch.randn(1, 2 * 4 * 4, 96), [2, 4, 4])
    with pytest.raises(_lib.MvitLibraryError):
        attention_pool(torch.randn(1, 1, 32, 96), torch.nn.MaxPool3d([1, 3, 3], [1, 2, 2], [0, 1, 1]), [2, 4, 4],
                       has_cls_embed=False)


def test_module_attribute_parity():
    a = MultiScaleAttention(96, num_heads=2, qkv_bias=True, kernel_q=[3, 3, 3], kernel_kv=[3, 3, 3],
                            stride_q=[1, 2, 2], stride_kv=[1, 4, 4], has_cls_embed=False, expand_channel=True,
                            expand_to_dim=192)
    assert [n for n, _ in a.named_parameters()] == [
        "qkv.weight", "qkv.bias", "proj.weight", "proj.bias", "pool_q.weight", "norm_q.weight", "norm_q.bias",
        "pool_k.weight", "norm_k.weight", "norm_k.bias", "pool_v.weight", "norm_v.weight", "norm_v.bias"]
    assert a.pool_q.weight.shape == (96, 1, 3, 3, 3) and a.norm_q.eps == 1e-5 and abs(a.scale - 96 ** -0.5) < 1e-12
    nopool = MultiScaleAttention(96, num_heads=1, kernel_q=(1, 1, 1), stride_q=(1, 1, 1))
    assert nopool.pool_q is None and nopool.norm_q is None


def test_install_into_reference_checkout():
    """aicity_action_b200.patch.install(): the reference's own build_model(cfg) then returns the drop-in MViT with a
    state_dict that loads into / from the reference model.  Needs /root/reference (absent on the GPU box -> skipped)."""
    if not os.path.isdir("/root/reference/slowfast"):
        pytest.skip("reference checkout not present")
    import ref_shims
    ref_shims.install()
    cfg = ref_shims.ref_cfg("MVITV2_FULL_B_16x4_CONV.yaml", tiny_cfg_overrides(MODEL_CASES[0]))
    ref_model = ref_shims.ref_build_model(cfg, seed=0)
    import aicity_action_b200.patch as b200
    from aicity_action_b200.mvit import MViT
    import slowfast.models.attention as ref_attn
    import slowfast.models.build as ref_build
    saved = (ref_build.MODEL_REGISTRY._obj_map["MViT"], ref_attn.attention_pool, ref_attn.MultiScaleAttention,
             ref_attn.MultiScaleBlock)
    try:
        assert b200.install() == {"registry": True, "attention": True}
        ours = ref_shims.ref_build_model(cfg, seed=0)
        assert isinstance(ours, MViT)
        a, b = ref_model.state_dict(), ours.state_dict()
        assert list(a.keys()) == list(b.keys())
        for k in a:
            assert torch.equal(a[k], b[k]), k
        ours.load_state_dict(a, strict=True)
    finally:
        ref_build.MODEL_REGISTRY._obj_map["MViT"] = saved[0]
        ref_attn.attention_pool, ref_attn.MultiScaleAttention, ref_attn.MultiScaleBlock = saved[1:]


def test_graphed_forward_refuses_cpu_input():
    from aicity_action_b200.graphed import GraphedForward
    c = MODEL_CASES[0]
    m = MViT(aicity_cfg(c["yaml"], tiny_cfg_overrides(c))).eval()
    with pytest.raises(_lib.MvitLibraryError):
        GraphedForward(m, torch.zeros(1, 3, 8, 64, 64))


def test_patch_embed_fold_geometry_equals_conv3d():
    """Host logic of the implicit-GEMM patch embedding: space-to-depth fold of the clip + the Conv3d weight scattered into
    the folded layout (PatchEmbed._folded_weight) reproduce F.conv3d exactly (fp32 emulation of what the kernel sums)."""
    import torch.nn.functional as F
    from aicity_action_b200.mvit import PatchEmbed
    torch.manual_seed(0)
    pe = PatchEmbed(dim_in=3, dim_out=16, kernel=(3, 7, 7), stride=(2, 4, 4), padding=(1, 3, 3))
    x = torch.randn(2, 3, 8, 32, 32)
    ref = F.conv3d(x, pe.proj.weight, pe.proj.bias, stride=(2, 4, 4), padding=(1, 3, 3))     # [B, 16, 4, 8, 8]
    k, s, p, lo, taps, creal, cf = pe._fold_geometry()
    assert (taps, lo, creal, cf) == ([2, 2, 2], [-1, -1, -1], 96, 128)
    B, C, T, H, W = x.shape
    Tf, Hf, Wf = T // s[0], H // s[1], W // s[2]
    # folded[b, t, h, w, ((ot*sh + oh)*sw + ow)*C + c] = x[b, c, t*st + ot, h*sh + oh, w*sw + ow]
    f = x.view(B, C, Tf, s[0], Hf, s[1], Wf, s[2]).permute(0, 2, 4, 6, 3, 5, 7, 1).reshape(B, Tf, Hf, Wf, creal)
    f = F.pad(f, (0, cf - creal))
    wf = pe._folded_weight().float().view(16, taps[0], taps[1], taps[2], cf)     # bf16-rounded weights
    wq = pe.proj.weight.detach().bfloat16().float()
    ref_q = F.conv3d(x, wq, pe.proj.bias, stride=(2, 4, 4), padding=(1, 3, 3))
    out = torch.zeros(B, Tf, Hf, Wf, 16)
    fp = F.pad(f, (0, 0, 1, 1, 1, 1, 1, 1))            # one folded block of zero padding on every side
    for a in range(taps[0]):
        for b in range(taps[1]):
            for c in range(taps[2]):
                dt, dh, dw = lo[0] + a, lo[1] + b, lo[2] + c
                sl = fp[:, 1 + dt:1 + dt + Tf, 1 + dh:1 + dh + Hf, 1 + dw:1 + dw + Wf]
                out += torch.einsum("bthwc,nc->bthwn", sl, wf[:, a, b, c])
    out = out + pe.proj.bias.detach()
    assert torch.allclose(out.permute(0, 4, 1, 2, 3), ref_q, atol=1e-4)
    assert (out.permute(0, 4, 1, 2, 3) - ref).abs().max() < 5e-2        # only the bf16 rounding of the weights apart


def test_drop_path_scale_consumes_rng_like_the_reference():
    """common.py:46-59: mask = floor(keep + rand([B,1,1])); x / keep * mask — same draw, same values."""
    from aicity_action_b200.common import drop_path_scale
    B, p = 6, 0.3
    torch.manual_seed(123)
    s = drop_path_scale(B, p, True, torch.device("cpu"))
    torch.manual_seed(123)
    keep = 1 - p
    mask = keep + torch.rand((B, 1, 1), dtype=torch.float32)
    mask.floor_()
    x = torch.ones(B, 1, 1)
    ref = (x.div(keep) * mask).reshape(B)
    assert torch.equal(s, ref)
    assert drop_path_scale(B, 0.0, True, torch.device("cpu")) is None
    assert drop_path_scale(B, p, False, torch.device("cpu")) is None


def test_weight_cache_refreshes_in_place_and_can_be_invalidated():
    """version-counted updates refresh the cached bf16 operand IN PLACE (a captured graph reading it by address
    stays valid); `.data` updates bypass the version counter and need invalidate_caches()."""
    import torch
    from aicity_action_b200.weights import cached_weight, cached_weight_t, invalidate_caches
    lin = torch.nn.Linear(8, 4)
    a = cached_weight(lin.weight, torch.bfloat16)
    at = cached_weight_t(lin.weight, torch.bfloat16)
    ptr, ptr_t = a.data_ptr(), at.data_ptr()
    with torch.no_grad():
        lin.weight.mul_(2)
    b = cached_weight(lin.weight, torch.bfloat16)
    assert b.data_ptr() == ptr and torch.equal(b, lin.weight.detach().bfloat16())
    lin.weight.data.mul_(2)                                   # invisible to the version counter
    assert not torch.equal(cached_weight(lin.weight, torch.bfloat16), lin.weight.detach().bfloat16())
    assert invalidate_caches(lin) == 1
    c, ct = cached_weight(lin.weight, torch.bfloat16), cached_weight_t(lin.weight, torch.bfloat16)
    assert c.data_ptr() == ptr and torch.equal(c, lin.weight.detach().bfloat16())
    assert ct.data_ptr() == ptr_t and torch.equal(ct, lin.weight.detach().bfloat16().t())


def test_rel_pos_is_off_by_default_and_adds_upstream_named_tables_when_on():
    """The relative-position extension (not in the reference) must not touch the default state_dict."""
    from aicity_action_b200.config import aicity_cfg
    from aicity_action_b200.mvit import MViT
    cfg = aicity_cfg("MVITV2_FULL_B_16x4_CONV.yaml")
    assert not cfg.MVIT.REL_POS_SPATIAL and not cfg.MVIT.REL_POS_TEMPORAL
    base = MViT(cfg).state_dict()
    assert not any("rel_pos" in k for k in base)
    on = MViT(aicity_cfg("MVITV2_FULL_B_16x4_CONV.yaml", ["MVIT.REL_POS_SPATIAL", True, "MVIT.REL_POS_TEMPORAL", True]))
    extra = {k: tuple(v.shape) for k, v in on.state_dict().items() if k not in base}
    assert len(extra) == 3 * len(on.blocks) and set(k.rsplit(".", 1)[1] for k in extra) == {"rel_pos_h", "rel_pos_w", "rel_pos_t"}
    # block 0 @224: q grid 8x56x56, k/v grid 8x7x7 -> [2*max-1, head_dim]
    assert extra["blocks.0.attn.rel_pos_h"] == (111, 96) and extra["blocks.0.attn.rel_pos_t"] == (15, 96)
    assert [k for k in on.state_dict() if k in base] == list(base)       # the reference's keys, in the reference's order


def test_pdl_launched_kernels_wait_on_their_predecessor():
    """A kernel launched with programmatic stream serialisation that never executes griddepcontrol.wait would run
    concurrently with the producer of its inputs.  Every kernel family that `launch_pdl` / `pdl_attr` launches must carry the
    wait (SASS: ACQBULK) and the early-launch trigger (PREEXIT) in every instantiation, and no other source may use them."""
    import glob
    import re
    import shutil
    import subprocess
    lib = os.path.join(ROOT, "aicity_action_b200", "lib", "libmvit_b200.so")
    if shutil.which("cuobjdump") is None or not os.path.exists(lib):
        pytest.skip("cuobjdump or the built library is not available")
    users = sorted(os.path.basename(f) for f in glob.glob(os.path.join(ROOT, "aicity_action_b200", "csrc", "*.cu"))
                   if re.search(r"\blaunch_pdl\(|\bpdl_attr\(", open(f).read()))
    assert users == ["attention_tc.cu", "gemm_tc.cu", "mlp_fused.cu"], users
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    waits, fn = {}, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1)
            waits.setdefault(fn, [0, 0])
        elif fn and "ACQBULK" in line:
            waits[fn][0] += 1
        elif fn and "PREEXIT" in line:
            waits[fn][1] += 1
    families = ("attention_tc_kernel", "linear_tc_kernel", "mlp_fused_kernel")
    checked = [f for f in waits if any(k in f for k in families)]
    assert len(checked) >= 15
    for f in checked:
        assert waits[f][0] >= 1 and waits[f][1] >= 1, f
