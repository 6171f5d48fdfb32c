"""Case tables shared by `oracle/make_golden.py` (which runs the reference) and the tests.

Shapes follow the MViTv2 stage shapes of SURVEY.md Appendix A scaled down so the CPU
oracle finishes in seconds; head_dim is always 96 as in every shipped Aicity config.
"""
from __future__ import annotations

K3 = [3, 3, 3]


def _pool(name, thw, stride, *, B=2, heads=2, d=96, mode="conv", kernel=None, cls=False, ndim=4, seed=11):
    return dict(name="pool_" + name, thw=list(thw), stride=list(stride), B=B, heads=heads, d=d, mode=mode,
                kernel=list(kernel or K3), cls=cls, ndim=ndim, seed=seed)


POOL_CASES = [
    _pool("s111", (4, 8, 8), (1, 1, 1)),
    _pool("s122", (4, 8, 8), (1, 2, 2)),
    _pool("s144", (4, 16, 16), (1, 4, 4), heads=1),
    _pool("s188", (2, 16, 16), (1, 8, 8), heads=1),
    _pool("s222", (4, 8, 8), (2, 2, 2), B=1),
    _pool("odd_s111", (3, 7, 5), (1, 1, 1), B=1, heads=3),
    _pool("odd_s122", (3, 7, 5), (1, 2, 2), B=1, heads=3),
    _pool("cls_s122", (4, 8, 8), (1, 2, 2), cls=True),
    _pool("skip_max", (4, 8, 8), (1, 2, 2), mode="max", kernel=[1, 3, 3], ndim=3, d=192, heads=1),
    _pool("skip_max_odd", (3, 7, 5), (1, 2, 2), mode="max", kernel=[1, 3, 3], ndim=3, d=96, heads=1, B=1),
    _pool("skip_max_222", (4, 8, 8), (2, 2, 2), mode="max", kernel=[3, 3, 3], ndim=3, d=96, heads=1, B=1),
]


def _attn(name, dim, dim_out, heads, thw, sq, skv, *, B=2, cls=False, residual=True, kq=K3, kkv=K3, seed=21):
    return dict(name="attn_" + name, dim=dim, dim_out=dim_out, heads=heads, thw=list(thw), stride_q=list(sq),
                stride_kv=list(skv), kernel_q=list(kq), kernel_kv=list(kkv), B=B, cls=cls,
                residual=residual, seed=seed)


ATTN_CASES = [
    _attn("blk0", 96, 96, 1, (4, 16, 16), (1, 1, 1), (1, 8, 8)),          # stage-1 shape (Lk tiny)
    _attn("blk1_expand", 96, 192, 2, (4, 16, 16), (1, 2, 2), (1, 4, 4)),  # expand + q stride 2
    _attn("mid", 192, 192, 2, (4, 8, 8), (1, 1, 1), (1, 2, 2)),
    _attn("last", 384, 384, 4, (2, 6, 6), (1, 1, 1), (1, 1, 1), B=1),     # Lq=Lk=72 ragged vs 128 tiles
    _attn("nopoolq", 192, 192, 2, (4, 8, 8), [], (1, 2, 2), residual=False, kq=[]),   # non-FULL variant
]


def _blk(name, dim, dim_out, heads, thw, sq, skv, *, B=2, cls=False, residual=True, expand_front=True,
         kq=K3, kkv=K3, seed=31):
    return dict(name="block_" + name, dim=dim, dim_out=dim_out, heads=heads, thw=list(thw), stride_q=list(sq),
                stride_kv=list(skv), kernel_q=list(kq), kernel_kv=list(kkv), B=B, cls=cls,
                residual=residual, expand_front=expand_front, seed=seed)


BLOCK_CASES = [
    _blk("plain", 96, 96, 1, (4, 16, 16), (1, 1, 1), (1, 8, 8)),
    _blk("expand_down", 96, 192, 2, (4, 16, 16), (1, 2, 2), (1, 4, 4)),
    _blk("deep", 384, 384, 4, (2, 8, 8), (1, 1, 1), (1, 2, 2), B=1),
    _blk("nonfull_nopoolq", 192, 192, 2, (4, 8, 8), [], (1, 2, 2), residual=False, kq=[]),
    _blk("v1_expand_back", 96, 192, 1, (4, 8, 8), (1, 1, 1), (1, 2, 2), expand_front=False, B=1),
]

TINY = ["DATA.TRAIN_CROP_SIZE", 64, "DATA.TEST_CROP_SIZE", 64, "DATA.NUM_FRAMES", 8,
        "MVIT.DEPTH", 4, "MVIT.DIM_MUL", [[1, 2.0], [2, 2.0]], "MVIT.HEAD_MUL", [[1, 2.0], [2, 2.0]],
        "MVIT.POOL_Q_STRIDE", [[1, 1, 2, 2], [2, 1, 2, 2]], "MVIT.DROPPATH_RATE", 0.1]

MODEL_CASES = [
    dict(name="tiny_full", yaml="MVITV2_FULL_B_16x4_CONV.yaml", tiny=True, B=2, seed=41),
    dict(name="tiny_nonfull", yaml="MVITV2_B_16x4_CONV.yaml", tiny=True, B=1, seed=42),
    dict(name="s224_full", yaml="MVITV2_FULL_B_16x4_CONV.yaml", tiny=False, B=1, seed=43),
    dict(name="b448_full", yaml="MVITV2_FULL_B_16x4_CONV_448.yaml", tiny=False, B=1, seed=44),
]


def tiny_cfg_overrides(case):
    return list(TINY) if case.get("tiny") else []
