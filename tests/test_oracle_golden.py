"""CPU: the oracle restatement reproduces the fixtures generated from the UNMODIFIED reference
(oracle/make_golden.py).  This is the parity pin of the oracle itself."""
import numpy as np
import pytest
import torch

import mvit_oracle as O
import window_oracle as WO
from tests.conftest import rel_inf
from tests.golden.cases import ATTN_CASES, BLOCK_CASES, MODEL_CASES, POOL_CASES, tiny_cfg_overrides
from tests.golden.synth import synth_clip, synth_input, synth_state_dict, synth_tensor

TOL = 2e-5


def pool_oracle(c):
    d = c["d"]
    L = c["thw"][0] * c["thw"][1] * c["thw"][2] + (1 if c["cls"] else 0)
    shape = (c["B"], c["heads"], L, d) if c["ndim"] == 4 else (c["B"], L, d)
    x = synth_input(c["seed"], c["name"], shape)
    w = ln = None
    if c["mode"] == "conv":
        w = synth_tensor(c["seed"], "pool_q.weight", (d, 1, *c["kernel"]))
        ln = (synth_tensor(c["seed"], "norm_q.weight", (d,)), synth_tensor(c["seed"], "norm_q.bias", (d,)), 1e-5)
    return x, w, ln, O.attention_pool(x, c["thw"], mode=c["mode"], kernel=c["kernel"], stride=c["stride"],
                                      weight=w, has_cls=c["cls"], ln=ln)


@pytest.mark.parametrize("c", POOL_CASES, ids=lambda c: c["name"])
def test_pool(c, golden):
    _, _, _, (out, thw) = pool_oracle(c)
    assert list(golden[c["name"] + ".thw"]) == thw
    assert rel_inf(out, torch.from_numpy(golden[c["name"]])) < TOL


def attn_shapes(c):
    C, Ci, d = c["dim_out"], c["dim"], 96
    s = {"qkv.weight": (3 * C, Ci), "qkv.bias": (3 * C,), "proj.weight": (C, C), "proj.bias": (C,)}
    for n, k in (("q", c["kernel_q"]), ("k", c["kernel_kv"]), ("v", c["kernel_kv"])):
        if k:
            s[f"pool_{n}.weight"] = (d, 1, *k)
            s[f"norm_{n}.weight"] = (d,)
            s[f"norm_{n}.bias"] = (d,)
    return s


def attn_spec(c, expand):
    spec = O.BlockSpec(c["dim"], c["dim_out"], c["heads"], c["kernel_q"], c["kernel_kv"], c["stride_q"],
                       c["stride_kv"], 0.0, expand=expand)
    mv = O.MViTSpec([], [], 0, [], [], [], c["cls"], True, "conv", c["residual"], 0, True)
    return spec, mv


@pytest.mark.parametrize("c", ATTN_CASES, ids=lambda c: c["name"])
def test_attention(c, golden):
    sd = synth_state_dict(attn_shapes(c), c["seed"])
    N = c["thw"][0] * c["thw"][1] * c["thw"][2] + (1 if c["cls"] else 0)
    x = synth_input(c["seed"], c["name"], (c["B"], N, c["dim"]))
    spec, mv = attn_spec(c, c["dim_out"] != c["dim"])
    out, _ = O.multiscale_attention(x, c["thw"], sd, "", spec, mv)
    assert rel_inf(out, torch.from_numpy(golden[c["name"]])) < TOL


def block_shapes(c):
    expand = c["expand_front"] and c["dim"] != c["dim_out"]
    ca = c["dim_out"] if expand else c["dim"]
    s = {"norm1.weight": (c["dim"],), "norm1.bias": (c["dim"],)}
    a = dict(c, dim_out=ca)
    s.update({"attn." + k: v for k, v in attn_shapes(a).items()})
    s.update({"norm2.weight": (ca,), "norm2.bias": (ca,), "mlp.fc1.weight": (4 * ca, ca), "mlp.fc1.bias": (4 * ca,),
              "mlp.fc2.weight": (c["dim_out"], 4 * ca), "mlp.fc2.bias": (c["dim_out"],)})
    if ca != c["dim_out"]:
        s.update({"proj.weight": (c["dim_out"], ca), "proj.bias": (c["dim_out"],)})
    if expand:
        s.update({"proj_max_pool.weight": (c["dim_out"], c["dim"]), "proj_max_pool.bias": (c["dim_out"],)})
    return s, expand


@pytest.mark.parametrize("c", BLOCK_CASES, ids=lambda c: c["name"])
def test_block(c, golden):
    shapes, expand = block_shapes(c)
    sd = synth_state_dict(shapes, c["seed"])
    N = c["thw"][0] * c["thw"][1] * c["thw"][2] + (1 if c["cls"] else 0)
    x = synth_input(c["seed"], c["name"], (c["B"], N, c["dim"]))
    spec, mv = attn_spec(c, expand)
    out, _ = O.multiscale_block(x, c["thw"], sd, "", spec, mv)
    assert rel_inf(out, torch.from_numpy(golden[c["name"]])) < TOL


@pytest.mark.parametrize("c", MODEL_CASES, ids=lambda c: c["name"])
def test_model(c, golden, golden_index):
    from aicity_action_b200.config import aicity_cfg   # cfg presets only (host logic, no compute)
    cfg = aicity_cfg(c["yaml"], tiny_cfg_overrides(c))
    shapes = golden_index["model_shapes"][c["name"]]
    sd = synth_state_dict(shapes, c["seed"])
    x = synth_clip(c["seed"], c["B"], cfg.DATA.NUM_FRAMES, cfg.DATA.TRAIN_CROP_SIZE)
    spec = O.derive_spec(cfg)
    dims = np.asarray([[b.attn_dim, b.dim_out, b.heads] for b in spec.blocks])
    assert (dims == golden[c["name"] + ".blockdims"]).all()
    with torch.no_grad():
        probs, extra = O.mvit_forward(x, sd, spec, return_features=True)
    assert rel_inf(probs, torch.from_numpy(golden[c["name"] + ".probs"])) < TOL
    assert rel_inf(extra["feat"], torch.from_numpy(golden[c["name"] + ".feat"])) < 1e-4


def test_windows(golden, golden_index):
    for key, n in golden_index["windows"].items():
        _, nf, length, stride = key.split("_")
        w = WO.window_list(int(nf), int(length), int(stride))
        assert len(w) == n and np.array_equal(np.asarray(w), golden[key])
        idx = [WO.frame_indices(t0, t1, 16, int(nf)) for t0, t1 in w]
        assert np.array_equal(np.asarray(idx), golden[key + ".idx"])
    # SURVEY.md Appendix C known answers
    assert len(WO.window_list(18000)) == 1125
    assert WO.frame_indices(0, 64, 16, 18000) == [0, 4, 8, 12, 17, 21, 25, 29, 34, 38, 42, 46, 51, 55, 59, 64]
    assert WO.frame_indices(17984, 18048, 16, 18000) == [17984, 17988, 17992, 17996] + [17999] * 12
    assert WO.fps_adjust(64, 16, 25.0, 30.0) == (53, 13) and WO.fps_adjust(64, 16, 29.97, 30.0) == (64, 16)


def test_chunks_and_aggregate(golden, golden_index):
    for ent in golden_index["chunks"]:
        got = [(a, b, n, float(m)) for a, b, n, m in WO.get_chunks(np.asarray(ent["scores"], np.float32), ent["thr"])]
        assert got == [tuple(c) for c in ent["chunks"]]
    wl = WO.window_list(300, 64, 16)
    preds = [(t0, t1, golden["agg_in"][i]) for i, (t0, t1) in enumerate(wl)]
    assert np.array_equal(WO.aggregate(preds, 18, "mean"), golden["agg_mean"])
    assert np.array_equal(WO.aggregate(preds, 18, "max"), golden["agg_max"])
