#!/usr/bin/env python3
"""Pin the CPU oracle against the UNMODIFIED reference and write tests/golden/.

Run here (the container that has /root/reference):   python oracle/make_golden.py

For every case below the reference's own module (`slowfast.models.attention.*`,
`slowfast.models.build_model`, `scripts/module_wrapper.py`, `scripts/aicity_inf_graph.py`)
is executed on deterministic synthetic inputs/weights (`tests/golden/synth.py`), the
oracle restatement (`oracle/mvit_oracle.py`, `oracle/window_oracle.py`) is asserted to
agree with it, and the reference OUTPUT is stored.  Inputs and weights are not stored:
they are regenerated from (seed, name, shape).  TEST INFRASTRUCTURE — not product code.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import ref_shims  # noqa: E402
import mvit_oracle as O  # noqa: E402
import window_oracle as WO  # noqa: E402
from tests.golden.synth import synth_state_dict, synth_input, synth_clip  # noqa: E402
from tests.golden.cases import POOL_CASES, ATTN_CASES, BLOCK_CASES, MODEL_CASES, tiny_cfg_overrides  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
TOL = 2e-5


def rel_err(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def load_synth(module, seed):
    shapes = {k: tuple(v.shape) for k, v in module.state_dict().items()}
    sd = synth_state_dict(shapes, seed)
    module.load_state_dict(sd, strict=True)
    return sd


def main():
    ref_shims.install()
    import torch.nn as nn
    from slowfast.models.attention import attention_pool, MultiScaleAttention, MultiScaleBlock

    torch.set_grad_enabled(False)
    blobs = {}
    report = {}

    # ---- A. attention_pool ------------------------------------------------
    for c in POOL_CASES:
        name, B, h, d, thw = c["name"], c["B"], c["heads"], c["d"], c["thw"]
        L = thw[0] * thw[1] * thw[2] + (1 if c["cls"] else 0)
        pad = [k // 2 for k in c["kernel"]]
        if c["mode"] == "conv":
            pool = nn.Conv3d(d, d, c["kernel"], stride=c["stride"], padding=pad, groups=d, bias=False)
            norm = nn.LayerNorm(d)
            holder = nn.ModuleDict({"pool_q": pool, "norm_q": norm})
            sd = load_synth(holder, c["seed"])
            ln = (sd["norm_q.weight"], sd["norm_q.bias"], 1e-5)
            w = sd["pool_q.weight"]
        else:
            pool = nn.MaxPool3d(c["kernel"], c["stride"], pad, ceil_mode=False)
            norm, ln, w = None, None, None
        shape = (B, h, L, d) if c["ndim"] == 4 else (B, L, d)
        x = synth_input(c["seed"], name, shape)
        ref, ref_thw = attention_pool(x, pool, list(thw), has_cls_embed=c["cls"], norm=norm)
        ora, ora_thw = O.attention_pool(x, thw, mode=c["mode"], kernel=c["kernel"], stride=c["stride"],
                                        weight=w, has_cls=c["cls"], ln=ln)
        assert list(ref_thw) == list(ora_thw) == O.pooled_thw(thw, c["kernel"], c["stride"]), name
        e = rel_err(ora, ref)
        assert e < TOL, (name, e)
        report[name] = e
        blobs[name] = ref.numpy()
        blobs[name + ".thw"] = np.asarray(ref_thw, dtype=np.int64)

    # ---- B. MultiScaleAttention ---------------------------------------------
    for c in ATTN_CASES:
        name = c["name"]
        m = MultiScaleAttention(c["dim"], num_heads=c["heads"], qkv_bias=True, kernel_q=c["kernel_q"],
                                kernel_kv=c["kernel_kv"], stride_q=c["stride_q"], stride_kv=c["stride_kv"],
                                norm_layer=nn.LayerNorm, has_cls_embed=c["cls"], mode="conv",
                                use_query_residual_pool=c["residual"], expand_channel=c["dim_out"] != c["dim"],
                                expand_to_dim=c["dim_out"]).eval()
        sd = load_synth(m, c["seed"])
        thw = c["thw"]
        N = thw[0] * thw[1] * thw[2] + (1 if c["cls"] else 0)
        x = synth_input(c["seed"], name, (c["B"], N, c["dim"]))
        ref, ref_thw = m(x, list(thw))
        spec = O.BlockSpec(c["dim"], c["dim_out"], c["heads"], c["kernel_q"], c["kernel_kv"],
                           c["stride_q"], c["stride_kv"], 0.0, expand=c["dim_out"] != c["dim"])
        mv = O.MViTSpec([], [], 0, [], [], [], c["cls"], True, "conv", c["residual"], 0, True)
        ora, ora_thw = O.multiscale_attention(x, thw, sd, "", spec, mv)
        assert list(ref_thw) == list(ora_thw), name
        e = rel_err(ora, ref)
        assert e < TOL, (name, e)
        report[name] = e
        blobs[name] = ref.numpy()

    # ---- C. MultiScaleBlock ---------------------------------------------------
    from functools import partial
    for c in BLOCK_CASES:
        name = c["name"]
        m = MultiScaleBlock(dim=c["dim"], dim_out=c["dim_out"], num_heads=c["heads"], mlp_ratio=4.0,
                            qkv_bias=True, drop_rate=0.0, drop_path=0.0,
                            norm_layer=partial(nn.LayerNorm, eps=1e-6),
                            kernel_q=c["kernel_q"], kernel_kv=c["kernel_kv"], stride_q=c["stride_q"],
                            stride_kv=c["stride_kv"], mode="conv", has_cls_embed=c["cls"],
                            use_query_residual_pool=c["residual"],
                            channel_expand_front=c["expand_front"]).eval()
        sd = load_synth(m, c["seed"])
        thw = c["thw"]
        N = thw[0] * thw[1] * thw[2] + (1 if c["cls"] else 0)
        x = synth_input(c["seed"], name, (c["B"], N, c["dim"]))
        ref, ref_thw = m(x, list(thw))
        spec = O.BlockSpec(c["dim"], c["dim_out"], c["heads"], c["kernel_q"], c["kernel_kv"],
                           c["stride_q"], c["stride_kv"], 0.0,
                           expand=c["expand_front"] and c["dim_out"] != c["dim"])
        mv = O.MViTSpec([], [], 0, [], [], [], c["cls"], True, "conv", c["residual"], 0, True)
        ora, ora_thw = O.multiscale_block(x, thw, sd, "", spec, mv)
        assert list(ref_thw) == list(ora_thw), name
        e = rel_err(ora, ref)
        assert e < TOL, (name, e)
        report[name] = e
        blobs[name] = ref.numpy()

    # ---- D/E. whole models ------------------------------------------------------
    shapes_index = {}
    for c in MODEL_CASES:
        name = c["name"]
        cfg = ref_shims.ref_cfg(c["yaml"], tiny_cfg_overrides(c))
        model = ref_shims.ref_build_model(cfg, seed=0).eval()
        shapes = {k: list(v.shape) for k, v in model.state_dict().items()}
        shapes_index[name] = shapes
        sd = synth_state_dict(shapes, c["seed"])
        model.load_state_dict(sd, strict=True)
        x = synth_clip(c["seed"], c["B"], cfg.DATA.NUM_FRAMES, cfg.DATA.TRAIN_CROP_SIZE)
        feats = {}
        hk = model.head.register_forward_pre_hook(lambda mod, inp: feats.__setitem__("feat", inp[0].detach().clone()))
        ref = model([x])
        hk.remove()
        spec = O.derive_spec(cfg)
        ora, extra = O.mvit_forward(x, sd, spec, return_features=True)
        e = rel_err(ora, ref)
        ef = rel_err(extra["feat"], feats["feat"])
        assert e < TOL and ef < 1e-4, (name, e, ef)
        report[name] = e
        report[name + ".feat"] = ef
        blobs[name + ".probs"] = ref.numpy()
        blobs[name + ".feat"] = feats["feat"].numpy()
        # structural known-answers: per-block thw / dims as the reference built them
        blobs[name + ".blockdims"] = np.asarray(
            [[b.dim, b.dim_out, b.attn.num_heads] for b in model.blocks], dtype=np.int64)
        ora_dims = np.asarray([[b.dim_out if b.expand else b.dim, b.dim_out, b.heads] for b in spec.blocks])
        assert (blobs[name + ".blockdims"] == ora_dims).all(), name
        print(f"{name}: probs rel {e:.2e} feat rel {ef:.2e} params {sum(v.numel() for v in sd.values())}")

    # ---- G. window indexing / post-processing -------------------------------------
    import types
    from module_wrapper import ActionProposalFromVideoTemporalDataset as DS
    import aicity_inf_graph as G

    win = {}
    for n_frames, length, stride in [(18000, 64, 16), (1000, 64, 16), (777, 53, 13), (65, 64, 16), (10, 64, 16)]:
        ref_props = DS._get_proposals(None, None, length, stride, n_frames)
        ref_w = [(p[1], p[2]) for p in ref_props]
        assert ref_w == WO.window_list(n_frames, length, stride)
        fake = types.SimpleNamespace(video_num_frame=n_frames)
        idx = [DS._get_frame_idxs_uniform(fake, t0, t1, 16) for t0, t1 in ref_w]
        assert idx == [WO.frame_indices(t0, t1, 16, n_frames) for t0, t1 in ref_w]
        key = f"win_{n_frames}_{length}_{stride}"
        blobs[key] = np.asarray(ref_w, dtype=np.int64)
        blobs[key + ".idx"] = np.asarray(idx, dtype=np.int64)
        win[key] = len(ref_w)
    rng = np.random.RandomState(7)
    chunk_cases = [[.9, .9, .1, .9, .9, .9, .1, .1, .9], [.1, .9, .9, .9], [.9], [.1, .1], [.9, .9, .9],
                   list(rng.rand(200).astype(np.float32)), list((rng.rand(500) > 0.3).astype(np.float32))]
    chunk_out = []
    for i, sc in enumerate(chunk_cases):
        for thr in (0.5, 0.2, 0.95):
            ref_c = [(a, b, n, float(m)) for a, b, n, m, _ in G.get_chunks(np.asarray(sc, dtype=np.float32), thr)]
            ora_c = [(a, b, n, float(m)) for a, b, n, m in WO.get_chunks(np.asarray(sc, dtype=np.float32), thr)]
            assert ref_c == ora_c, (i, thr)
            chunk_out.append({"scores": [float(v) for v in sc], "thr": thr, "chunks": ref_c})
    preds = [(0, 4, np.array([1, 0], np.float32)), (2, 6, np.array([0, 1], np.float32))]
    assert np.array_equal(G.aggregate_predictions(preds, np.mean, 2), WO.aggregate(preds, 2, "mean"))
    wl = WO.window_list(300, 64, 16)
    pr = [(t0, t1, rng.rand(18).astype(np.float32)) for t0, t1 in wl]
    for how, fn in (("mean", np.mean), ("max", np.max)):
        ref_a = G.aggregate_predictions(pr, fn, 18)
        assert np.array_equal(ref_a, WO.aggregate(pr, 18, how))
        blobs["agg_" + how] = ref_a
    blobs["agg_in"] = np.stack([p[2] for p in pr])

    np.savez_compressed(os.path.join(OUT, "golden.npz"), **blobs)
    with open(os.path.join(OUT, "golden_index.json"), "w") as f:
        json.dump({"oracle_vs_reference_rel_err": report, "model_shapes": shapes_index,
                   "windows": win, "chunks": chunk_out,
                   "torch": torch.__version__}, f, indent=1)
    print(json.dumps(report, indent=1))
    print("wrote", os.path.join(OUT, "golden.npz"), os.path.getsize(os.path.join(OUT, "golden.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
