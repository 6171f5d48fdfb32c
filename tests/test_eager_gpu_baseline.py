"""The "before" number for a GPU user: the reference algorithm as stock PyTorch ops (the oracle restatement,
bit-identical to the reference modules on CPU) run eagerly on the same B200, next to the CUDA path, same weights and
inputs.  Writes gpurun_out/eager_gpu_baseline.json when that directory exists; asserts parity and that the fused path
is not slower.  MViTv2-B 16x4 @448, batch 8, bf16."""
import json
import os
import time

import pytest
import torch
import torch.nn.functional as F

import mvit_oracle as O
from aicity_action_b200.config import aicity_cfg
from aicity_action_b200.mvit import MViT

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _time(fn, warm=2, reps=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def test_eager_pytorch_on_gpu_vs_fused_path():
    if torch.cuda.get_device_properties(0).total_memory < 100 * 2 ** 30:
        pytest.skip("needs a large-memory GPU: the eager path materialises every score tensor")
    B = 8
    cfg = aicity_cfg("MVITV2_FULL_B_16x4_CONV_448.yaml", ["MVIT.DROPPATH_RATE", 0.0, "MODEL.DROPOUT_RATE", 0.0])
    torch.manual_seed(0)
    model = MViT(cfg).cuda().eval()
    spec = O.derive_spec(cfg)
    sd16 = {k: v.detach().bfloat16() for k, v in model.state_dict().items()}
    x = torch.randn(B, 3, cfg.DATA.NUM_FRAMES, 448, 448, device="cuda").bfloat16()
    labels = torch.randint(0, cfg.MODEL.NUM_CLASSES, (B,), device="cuda")

    with torch.no_grad():
        ref = O.mvit_forward(x, sd16, spec).float()
        got = model([x]).float()
        t_eager_fwd = _time(lambda: O.mvit_forward(x, sd16, spec))
        t_ours_fwd = _time(lambda: model([x]))
    err = ((got - ref).abs().max() / ref.abs().max()).item()
    assert err < 2e-2, err

    # training step: forward + backward (no optimizer: identical on both sides)
    sdg = {k: v.detach().bfloat16().requires_grad_(True) for k, v in model.state_dict().items()}

    def eager_step():
        for v in sdg.values():
            v.grad = None
        F.cross_entropy(O.mvit_forward(x, sdg, spec, training=True).float(), labels).backward()

    model.train()

    def ours_step():
        model.zero_grad(set_to_none=True)
        F.cross_entropy(model([x]).float(), labels).backward()

    torch.cuda.reset_peak_memory_stats()
    t_eager_trn = _time(eager_step, warm=1, reps=3)
    mem_eager = torch.cuda.max_memory_allocated() / 2 ** 30
    for v in sdg.values():
        v.grad = None
    torch.cuda.empty_cache()
    torch.cuda.reset_peak_memory_stats()
    t_ours_trn = _time(ours_step, warm=1, reps=3)
    mem_ours = torch.cuda.max_memory_allocated() / 2 ** 30
    res = {"config": "MVITV2_FULL_B_16x4_CONV_448, batch 8, bf16, one B200",
           "eager_pytorch_forward_ms": t_eager_fwd, "fused_forward_ms": t_ours_fwd,
           "forward_speedup": t_eager_fwd / t_ours_fwd, "forward_rel_inf_error": err,
           "eager_pytorch_fwd_bwd_ms": t_eager_trn, "fused_fwd_bwd_ms": t_ours_trn,
           "fwd_bwd_speedup": t_eager_trn / t_ours_trn,
           "eager_peak_mem_gb": mem_eager, "fused_peak_mem_gb": mem_ours,
           "note": "eager = the oracle restatement of attention.py / video_model_builder.py as stock torch ops in bf16"}
    print(json.dumps(res))
    out_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, "eager_gpu_baseline.json"), "w") as f:
            json.dump(res, f, indent=1)
    assert t_ours_fwd < t_eager_fwd and t_ours_trn < t_eager_trn
