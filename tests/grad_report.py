"""Per-parameter gradient error of the tiny MViT (CUDA backward vs autograd through the CPU oracle)."""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # tests/ -> repo root
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import mvit_oracle as O  # noqa: E402
from aicity_action_b200.config import aicity_cfg  # noqa: E402
from aicity_action_b200.mvit import MViT  # noqa: E402
from tests.golden.cases import MODEL_CASES, tiny_cfg_overrides  # noqa: E402
from tests.golden.synth import synth_clip, synth_state_dict  # noqa: E402

c = MODEL_CASES[0]
cfg = aicity_cfg(c["yaml"], tiny_cfg_overrides(c) + ["MVIT.DROPPATH_RATE", 0.0, "MODEL.DROPOUT_RATE", 0.0])
m = MViT(cfg).train()
sd = synth_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, c["seed"])
m.load_state_dict(sd)
m = m.cuda()
x = synth_clip(c["seed"], c["B"], cfg.DATA.NUM_FRAMES, cfg.DATA.TRAIN_CROP_SIZE)
labels = torch.arange(c["B"]) % cfg.MODEL.NUM_CLASSES
sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
F.cross_entropy(O.mvit_forward(x, sdr, O.derive_spec(cfg), training=True), labels).backward()
gmax = max(v.grad.abs().max().item() for v in sdr.values())
for dtype in (torch.float32, torch.bfloat16):
    m.zero_grad(set_to_none=True)
    F.cross_entropy(m([x.cuda().to(dtype)]), labels.cuda()).backward()
    print(f"--- {dtype}  (global max |grad| {gmax:.3e})")
    for k, p in m.named_parameters():
        r = sdr[k].grad
        e = (p.grad.cpu() - r).abs().max().item()
        print(f"{k:45s} ref {r.abs().max().item():.3e}  err/ref {e / max(r.abs().max().item(), 1e-30):.3e}  err/global {e / gmax:.3e}")
