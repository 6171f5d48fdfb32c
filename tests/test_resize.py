"""uint8 bilinear resize (the reference's `cv2.resize(frame, (S, S), INTER_LINEAR)`, scripts/utils.py:207-211): bit-exact.

CPU: the NumPy restatement (oracle/resize_oracle.py) == the committed cv2 outputs (tests/golden/resize_golden.*) and, where
cv2 is importable, == cv2 itself on fresh random geometries.  GPU: mvit_resize_gather_u8 == restatement == fixtures == cv2,
including the frame-index gather (duplicates, reordering) and the full 540p -> 448 production geometry."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from resize_oracle import resize_linear_u8
from tests.golden.resize_cases import CASES, case_image

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "resize_golden.npz"))
INDEX = json.load(open(os.path.join(HERE, "golden", "resize_golden.json")))["cases"]


def _check_against_fixture(name, out):
    assert list(out.shape) == INDEX[name]["shape"]
    assert hashlib.sha256(np.ascontiguousarray(out).tobytes()).hexdigest() == INDEX[name]["sha256"], name
    g = GOLD[name]
    assert np.array_equal(out[: g.shape[0]], g)


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_matches_cv2_fixture(name):
    _, _, oh, ow = CASES[name]
    _check_against_fixture(name, resize_linear_u8(case_image(name), oh, ow))


def test_oracle_matches_cv2_live():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(7)
    for _ in range(25):
        H, W, oh, ow = (int(v) for v in rng.integers(2, 200, 4))
        img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        assert np.array_equal(resize_linear_u8(img, oh, ow), cv2.resize(img, (ow, oh), interpolation=cv2.INTER_LINEAR)), (H, W, oh, ow)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_cuda_resize_bit_exact(name):
    from aicity_action_b200 import ops
    _, _, oh, ow = CASES[name]
    img = case_image(name)
    out = ops.resize_gather_u8(torch.from_numpy(img)[None].cuda(), None, (oh, ow))[0].cpu().numpy()
    _check_against_fixture(name, out)
    assert np.array_equal(out, resize_linear_u8(img, oh, ow))


@pytest.mark.gpu
def test_cuda_resize_gathers_frames_by_index_like_the_reference_window_reader():
    """A window = 16 frame indices (module_wrapper.py:384-397, clamped repeats at the end of the video) -> resized clip."""
    from aicity_action_b200 import ops
    from aicity_action_b200.sliding_window import frame_indices
    rng = np.random.default_rng(3)
    frames = rng.integers(0, 256, (40, 54, 96, 3), dtype=np.uint8)
    idx = frame_indices(16, 80, 16, 40) + frame_indices(0, 64, 16, 40)          # the first one clamps to frame 39
    uniq = sorted(set(idx))
    pos = {f: i for i, f in enumerate(uniq)}
    dev_frames = torch.from_numpy(frames[uniq]).cuda()                          # only the needed frames travel
    sel = torch.tensor([pos[f] for f in idx], dtype=torch.int32, device="cuda")
    out = ops.resize_gather_u8(dev_frames, sel, (64, 64)).cpu().numpy()
    assert out.shape == (32, 64, 64, 3)
    for k, f in enumerate(idx):
        assert np.array_equal(out[k], resize_linear_u8(frames[f], 64, 64)), (k, f)
    try:
        import cv2
    except ImportError:
        return
    assert np.array_equal(out[5], cv2.resize(frames[idx[5]], (64, 64), interpolation=cv2.INTER_LINEAR))


@pytest.mark.gpu
def test_cuda_resize_random_geometries_vs_cv2():
    cv2 = pytest.importorskip("cv2")
    from aicity_action_b200 import ops
    rng = np.random.default_rng(11)
    for _ in range(20):
        H, W, oh, ow = (int(v) for v in rng.integers(2, 300, 4))
        img = rng.integers(0, 256, (2, H, W, 3), dtype=np.uint8)
        out = ops.resize_gather_u8(torch.from_numpy(img).cuda(), None, (oh, ow)).cpu().numpy()
        for k in range(2):
            assert np.array_equal(out[k], cv2.resize(img[k], (ow, oh), interpolation=cv2.INTER_LINEAR)), (H, W, oh, ow)
