"""CPU restatement (NumPy integer arithmetic) of OpenCV's `cv2.resize(u8 image, (W', H'), interpolation=INTER_LINEAR)`.

TEST INFRASTRUCTURE ONLY — the checker for the CUDA resize kernel (aicity_action_b200/csrc/resize.cu); never imported by
the product.  The reference resizes each uint8 frame with this call before casting to float (scripts/utils.py:207-211,
via module_wrapper.py:326-331 with keep_scale=False), so "bit-exact" input parity means reproducing OpenCV's fixed-point
bilinear path: the algorithm lives in the third-party dependency opencv-python (unpinned by the reference; 4.13.0 in this
image), modules/imgproc/src/resize.cpp — `hal::resize` coefficient tables + `HResizeLinear<uchar,int,short,2048>` +
`VResizeLinear<uchar,int,short,FixedPtCast<int,uchar,22>>`:

  fx = float((dx + 0.5) * scale_x - 0.5);  sx = floor(fx);  fx -= sx          (scale_x = 1 / (W' / W) in double)
  sx < 0 -> (sx, fx) = (0, 0);   sx >= W-1 -> (sx, fx) = (W-1, 0)              (columns clamp, weight goes to the edge pixel)
  alpha = (round_half_even((1 - fx) * 2048), round_half_even(fx * 2048)) as int16; rows likewise (beta), but rows only
  clamp their INDEX (sy, sy+1 -> [0, H-1]) and keep their weights
  horizontal: D[dx] = S[sx] * alpha0 + S[sx + 1] * alpha1                       (int32, scale 2^11)
  vertical:   out = (((beta0 * (D0 >> 4)) >> 16) + ((beta1 * (D1 >> 4)) >> 16) + 2) >> 2

Pinned against cv2 itself: `tests/test_resize_cpu.py` compares this restatement with `cv2.resize` (installed in this image)
on random images over the geometries the pipeline uses, and tests/golden/resize_golden.npz stores cv2 outputs made by
`oracle/make_golden_resize.py`.
"""
from __future__ import annotations

import numpy as np


def linear_coeffs(src: int, dst: int, clamp_weights: bool):
    """(index int32[dst], coefficient int16[dst, 2]) of OpenCV's fixed-point linear resize along one axis."""
    scale = 1.0 / (float(dst) / float(src))                       # double, as hal::resize derives it from inv_scale
    d = np.arange(dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int32)
    f = (f - s.astype(np.float32)).astype(np.float32)
    if clamp_weights:                                             # x axis: hal::resize zeroes fx at both borders
        lo, hi = s < 0, s >= src - 1
        f = np.where(lo | hi, np.float32(0), f)
        s = np.where(lo, 0, np.where(hi, src - 1, s)).astype(np.int32)
    c0 = np.rint((np.float32(1.0) - f) * np.float32(2048.0)).astype(np.int32)     # saturate_cast<short>: nearest-even
    c1 = np.rint(f * np.float32(2048.0)).astype(np.int32)
    coef = np.clip(np.stack([c0, c1], 1), -32768, 32767).astype(np.int16)
    return s, coef


def resize_linear_u8(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """img [H, W, C] uint8 -> [out_h, out_w, C] uint8, == cv2.resize(img, (out_w, out_h), interpolation=cv2.INTER_LINEAR)."""
    assert img.dtype == np.uint8 and img.ndim == 3
    H, W, _ = img.shape
    if (H, W) == (out_h, out_w):
        return img.copy()
    sx, ax = linear_coeffs(W, out_w, True)
    sy, ay = linear_coeffs(H, out_h, False)
    x1 = np.minimum(sx + 1, W - 1)
    src = img.astype(np.int32)
    rows = src[:, sx, :] * ax[:, 0].astype(np.int32)[None, :, None] + src[:, x1, :] * ax[:, 1].astype(np.int32)[None, :, None]
    y0, y1 = np.clip(sy, 0, H - 1), np.clip(sy + 1, 0, H - 1)
    b0, b1 = ay[:, 0].astype(np.int32)[:, None, None], ay[:, 1].astype(np.int32)[:, None, None]
    out = (((b0 * (rows[y0] >> 4)) >> 16) + ((b1 * (rows[y1] >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)
