"""Geometries / seeded inputs of the uint8 resize parity cases (shared by oracle/make_golden_resize.py and the tests)."""
import zlib

import numpy as np

# name -> (H, W, out_h, out_w)
CASES = {
    "p540_to_448": (540, 960, 448, 448),        # BASELINE config 3: 540p views -> 448 x 448, aspect ignored
    "p540_to_224": (540, 960, 224, 224),
    "p1080_to_448": (1080, 1920, 448, 448),
    "small_down": (72, 96, 64, 64),
    "small_mixed": (54, 96, 64, 64),             # height up, width down
    "odd_up": (37, 53, 64, 80),
    "half": (64, 96, 32, 48),                    # exact 2x
    "same": (48, 48, 48, 48),
    "tiny_src": (2, 3, 16, 16),                  # every border clamp
}


def case_image(name: str) -> np.ndarray:
    H, W, _, _ = CASES[name]
    rng = np.random.default_rng(zlib.crc32(name.encode()))
    img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    img[: H // 3] = (img[: H // 3] // 64) * 64 + 31          # flat-ish regions and hard edges as well as noise
    return img
