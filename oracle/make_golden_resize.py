#!/usr/bin/env python3
"""Golden vectors for the uint8 bilinear resize: outputs of `cv2.resize(img, (W', H'), interpolation=cv2.INTER_LINEAR)`
(the call the reference makes, scripts/utils.py:207-211) on seeded random images -> tests/golden/resize_golden.npz.

    python oracle/make_golden_resize.py           (needs opencv-python; 4.13.0 in this image)

Small geometries are stored whole; the two production geometries (540p -> 448 / 224) are stored as SHA-256 digests plus
their first 8 rows.  Inputs are regenerated from the seed by tests/golden/resize_cases.py.  TEST INFRASTRUCTURE."""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from tests.golden.resize_cases import CASES, case_image  # noqa: E402
from resize_oracle import resize_linear_u8  # noqa: E402

if __name__ == "__main__":
    import cv2
    cv2.setNumThreads(1)
    blobs, index = {}, {"opencv": cv2.__version__, "cases": {}}
    for name, (H, W, oh, ow) in CASES.items():
        img = case_image(name)
        ref = cv2.resize(img, (ow, oh), interpolation=cv2.INTER_LINEAR)
        assert np.array_equal(resize_linear_u8(img, oh, ow), ref), name       # pins the restatement to cv2
        index["cases"][name] = {"sha256": hashlib.sha256(ref.tobytes()).hexdigest(), "shape": list(ref.shape)}
        blobs[name] = ref if ref.size <= 64 * 1024 else ref[:8]
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "resize_golden.npz"), **blobs)
    with open(os.path.join(ROOT, "tests", "golden", "resize_golden.json"), "w") as f:
        json.dump(index, f, indent=1)
    print(json.dumps(index["cases"], indent=1))
