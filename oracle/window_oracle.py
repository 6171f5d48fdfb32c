"""CPU oracle for the integer / post-processing side of sliding-window localisation.

*** TEST INFRASTRUCTURE *** — restates, in plain Python/NumPy loops, what the
reference does around the model in `scripts/` (citations are file:line under
/root/reference).  Pinned by `oracle/make_golden.py` against the reference's own
functions and by the hand-checked vectors of SURVEY.md Appendix C.
"""
from __future__ import annotations

import numpy as np
import torch


def window_list(num_frames: int, length: int = 64, stride: int = 16):
    """module_wrapper.py:246-253 — windows start every `stride` frames; t1 may pass the end."""
    out = []
    i = 0
    while i < num_frames:
        out.append((i, i + length))
        i += stride
    return out


def fps_adjust(length: int, stride: int, video_fps: float, target_fps: float):
    """module_wrapper.py:215-232 — rescale only when |fps - target| > 2; int() truncation."""
    if abs(video_fps - target_fps) > 2.0:
        r = video_fps / target_fps
        return int(r * length), int(r * stride)
    return length, stride


def frame_indices(t0: int, t1: int, num: int, num_frames: int):
    """module_wrapper.py:384-397 — float32 linspace (endpoint inclusive), clamp, truncate."""
    idx = torch.linspace(t0, t1, num)
    idx = torch.clamp(idx, 0, num_frames - 1).long()
    return idx.numpy().tolist()


def get_chunks(scores, threshold):
    """aicity_inf_graph.py:288-309 — (start, end, length, mean) runs of score >= threshold.

    Quirks kept: the below-threshold frame that closes a run is INCLUDED in (end, length,
    mean); a run touching the last frame is emitted only if it began earlier."""
    s = np.asarray(scores)
    n = len(s)
    chunks = []
    start = None
    for f in range(n):
        if s[f] >= threshold:
            if start is None:
                start = f
            elif f == n - 1:
                chunks.append((start, f, f - start + 1, np.mean(s[start:f + 1])))
                start = None
        elif start is not None:
            chunks.append((start, f, f - start + 1, np.mean(s[start:f + 1])))
            start = None
    return chunks


def aggregate(pred_list, num_class, how="mean"):
    """aicity_inf_graph.py:313-351 — per-frame mean/max over every window covering the frame.

    Frame axis = [min t0, max t1); uncovered frames are zeros.  float32 throughout."""
    lo = min(min(p[0] for p in pred_list), min(p[1] for p in pred_list))
    hi = max(max(p[0] for p in pred_list), max(p[1] for p in pred_list))
    per_frame = [[] for _ in range(hi - lo)]
    for t0, t1, sc in pred_list:
        assert len(sc) == num_class
        for t in range(t0, t1):
            per_frame[t - lo].append(np.asarray(sc, dtype=np.float32))
    fn = np.mean if how == "mean" else np.max
    rows = []
    for lst in per_frame:
        if not lst:
            rows.append(np.zeros((num_class,), dtype=np.float32))
        else:
            rows.append(fn(np.vstack(lst), axis=0))
    return np.vstack(rows)
