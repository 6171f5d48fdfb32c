"""Post-processing of sliding-window scores into action segments (scripts/aicity_inf_graph.py:288-351,
scripts/aicity_inf.py:87-129).  Host-side NumPy; semantics — including the reference's quirks — are kept
verbatim because "segment boundaries bit-exact" is part of the parity contract."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def get_chunks(score_list: Sequence[float], threshold: float):
    """aicity_inf_graph.py:288-309.  Runs of score >= threshold as (start, end, length, mean, scores).
    A run closed by a below-threshold frame INCLUDES that frame; a run that reaches the last frame is emitted
    only if it began before it (a lone above-threshold final frame is dropped)."""
    s = np.asarray(score_list)
    n = len(s)
    above = s >= threshold
    chunks = []
    start = None
    for f in range(n):
        if above[f]:
            if start is None:
                start = f
            elif f == n - 1:
                chunks.append((start, f, f - start + 1, np.mean(s[start:f + 1]), s[start:f + 1]))
                start = None
        elif start is not None:
            chunks.append((start, f, f - start + 1, np.mean(s[start:f + 1]), s[start:f + 1]))
            start = None
    return chunks


def aggregate_predictions(pred_list, aggregate_func, num_class: int) -> np.ndarray:
    """aicity_inf_graph.py:313-351.  Per-frame aggregation (np.mean / np.max over axis 0) of every window that
    covers the frame; frame axis = [min t0, max t1); uncovered frames are zeros.

    Vectorised per run of frames covered by the same window set (the reference loops frame by frame and window by
    window in Python); within a run the stacked scores are identical, so calling `aggregate_func` once per run gives
    bit-identical float32 results."""
    lo = min(min(p[0] for p in pred_list), min(p[1] for p in pred_list))
    hi = max(max(p[0] for p in pred_list), max(p[1] for p in pred_list))
    n = hi - lo
    out = np.zeros((n, num_class), dtype=np.float32)
    if n <= 0:
        return out
    # frames where the covering set changes
    cuts = sorted({lo, hi} | {min(max(p[0], lo), hi) for p in pred_list} | {min(max(p[1], lo), hi) for p in pred_list})
    scores = [np.asarray(p[2]) for p in pred_list]
    for sc in scores:
        assert len(sc) == num_class
    for a, b in zip(cuts[:-1], cuts[1:]):
        if a >= b:
            continue
        cover = [scores[k] for k, p in enumerate(pred_list) if p[0] <= a and a < p[1]]   # reference append order
        if cover:
            out[a - lo:b - lo] = aggregate_func(np.vstack(cover), axis=0)
    return out


def boundaries_to_seconds(start_frame: int, end_frame: int, fps: float = 30.0) -> Tuple[float, float]:
    """aicity_inf.py:99,123: frames -> seconds, Python round() (banker's), then +1 / -1 second."""
    return round(start_frame / fps) + 1.0, round(end_frame / fps) - 1.0


def top_chunk_per_class(agg_scores: np.ndarray, thresholds: Sequence[float], by: str = "score"):
    """aicity_inf.py:87-99: per class, the best chunk of the per-frame score track (by mean score or by length)."""
    out = {}
    for c in range(agg_scores.shape[1]):
        chunks = get_chunks(agg_scores[:, c], thresholds[c])
        if not chunks:
            continue
        key = (lambda ch: ch[3]) if by == "score" else (lambda ch: ch[2])
        out[c] = sorted(chunks, key=key, reverse=True)[0]
    return out
