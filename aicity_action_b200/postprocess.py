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


# ----------------------------------------------------------------------------- 3-view merge + submission writer
def localise_actions(preds_by_file, thresholds, views_by_vid, num_class: int = 18, agg_method: str = "avg",
                     sort_single: str = "score", sort_multi: str = "length", use_num_chunk: int = 1,
                     video_fps: float = 30.0, log=None):
    """scripts/aicity_inf.py:36-126, the whole post-processing of one submission.

    preds_by_file: {file_id: [(t0, t1, scores[num_class])]} — what the sliding-window runner pickles per video;
    thresholds:    {action_id: threshold} in the order of the threshold file (it fixes the output order);
    views_by_vid:  {vid: [file_id of view 1, 2, 3]} in the order of the video-id csv.
    Per file and action: per-frame aggregation, threshold chunks, stable sort by mean score (or length), top-k, frames ->
    seconds.  Per vid and action: the chunks of the three synchronised views concatenated in view order, stable sort by
    LENGTH (default) or score, top-k, then Python `round()` (banker's) +1 / -1 second.
    Returns [(vid, action_id, start_sec, end_sec)] in the reference's order."""
    aggregate_func = np.mean if agg_method == "avg" else np.max
    say = log if log is not None else (lambda msg: None)
    action_chunks = {}
    for file_id, pred in preds_by_file.items():
        frames = aggregate_predictions(pred, aggregate_func, num_class)
        per_action = {}
        for action_id, thr in thresholds.items():
            chunks = get_chunks(frames[:, action_id], thr)
            if not chunks:
                say("warning, %s %s got no action chunks" % (file_id, action_id))
                continue
            chunks.sort(key=(lambda c: c[2]) if sort_single == "length" else (lambda c: c[3]), reverse=True)
            per_action[action_id] = [(s / video_fps, e / video_fps, n, m) for s, e, n, m, _ in chunks[:use_num_chunk]]
        action_chunks[file_id] = per_action
    outputs = []
    for vid, files in views_by_vid.items():
        for action_id in thresholds:
            merged = [c for f in files for c in action_chunks[f].get(action_id, [])]
            if not merged:
                say("warning, %s %s has no action chunks" % (vid, action_id))
                continue
            merged.sort(key=(lambda c: c[2]) if sort_multi == "length" else (lambda c: c[3]), reverse=True)
            for c in merged[:use_num_chunk]:
                outputs.append((vid, action_id, round(c[0]) + 1.0, round(c[1]) - 1.0))
    return outputs


def write_submission(outputs, path: str) -> None:
    """aicity_inf.py:128-130: one `vid action_id start end` line per segment, seconds with six decimals."""
    with open(path, "w") as f:
        for vid, action_id, start, end in outputs:
            f.write("%s %s %.6f %.6f\n" % (vid, action_id, start, end))


def read_thresholds(path: str) -> dict:
    """aicity_inf.py:47-50: `action_id threshold` per line; insertion order is the output order."""
    out = {}
    with open(path) as f:
        for line in f:
            if line.strip():
                a, t = line.strip().split()
                out[int(a)] = float(t)
    return out


def read_video_ids(path: str) -> dict:
    """aicity_inf.py:52-58: csv with a header line, then `vid,file1,file2,file3`."""
    out = {}
    with open(path) as f:
        for line in f.readlines()[1:]:
            if line.strip():
                vid, f1, f2, f3 = line.strip().split(",")
                out[vid] = [f1, f2, f3]
    return out


def main(argv=None) -> int:
    """Same positional arguments and options as scripts/aicity_inf.py (argparse block at :15-34)."""
    import argparse
    import os
    import pickle
    ap = argparse.ArgumentParser(prog="python -m aicity_action_b200.postprocess")
    ap.add_argument("pred_pickle_path")
    ap.add_argument("thres_file")
    ap.add_argument("vid_csv")
    ap.add_argument("output_file")
    ap.add_argument("--num_class", default=18, type=int)
    ap.add_argument("--agg_method", default="avg", choices=["avg", "max"])
    ap.add_argument("--chunk_sort_base_single_vid", default="score", choices=["score", "length"])
    ap.add_argument("--chunk_sort_base_multi_vid", default="length", choices=["score", "length"])
    ap.add_argument("--use_num_chunk", default=1, type=int)
    a = ap.parse_args(argv)
    views = read_video_ids(a.vid_csv)
    preds = {}
    for files in views.values():
        for fid in files:
            with open(os.path.join(a.pred_pickle_path, "%s.pkl" % fid), "rb") as f:
                preds[fid] = pickle.load(f)
    outputs = localise_actions(preds, read_thresholds(a.thres_file), views, a.num_class, a.agg_method,
                               a.chunk_sort_base_single_vid, a.chunk_sort_base_multi_vid, a.use_num_chunk, log=print)
    print("total pred %s" % len(outputs))
    write_submission(outputs, a.output_file)
    return 0


if __name__ == "__main__":
    import sys
    sys.exit(main())
