"""Synthetic 3-view sliding-window scores for the end-to-end post-processing parity case (scripts/aicity_inf.py).

Pure function of the seed, shared by `oracle/make_golden_postprocess.py` (which feeds them to the UNMODIFIED reference
script and stores its output as tests/golden/aicity_inf_expected.txt) and by the tests."""
from __future__ import annotations

import numpy as np

VIDS = {"1": ["Dashboard_user_id_1_NoAudio_3.MP4", "Rear_view_user_id_1_NoAudio_3.MP4", "Rightside_window_user_id_1_NoAudio_3.MP4"],
        "2": ["Dashboard_user_id_2_NoAudio_5.MP4", "Rear_view_user_id_2_NoAudio_5.MP4", "Rightside_window_user_id_2_NoAudio_5.MP4"]}
NUM_CLASS = 18
N_WINDOWS = 120                      # 120 windows of 64 frames, stride 16 -> 1968 frames (~66 s)
THRESHOLDS = {c: (0.35 if c % 3 else 0.5) for c in range(1, NUM_CLASS)}


def window_scores(file_id: str, seed: int = 2022):
    """[(t0, t1, float32 scores[18])]: per class one or two plateaus whose position / length differ per view."""
    rng = np.random.default_rng([seed, sum(file_id.encode())])
    s = 0.02 + 0.1 * rng.random((N_WINDOWS, NUM_CLASS))
    for c in range(1, NUM_CLASS):
        for _ in range(int(rng.integers(0, 3))):
            a = int(rng.integers(0, N_WINDOWS - 12))
            n = int(rng.integers(3, 12))
            s[a:a + n, c] = 0.55 + 0.4 * rng.random(n)
    s = (s / s.sum(1, keepdims=True)).astype(np.float32) * 1.0
    s = np.minimum(s * 4.0, 1.0).astype(np.float32)
    return [(16 * i, 16 * i + 64, s[i]) for i in range(N_WINDOWS)]


def write_inputs(d: str):
    """pickles + threshold file + video-id csv in the layout scripts/aicity_inf.py reads; returns the three paths."""
    import os
    import pickle
    os.makedirs(os.path.join(d, "pkl"), exist_ok=True)
    for files in VIDS.values():
        for f in files:
            with open(os.path.join(d, "pkl", f + ".pkl"), "wb") as fh:
                pickle.dump(window_scores(f), fh)
    thr = os.path.join(d, "thres.txt")
    with open(thr, "w") as fh:
        for c, t in THRESHOLDS.items():
            fh.write(f"{c} {t}\n")
    csv = os.path.join(d, "video_ids.csv")
    with open(csv, "w") as fh:
        fh.write("video_id,video_files\n")
        for vid, files in VIDS.items():
            fh.write(",".join([vid] + files) + "\n")
    return os.path.join(d, "pkl"), thr, csv
