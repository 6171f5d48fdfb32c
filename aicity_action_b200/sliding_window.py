"""Sliding-window temporal localisation over long driver videos, sharded across the GPUs of one box.

B200 form of scripts/run_action_classification_temporal_inf.py:75-130 + the window/frame logic of
scripts/module_wrapper.py:215-253, 304-397 (reference file:line cited per function):

  * window list and per-window frame indices are computed on the host exactly as the reference does
    (integer loop; float32 `torch.linspace` + clamp + truncate) — bit-exact by construction;
  * windows are dealt to ranks round-robin (window w -> rank w % R); every rank runs the replicated model
    on batches of its own windows: uint8 frames go up over PCIe (1 byte/sample), normalisation happens on
    the device (mvit_preprocess_u8_fwd), the forward is the tcgen05 path;
  * per-window probabilities are exchanged with ONE collective per video — an all_gather of a padded
    [ceil(n/R), 2 + classes] fp32/int64 pair over NCCL (NVLink/NVSwitch; gloo in CPU tests) — and put back in
    window order, which equals the reference's `pred_list.sort(key=t0)` order (module_wrapper windows have
    strictly increasing t0), so rank 0 writes byte-identical (t0, t1, scores) lists.
"""
from __future__ import annotations

import os
import pickle
from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import torch


# ----------------------------------------------------------------------------- host-side indexing
def fps_adjusted_window(length: int, stride: int, video_fps: float, target_fps: float = 30.0) -> Tuple[int, int]:
    """module_wrapper.py:215-232: windows are rescaled only when |video_fps - target_fps| > 2 (int() truncation)."""
    if abs(video_fps - target_fps) > 2.0:
        r = video_fps / target_fps
        return int(r * length), int(r * stride)
    return length, stride


def window_list(num_frames: int, length: int = 64, stride: int = 16) -> List[Tuple[int, int]]:
    """module_wrapper.py:246-253: (t0, t1) for t0 in range(0, num_frames, stride); t1 may pass the end."""
    return [(t0, t0 + length) for t0 in range(0, num_frames, stride)]


def frame_indices(t0: int, t1: int, num: int, num_frames: int) -> List[int]:
    """module_wrapper.py:384-397: `num` points uniformly in [t0, t1] (endpoint inclusive, float32), clamped to
    the video and truncated.  torch.linspace is called as the reference calls it, not re-derived."""
    idx = torch.linspace(t0, t1, num)
    return torch.clamp(idx, 0, num_frames - 1).long().numpy().tolist()


def coalesce_rows(rows: Sequence[int]) -> List[List[int]]:
    """[start, length] runs of consecutive rows, in order: the uploads of one batch from a pinned frame store."""
    runs: List[List[int]] = []
    for r in rows:
        if runs and runs[-1][0] + runs[-1][1] == r:
            runs[-1][1] += 1
        else:
            runs.append([int(r), 1])
    return runs


def shard_windows(n_windows: int, rank: int, world: int) -> List[int]:
    """Round-robin deal: window w belongs to rank w % world."""
    return list(range(rank, n_windows, world))


# ----------------------------------------------------------------------------- video sources
class SyntheticVideo:
    """Procedural stand-in for a decoded + resized video: frame f of video `seed` is a deterministic function of
    (seed, f) only, so any rank can materialise any frame without I/O.  Frames are uint8 [S, S, 3] RGB — what the
    reference holds after `cv2.resize(frame, (S, S))` (scripts/utils.py:207-211)."""

    PERIOD = 61      # distinct frames per video: frame f is texture f % PERIOD (enough that every window is unique)

    def __init__(self, seed: int, num_frames: int, size: int, fps: float = 30.0, raw_hw=None):
        """`raw_hw=(H, W)`: the video is held at its native resolution (e.g. (540, 960)) and resized to `size` per window
        like the reference does — on the device by the runner (`get_raw_into`), or on the host with cv2 (`get_batch`)."""
        self.seed, self.num_frames, self.size, self.fps = seed, num_frames, size, fps
        self.raw_hw = tuple(raw_hw) if raw_hw is not None else None
        h, w = self.raw_hw if self.raw_hw is not None else (size, size)
        g = torch.Generator().manual_seed(seed)
        self._base = torch.randint(0, 256, (h, w, 3), dtype=torch.uint8, generator=g)
        self._ramp = torch.arange(h, dtype=torch.int32).view(h, 1, 1)
        self._hw = (h, w)
        self._pool = None
        self.pinned_store = False          # True: the frame store is pinned and offered to the runner (raw_frames_pinned)

    def __len__(self):
        return self.num_frames

    def _texture(self, k: int) -> torch.Tensor:
        # cheap, deterministic: circular shift of a base texture plus a ramp
        img = torch.roll(self._base, shifts=(k % self._hw[0], (3 * k) % self._hw[1]), dims=(0, 1)).to(torch.int32)
        return ((img + (self._ramp * (k % 7))) % 256).to(torch.uint8)

    def frame(self, f: int) -> torch.Tensor:
        return self._texture(int(f) % self.PERIOD)

    def get_batch(self, idxs: Sequence[int]) -> torch.Tensor:
        """[T, S, S, 3] uint8: one gather from the (lazily built) texture pool, so that host-side frame synthesis does not
        bound the benchmark the way a Python loop over frames did."""
        raw = self._textures()[torch.as_tensor([int(i) % self.PERIOD for i in idxs])]
        if self.raw_hw is None or self._hw == (self.size, self.size):
            return raw
        import cv2
        out = np.empty((len(idxs), self.size, self.size, 3), dtype=np.uint8)
        for n in range(len(idxs)):                       # the reference's host path (scripts/utils.py:207-211)
            cv2.resize(raw[n].numpy(), (self.size, self.size), dst=out[n], interpolation=cv2.INTER_LINEAR)
        return torch.from_numpy(out)

    def _textures(self) -> torch.Tensor:
        if self._pool is None:
            self._pool = torch.stack([self._texture(k) for k in range(self.PERIOD)])
        return self._pool

    def get_batch_into(self, idxs: Sequence[int], out: torch.Tensor) -> None:
        """Same frames as `get_batch`, written straight into `out` ([T, S, S, 3] uint8, e.g. a slice of a pinned staging
        buffer): one copy per frame instead of gather + stack + pin."""
        if self.raw_hw is not None and self._hw != (self.size, self.size):
            out.copy_(self.get_batch(idxs))
            return
        torch.index_select(self._textures(), 0, torch.as_tensor([int(i) % self.PERIOD for i in idxs]), out=out)

    def get_raw_into(self, idxs: Sequence[int], out: torch.Tensor) -> None:
        """Native-resolution frames into `out` [n, H, W, 3] (pinned staging); the runner resizes them on the device."""
        torch.index_select(self._textures(), 0, torch.as_tensor([int(i) % self.PERIOD for i in idxs]), out=out)

    def raw_frames_pinned(self):
        """(frame store [P, H, W, 3] uint8 in PINNED host memory, frame index -> row) when the decoded frames live in pinned
        memory (a decoder's pinned output ring, a pre-decoded video): the runner then uploads each frame straight from there
        (one DMA per frame, no staging copy on the host).  None: frames go through `get_raw_into` and the staging ring."""
        if self.raw_hw is None or not self.pinned_store or not torch.cuda.is_available():
            return None
        pool = self._textures()
        if not pool.is_pinned():
            self._pool = pool = pool.pin_memory()
        return pool, (lambda f: int(f) % self.PERIOD)


class ArrayVideo:
    """Decoded RGB frames held in host memory ([N, H, W, 3] uint8 array or memmap) — what the reference's decord reader
    yields.  `get_batch` resizes each requested frame the way the reference does before the cast to float: uint8
    `cv2.resize(frame, (S, S), INTER_LINEAR)`, aspect ratio ignored (scripts/utils.py:207-211 with keep_scale=False,
    module_wrapper.py:326-331), so the clips the model sees are bit-identical to the reference's."""

    def __init__(self, frames, size: int, fps: float = 30.0):
        self.frames, self.size, self.fps = frames, size, fps
        self.raw_hw = (int(frames.shape[1]), int(frames.shape[2]))

    def __len__(self):
        return int(self.frames.shape[0])

    def get_raw_into(self, idxs: Sequence[int], out: torch.Tensor) -> None:
        """Decoded frames at native resolution into `out` [n, H, W, 3] (pinned staging): the device resizes them."""
        dst = out.numpy()
        src = self.frames.numpy() if isinstance(self.frames, torch.Tensor) else self.frames
        for n, i in enumerate(idxs):
            dst[n] = src[int(i)]

    def raw_frames_pinned(self):
        """See SyntheticVideo.raw_frames_pinned: offered when `frames` is a pinned uint8 torch tensor."""
        f = self.frames
        if isinstance(f, torch.Tensor) and f.dtype == torch.uint8 and f.is_pinned() and f.is_contiguous():
            return f, int
        return None

    def get_batch(self, idxs: Sequence[int]) -> torch.Tensor:
        import cv2
        out = np.empty((len(idxs), self.size, self.size, 3), dtype=np.uint8)
        for n, i in enumerate(idxs):
            f = np.asarray(self.frames[int(i)])
            if f.shape[0] == self.size and f.shape[1] == self.size:
                out[n] = f
            else:
                cv2.resize(f, (self.size, self.size), dst=out[n], interpolation=cv2.INTER_LINEAR)
        return torch.from_numpy(out)

    def get_batch_into(self, idxs: Sequence[int], out: torch.Tensor) -> None:
        out.copy_(self.get_batch(idxs))


# ----------------------------------------------------------------------------- the runner
@dataclass
class WindowPrediction:
    t0: int
    t1: int
    scores: np.ndarray       # [num_classes] float32


class SlidingWindowRunner:
    """Runs the replicated classifier over one rank's share of the windows of each video and gathers scores.

    `model` is anything with the reference inference contract `model([clip]) -> [B, classes]` on CUDA clips
    (aicity_action_b200.mvit.MViT); `preprocess(frames_u8_device) -> clip` defaults to the on-device normalisation.
    With `device=None` the runner only does the host-side logic (used by the gloo CPU tests with a stub model)."""

    def __init__(self, model: Callable, num_frames: int = 16, sampling_rate: int = 4, proposal_stride: int = 16,
                 batch_size: int = 8, dtype: torch.dtype = torch.bfloat16, device: Optional[torch.device] = None,
                 rank: int = 0, world: int = 1, group=None, preprocess: Optional[Callable] = None,
                 use_cuda_graph: bool = False, host_threads: Optional[int] = None, n_stage: Optional[int] = None,
                 device_resize=None, direct_upload: Optional[bool] = None):
        self.model, self.T, self.rate = model, num_frames, sampling_rate
        self.length, self.stride = num_frames * sampling_rate, proposal_stride   # run_action...py:76
        self.batch_size, self.dtype, self.device = batch_size, dtype, device
        self.rank, self.world, self.group = rank, world, group
        self.use_cuda_graph = use_cuda_graph and device is not None
        self._graphed = [None, None]
        self._dbuf = [None, None]
        self._copy_stream = None
        # Host side of the pipeline: worker threads that gather the frames of the next batches into pinned staging.  A batch
        # of eight 540p windows is ~190 MB of frame copies (~35 ms on one core) against ~12 ms of GPU time, so the default takes
        # the cores this rank can expect (cores / ranks on the box, one left for the launching thread), between 3 and 8.
        if host_threads is None:
            host_threads = min(8, max(3, (os.cpu_count() or 8) // max(1, world) - 1))
        if n_stage is None:
            n_stage = host_threads + 1
        self.host_threads, self.n_stage = max(1, host_threads), max(2, n_stage)
        self._stage, self._stage_free = [None], None
        # None: resize on the device whenever the video offers raw frames whose size differs from the model's; True / False force
        self.device_resize = device_resize
        # None: upload frames straight from a video's pinned frame store when every video of a call offers one
        # (`raw_frames_pinned`); False: always through get_raw_into + the staging ring
        self.direct_upload = direct_upload
        self.h2d_bytes = 0                    # bytes uploaded so far (frames + index lists), for the bench's accounting
        if preprocess is None and device is not None and dtype != torch.bfloat16:
            from . import ops
            preprocess = lambda u8: ops.preprocess_u8(u8, dtype)
        # bf16 (the default): the model takes the uint8 frames as they are and normalises inside its patch-embed fold, in
        # the eager and in the graph-replay path alike, so both give identical bits
        self.preprocess = preprocess

    def windows_for(self, video) -> List[Tuple[int, int]]:
        length, stride = fps_adjusted_window(self.length, self.stride, getattr(video, "fps", 30.0))
        return window_list(len(video), length, stride)

    def local_scores(self, video, windows: List[Tuple[int, int]]) -> Tuple[List[int], torch.Tensor]:
        """Scores of this rank's windows of one video: (window ids, [n_local, classes] float32 on CPU)."""
        return self.local_scores_multi([video], [windows])[0]

    @torch.no_grad()
    def local_scores_multi(self, videos, windows_list) -> List[Tuple[List[int], torch.Tensor]]:
        """Scores of this rank's windows of several videos (e.g. the three synchronised views of one recording) as ONE
        pipelined stream of batches: per video (window ids, [n_local, classes] float32 on CPU).

        On a device the batches are pipelined: host threads gather the uint8 frames of the next batches into pinned staging
        while batch i runs, uploads go through a copy stream into one of two device buffers, and the per-batch probabilities
        stay on the device until every video is done (one synchronisation per call; the pipeline does not drain between
        videos).  A ragged last batch of a video is padded by repeating its last window so that it, too, replays the captured
        graph; the padded rows are dropped."""
        mines = [shard_windows(len(w), self.rank, self.world) for w in windows_list]
        jobs = [(v, mine[b0:b0 + self.batch_size]) for v, mine in enumerate(mines)
                for b0 in range(0, len(mine), self.batch_size)]

        if self.device is None:                          # host-only logic (CPU tests with a stub model)
            outs = [[] for _ in videos]
            for v, ids in jobs:
                video, windows = videos[v], windows_list[v]
                frames = torch.stack([video.get_batch(frame_indices(*windows[w], self.T, len(video))) for w in ids])
                clip = self.preprocess(frames) if self.preprocess is not None else frames
                outs[v].append(self.model([clip]).float().cpu())
            return [(mine, torch.cat(o) if o else torch.zeros((0, 0))) for mine, o in zip(mines, outs)]
        if not jobs:
            return [(mine, torch.zeros((0, 0))) for mine in mines]

        from collections import deque
        from concurrent.futures import ThreadPoolExecutor
        from . import ops
        cur = torch.cuda.current_stream(self.device)
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        copy = self._copy_stream
        # Raw-frame path (N1): a video that exposes its decoded frames at native resolution (`raw_hw`, `get_raw_into`) is
        # resized ON THE DEVICE: every frame a batch needs travels once (duplicates removed), and mvit_resize_gather_u8
        # gathers the windows' frames by index and applies OpenCV's uint8 INTER_LINEAR arithmetic bit for bit
        # (scripts/utils.py:207-211) straight into the clip buffer.  Otherwise frames arrive at the model resolution.
        v0 = videos[0]
        raw_hw = getattr(v0, "raw_hw", None)
        size = int(getattr(v0, "size", 0)) or int(v0.get_batch([0]).shape[1])
        for v in videos[1:]:
            if getattr(v, "raw_hw", None) != raw_hw or (int(getattr(v, "size", 0)) or size) != size:
                raise ValueError("local_scores_multi: the videos of one call must share resolution (run them one by one)")
        dev_resize = (self.device_resize is not False and raw_hw is not None and hasattr(v0, "get_raw_into")
                      and (tuple(raw_hw) != (size, size) or self.device_resize is True))
        stores = [v.raw_frames_pinned() if hasattr(v, "raw_frames_pinned") else None for v in videos] \
            if (dev_resize and self.direct_upload is not False) else [None]
        direct = all(st is not None for st in stores)
        shape = (self.batch_size, self.T, size, size, 3)
        nf_max = self.batch_size * self.T
        up_shape = (nf_max,) + tuple(raw_hw) + (3,) if dev_resize else shape
        # Device side: two upload buffers that live as long as the runner (the captured graphs read them by address, and a
        # buffer allocated mid-stream could be a block that kernels in flight on the compute stream still use).  Host side:
        # a ring of pinned staging buffers that the worker threads fill in place: no per-batch cudaHostAlloc, one memcpy
        # per frame.
        if self._dbuf[0] is None or tuple(self._dbuf[0].shape) != shape or tuple(self._stage[0].shape) != up_shape:
            cur.synchronize()
            if self._dbuf[0] is None or tuple(self._dbuf[0].shape) != shape:
                self._dbuf = [torch.empty(shape, dtype=torch.uint8, device=self.device) for _ in range(2)]
                self._graphed = [None, None]
            self._stage = [torch.empty(up_shape, dtype=torch.uint8).pin_memory() for _ in range(self.n_stage)]
            self._stage_idx = [torch.zeros((nf_max,), dtype=torch.int32).pin_memory() for _ in range(self.n_stage)]
            self._stage_free = [torch.cuda.Event() for _ in range(self.n_stage)]
            self._draw = ([torch.empty(up_shape, dtype=torch.uint8, device=self.device) for _ in range(2)]
                          if dev_resize else [None, None])
            self._didx = [torch.zeros((nf_max,), dtype=torch.int32, device=self.device) for _ in range(2)]
        dbuf, stage, stage_free = self._dbuf, self._stage, self._stage_free
        stage_idx, draw, didx = self._stage_idx, self._draw, self._didx
        if self.use_cuda_graph:
            # capture BEFORE any worker thread exists: CUDA calls from other threads during a global-mode stream capture
            # (cudaHostAlloc, event queries) can invalidate it
            from .graphed import GraphedForward
            for slot in range(2):
                if self._graphed[slot] is None:
                    pool = next((g.pool for g in self._graphed if g is not None), None)
                    dbuf[slot].zero_()
                    self._graphed[slot] = GraphedForward(self.model, dbuf[slot], pool=pool)
        pad = self.use_cuda_graph                        # ragged batches are padded to the captured batch size

        def fill(j, v, ids):
            """Worker thread: gather the frames of job j into staging buffer j % n_stage once its last upload is done.
            Returns (pinned frames view, clips to run, clips to keep, number of distinct raw frames or None)."""
            video, windows = videos[v], windows_list[v]
            k = j % self.n_stage
            stage_free[k].synchronize()
            run_ids = ids + [ids[-1]] * (self.batch_size - len(ids)) if pad else ids
            if dev_resize:
                flat = [f for w in run_ids for f in frame_indices(*windows[w], self.T, len(video))]
                uniq = sorted(set(flat))
                pos = {f: n for n, f in enumerate(uniq)}
                stage_idx[k][:len(flat)] = torch.tensor([pos[f] for f in flat], dtype=torch.int32)
                if direct:
                    # no host copy: (start row, run length) of the frame store per upload, consecutive rows coalesced
                    row_of = stores[v][1]
                    return (v, coalesce_rows([row_of(f) for f in uniq])), len(run_ids), len(ids), len(uniq)
                video.get_raw_into(uniq, stage[k][:len(uniq)])
                return stage[k][:len(uniq)], len(run_ids), len(ids), len(uniq)
            buf = stage[k][:len(run_ids)]
            into = getattr(video, "get_batch_into", None)
            for n, w in enumerate(ids):
                idxs = frame_indices(*windows[w], self.T, len(video))
                if into is not None:
                    into(idxs, buf[n])
                else:
                    buf[n].copy_(torch.as_tensor(video.get_batch(idxs)))
            for n in range(len(ids), len(run_ids)):
                buf[n].copy_(buf[len(ids) - 1])
            return buf, len(run_ids), len(ids), None

        workers = ThreadPoolExecutor(max_workers=self.host_threads)
        pending, todo = deque(), iter(enumerate(jobs))

        def submit_next():
            nxt = next(todo, None)
            if nxt is not None:
                pending.append(workers.submit(fill, nxt[0], *nxt[1]))

        for ev in stage_free:
            ev.record(cur)
        for _ in range(self.n_stage - 1):
            submit_next()
        ready = [torch.cuda.Event(), torch.cuda.Event()]
        freed = [torch.cuda.Event(), torch.cuda.Event()]
        for ev in freed:
            ev.record(cur)
        outs, i = [[] for _ in videos], 0
        while pending:
            frames, n_clips, n_keep, n_uniq = pending.popleft().result()
            slot, k = i % 2, i % self.n_stage
            dev_frames = dbuf[slot][:n_clips]
            with torch.cuda.stream(copy):
                copy.wait_event(freed[slot])
                if dev_resize:
                    if direct:
                        store, n0 = stores[frames[0]][0], 0
                        for r, ln in frames[1]:
                            draw[slot][n0:n0 + ln].copy_(store[r:r + ln], non_blocking=True)
                            n0 += ln
                    else:
                        draw[slot][:n_uniq].copy_(frames, non_blocking=True)
                    didx[slot][:n_clips * self.T].copy_(stage_idx[k][:n_clips * self.T], non_blocking=True)
                else:
                    dev_frames.copy_(frames, non_blocking=True)
                ready[slot].record(copy)
                stage_free[k].record(copy)
            self.h2d_bytes += (n_uniq * draw[slot][0].numel() if (dev_resize and direct) else frames.numel()) \
                + (4 * n_clips * self.T if dev_resize else 0)
            submit_next()
            cur.wait_event(ready[slot])
            if dev_resize:
                ops.resize_gather_u8(draw[slot][:n_uniq], didx[slot][:n_clips * self.T], (size, size),
                                     out=dev_frames.view(n_clips * self.T, size, size, 3))
            if self.use_cuda_graph and n_clips == self.batch_size:
                # a captured graph per upload buffer (uint8 frames in, normalisation fused into the patch embed)
                probs = self._graphed[slot]()[:n_keep].clone()
            else:
                clip = self.preprocess(dev_frames) if self.preprocess is not None else dev_frames
                probs = self.model([clip])[:n_keep]
            freed[slot].record(cur)
            outs[jobs[i][0]].append(probs.float())
            i += 1
        workers.shutdown(wait=True)
        host = [torch.cat(o).cpu() if o else torch.zeros((0, 0)) for o in outs]    # the one synchronisation of the call
        ops.device_fault_check()
        return list(zip(mines, host))

    def gather(self, n_windows: int, mine: List[int], scores: torch.Tensor, num_classes: int) -> Optional[torch.Tensor]:
        """One all_gather per video of a padded [ceil(n/R), 1 + classes] block (window id, scores); returns the
        [n_windows, classes] table in window order on every rank."""
        if self.world == 1:
            return scores
        import torch.distributed as dist
        per = (n_windows + self.world - 1) // self.world
        block = torch.full((per, 1 + num_classes), -1.0, dtype=torch.float32)
        if len(mine):
            block[:len(mine), 0] = torch.tensor(mine, dtype=torch.float32)
            block[:len(mine), 1:] = scores
        dev = self.device if (self.device is not None and dist.get_backend(self.group) == "nccl") else torch.device("cpu")
        block = block.to(dev)
        out = [torch.empty_like(block) for _ in range(self.world)]
        dist.all_gather(out, block, group=self.group)
        table = torch.zeros((n_windows, num_classes), dtype=torch.float32)
        for blk in out:
            blk = blk.cpu()
            ok = blk[:, 0] >= 0
            table[blk[ok, 0].long()] = blk[ok, 1:]
        return table

    def run_video(self, video, num_classes: int) -> List[Tuple[int, int, np.ndarray]]:
        """The reference's per-video result: [(t0, t1, scores[classes])] sorted by t0 (run_action...py:110-125)."""
        windows = self.windows_for(video)
        mine, scores = self.local_scores(video, windows)
        table = self.gather(len(windows), mine, scores, num_classes)
        preds = [(int(t0), int(t1), table[w].numpy()) for w, (t0, t1) in enumerate(windows)]
        preds.sort(key=lambda x: x[0])
        return preds

    def run_videos(self, videos, num_classes: int) -> List[List[Tuple[int, int, np.ndarray]]]:
        """`run_video` for several videos of the same resolution (the 3 views of a recording): their batches form one
        pipelined stream, then ONE all_gather per video exchanges the scores."""
        windows_list = [self.windows_for(v) for v in videos]
        local = self.local_scores_multi(videos, windows_list)
        results = []
        for windows, (mine, scores) in zip(windows_list, local):
            table = self.gather(len(windows), mine, scores, num_classes)
            preds = [(int(t0), int(t1), table[w].numpy()) for w, (t0, t1) in enumerate(windows)]
            preds.sort(key=lambda x: x[0])
            results.append(preds)
        return results

    def run_and_save(self, video, video_name: str, out_dir: str, num_classes: int):
        preds = self.run_video(video, num_classes)
        if self.rank == 0:
            os.makedirs(out_dir, exist_ok=True)
            with open(os.path.join(out_dir, f"{video_name}.pkl"), "wb") as f:   # run_action...py:128-130
                pickle.dump(preds, f)
        return preds
