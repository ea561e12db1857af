"""BYTETracker with the reference's constructor and update() contract
(boxmot/trackers/bytetrack/byte_tracker.py:114-281), backed by a one-stream device context.
For throughput track many streams with one `BatchedTracker` instead."""
from __future__ import annotations

import numpy as np

from ..batch import BatchedTracker


def _device_index(device):
    if device is None:
        return 0
    if isinstance(device, int):
        return device
    s = str(device).lower().replace("cuda:", "").strip()
    if s in ("", "cuda"):
        return 0
    if s == "cpu":
        raise RuntimeError("yolo_tracking_b200 runs on CUDA devices only (no CPU fallback)")
    return int(s.split(",")[0])


class _SingleStreamTracker:
    kind = ""

    def _make(self, device, max_tracks, max_dets, feat_dim=0, **params):
        self._max_dets = max_dets
        self._batch = BatchedTracker(self.kind, 1, max_tracks=max_tracks, max_dets=max_dets,
                                     device=_device_index(device), feat_dim=feat_dim, **params)
        self.frame_id = 0

    @staticmethod
    def _check(dets):
        assert isinstance(dets, np.ndarray), f"Unsupported 'dets' input format '{type(dets)}', valid format is np.ndarray"
        assert len(dets.shape) == 2, "Unsupported 'dets' dimensions, valid number of dimensions is two"
        assert dets.shape[1] == 6, "Unsupported 'dets' 2nd dimension lenght, valid lenghts is 6"

    @staticmethod
    def _as_rows(dets):
        """float32 detections travel as float32 (widened on the device, exact); everything else as float64."""
        return dets if dets.dtype == np.float32 else np.asarray(dets, dtype=np.float64)

    def _step(self, dets, img_hw=(0, 0), feats=None, warp=None):
        """One frame through the packed host interface: one copy in, the step, one copy of the compact rows out; a
        capacity overflow of this very step raises here."""
        n = len(dets)
        if n > self._max_dets:
            raise ValueError(f"{n} detections exceed max_dets={self._max_dets}")
        rows = self._batch.update_frames([dets], feats=None if feats is None else [feats],
                                         warps=None if warp is None else np.asarray(warp, dtype=np.float64).reshape(1, 6),
                                         img_hw=img_hw, dtype=dets.dtype)[0]
        self.frame_id += 1
        return rows

    def state(self):
        return self._batch.state(0)


class BYTETracker(_SingleStreamTracker):
    kind = "bytetrack"

    def __init__(self, track_thresh=0.45, match_thresh=0.8, track_buffer=25, frame_rate=30,
                 device=0, max_tracks=256, max_dets=256):
        self.track_thresh = track_thresh
        self.match_thresh = match_thresh
        self.track_buffer = track_buffer
        self.det_thresh = track_thresh
        self.buffer_size = int(frame_rate / 30.0 * track_buffer)
        self.max_time_lost = self.buffer_size
        self._make(device, max_tracks, max_dets, track_thresh=track_thresh, match_thresh=match_thresh,
                   track_buffer=track_buffer, frame_rate=frame_rate)

    def update(self, dets, _=None):
        self._check(dets)
        rows = self._step(self._as_rows(dets))
        return rows if len(rows) else np.asarray([])        # byte_tracker.py:280: empty -> shape (0,)
