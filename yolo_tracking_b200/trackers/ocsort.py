"""OCSort with the reference's constructor and update() contract
(boxmot/trackers/ocsort/ocsort.py:190-379), backed by a one-stream device context."""
from __future__ import annotations

import numpy as np

from .bytetrack import _SingleStreamTracker


class OCSort(_SingleStreamTracker):
    kind = "ocsort"

    def __init__(self, per_class=True, det_thresh=0.2, max_age=30, min_hits=3, asso_threshold=0.3, delta_t=3,
                 asso_func="iou", inertia=0.2, use_byte=False, device=0, max_tracks=256, max_dets=256):
        self.per_class = per_class          # accepted and unused, like the reference (ocsort.py:193)
        self.max_age, self.min_hits, self.asso_threshold = max_age, min_hits, asso_threshold
        self.det_thresh, self.delta_t, self.inertia, self.use_byte = det_thresh, delta_t, inertia, use_byte
        self.frame_count = 0
        self._make(device, max_tracks, max_dets, det_thresh=det_thresh, max_age=max_age, min_hits=min_hits,
                   asso_threshold=asso_threshold, delta_t=delta_t, asso_func=asso_func, inertia=inertia,
                   use_byte=use_byte)

    def update(self, dets, img):
        self._check(dets)
        h, w = img.shape[0:2] if hasattr(img, "shape") else img       # only the frame size is used (ocsort.py:239)
        rows = self._step(self._as_rows(dets), img_hw=(h, w))
        self.frame_count += 1
        return rows if len(rows) else np.array([])                   # ocsort.py:379
