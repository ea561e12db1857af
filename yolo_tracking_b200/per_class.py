"""Per-class tracking (the reference's PerClassDecorator, boxmot/utils/__init__.py:22-61, feeds one
class of detections at a time so that classes are never mixed).  Here the class is an extra stream axis: one device
context holds `n_classes` independent streams and ONE kernel launch per frame advances all of them, instead of one
update() per class.

Differences from the decorator, which shares a single tracker object between the classes: every class has its own
track lists and id counter (ids are made unique as (id - 1) * n_classes + cls + 1), classes are always stepped (an
absent class sees an empty frame, exactly what a per-class tracker object would see), output rows are ordered by class.
"""
from __future__ import annotations

import numpy as np

from .batch import BatchedTracker
from .trackers.bytetrack import _SingleStreamTracker, _device_index


class PerClassTracker:
    def __init__(self, kind, n_classes=80, device=0, max_tracks=128, max_dets=128, **params):
        self.kind, self.n_classes, self._max_dets = kind, int(n_classes), max_dets
        self._batch = BatchedTracker(kind, self.n_classes, max_tracks=max_tracks, max_dets=max_dets,
                                     device=_device_index(device), **params)
        self._dets = np.zeros((self.n_classes, max_dets, 6), dtype=np.float64)
        self._nd = np.zeros((self.n_classes,), dtype=np.int32)
        self._src = np.zeros((self.n_classes, max_dets), dtype=np.int64)
        self.frame_id = 0

    def update(self, dets, img=None):
        _SingleStreamTracker._check(dets)
        dets = np.asarray(dets, dtype=np.float64)
        cls = dets[:, 5].astype(np.int64)
        if len(dets) and (cls.min() < 0 or cls.max() >= self.n_classes or np.any(cls != dets[:, 5])):
            raise ValueError(f"class ids must be integers in [0, {self.n_classes})")
        self._nd[:] = 0
        for c in np.unique(cls):
            idx = np.nonzero(cls == c)[0]
            if len(idx) > self._max_dets:
                raise ValueError(f"{len(idx)} detections of class {c} exceed max_dets={self._max_dets}")
            self._dets[c, :len(idx)] = dets[idx]
            self._src[c, :len(idx)] = idx
            self._nd[c] = len(idx)
        hw = tuple(img.shape[:2]) if hasattr(img, "shape") else (img if isinstance(img, tuple) else (0, 0))
        out, nout = self._batch.update_batch(self._dets, self._nd, img_hw=hw)
        self._batch.sync()
        self.frame_id += 1
        rows = []
        for c in np.nonzero(nout)[0]:
            r = out[c, :nout[c]].copy()
            r[:, 4] = (r[:, 4] - 1) * self.n_classes + c + 1
            r[:, 7] = self._src[c, r[:, 7].astype(np.int64)]          # det_ind refers to the caller's rows
            rows.append(r)
        return np.concatenate(rows, axis=0) if rows else np.empty((0, 8))

    def close(self):
        self._batch.close()
