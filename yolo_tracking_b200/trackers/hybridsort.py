"""HybridSORT with the reference's constructor and update() contract
(boxmot/trackers/hybridsort/hybridsort.py:336-570), backed by a one-stream context of the fused HybridSORT frame step
(csrc/hybridsort_step.cu): predict, the four-corner association with the appearance term, the long-term-ReID correction,
the recovery round, the Kalman update with the observation-centric re-update, the float32 embedding blend, births and
the output scan are ONE kernel launch per frame; this class only hands the frame over.  For throughput track many
streams with one `BatchedTracker("hybridsort", ...)` instead.

`update` is wrapped by the reference's PerClassDecorator (boxmot/utils/__init__.py:22-61) and HybridSORT sets
`per_class = True` itself (hybridsort.py:346): with detections of several classes the undecorated update runs once per
class that has detections or live trackers - every call a full frame for ALL trackers.  That wrapper is restated here
around the device step, with the same set / dict expressions so that the classes come in the same order.

`use_byte=True` raises: tracker_zoo.py:100-115 never forwards it and the reference's branch passes an embedding row where
the class goes (hybridsort.py:470-474).  The ReID network is out of scope (BASELINE.json): embeddings come from a `model`
with get_features(xyxys, img) or from update(..., feats=...) - one row per detection, what get_features returns for all
boxes of the call (hybridsort.py:394).  ECC is off in the reference (:363).
"""
from __future__ import annotations

import numpy as np

from .. import _lib
from ..batch import BatchedTracker
from .bytetrack import _SingleStreamTracker, _device_index
from .deepocsort import _TrackView


class HybridSORT:
    def __init__(self, reid_weights=None, device=0, half=False, det_thresh=0.0, max_age=30, min_hits=3, iou_threshold=0.3,
                 delta_t=3, asso_func="iou", inertia=0.2, use_byte=False, model=None, max_tracks=256, max_dets=256, **kwargs):
        if use_byte:
            raise NotImplementedError("HybridSORT's use_byte branch cannot produce a result row in the reference (hybridsort.py:470-474)")
        if asso_func not in _lib.SIM:
            raise ValueError("Invalid function specified. Must be either '(g,d,c, )iou_batch' or 'centroid_batch'.")
        self.device = _device_index(device)
        self.max_age, self.min_hits, self.iou_threshold, self.det_thresh = max_age, min_hits, iou_threshold, det_thresh
        self.delta_t, self.asso_func, self.inertia, self.use_byte = delta_t, asso_func, inertia, use_byte
        self.per_class = True                                # hybridsort.py:346
        self.model = model
        self.frame_count = 0
        self._max_tracks, self._max_dets = max_tracks, max_dets
        self._batch = None                                   # created on the first frame that carries embeddings (their size
        self._pending_empty = 0                              # is only known then)
        _lib.load()                                          # fail loudly without the library

    def _context(self, feat_dim):
        if self._batch is None:
            self._feat_dim = int(feat_dim)
            pad = (-self._feat_dim) % 4                      # rows are padded with zeros to a multiple of 4 (norms / dot products unchanged)
            self._batch = BatchedTracker("hybridsort", 1, max_tracks=self._max_tracks, max_dets=self._max_dets, device=self.device,
                                         feat_dim=self._feat_dim + pad, det_thresh=self.det_thresh, max_age=self.max_age,
                                         min_hits=self.min_hits, iou_threshold=self.iou_threshold, delta_t=self.delta_t,
                                         asso_func=self.asso_func, inertia=self.inertia)
            D, Fp = self._max_dets, self._batch.feat_dim
            self._dets = np.zeros((1, D, 6))
            self._nd = np.zeros(1, dtype=np.int32)
            self._feats = np.zeros((1, D, Fp), dtype=np.float32)
            for _ in range(self._pending_empty):             # frames seen before the first embedding: they only count
                self._batch.update_batch(self._dets, self._nd, feats=self._feats)
            self._pending_empty = 0
        return self._batch

    # ------------------------------------------------------------------ hybridsort.py:373-570, one class
    def _update(self, dets, img, feats):
        self.frame_count += 1
        h, w = img.shape[:2] if isinstance(img, np.ndarray) else img
        n = len(dets)
        if n > self._max_dets:
            raise ValueError(f"{n} detections exceed max_dets={self._max_dets}")
        if feats is None and n:
            if self.model is None:
                raise ValueError("HybridSORT needs `feats` or a `model` with get_features(xyxys, img)")
            feats = self.model.get_features(dets[:, 0:4], img)
        if self._batch is None and n == 0:
            self._pending_empty += 1
            return np.empty((0, 7))
        feats = np.asarray(feats, dtype=np.float32).reshape(n, -1) if n else None
        trk = self._context(feats.shape[1] if n else self._feat_dim)
        if n and feats.shape[1] != self._feat_dim:
            raise ValueError(f"embedding size changed from {self._feat_dim} to {feats.shape[1]}")
        self._dets[0, :n] = dets
        self._nd[0] = n
        if n:
            self._feats[0, :n, :feats.shape[1]] = feats
        out, nout = trk.update_batch(self._dets, self._nd, feats=self._feats, img_hw=(h, w))
        m = int(nout[0])
        return out[0, :m].copy() if m else np.empty((0, 7))

    # ------------------------------------------------------------------ PerClassDecorator (boxmot/utils/__init__.py:22-61)
    def update(self, dets, img, feats=None):
        """`feats`: one embedding row per detection row (what self.model.get_features(dets[:, :4], img) returns)."""
        _SingleStreamTracker._check(dets)
        assert isinstance(img, (np.ndarray, tuple)), f"Unsupported 'img' input type '{type(img)}', valid format is np.ndarray"
        dets = np.asarray(dets, dtype=np.float64)
        if self.per_class is True and dets.size != 0:
            idx_of = {class_id: np.array([i for i, det in enumerate(dets) if det[5] == class_id], dtype=np.int64)
                      for class_id in set(det[5] for det in dets)}
            detected_classes = set(idx_of.keys())
            active_classes = set(np.float64(c) for c in (self._batch.live_classes(0) if self._batch is not None else ()))
            relevant_classes = active_classes.union(detected_classes)
            mc_dets = np.empty(shape=(0, 8))
            for class_id in relevant_classes:
                idx = idx_of.get(int(class_id), np.zeros(0, dtype=np.int64))
                out = self._update(dets[idx].reshape(-1, 6), img, None if feats is None else np.asarray(feats)[idx])
                if out.size != 0:
                    mc_dets = np.append(mc_dets, out, axis=0)
            return mc_dets
        return self._update(dets.reshape(-1, 6), img, feats)

    # ------------------------------------------------------------------ probes
    @property
    def stats(self):
        if self._batch is None:
            return dict(lap_frames=0, ocr_frames=0, oru=0, corrections=0)
        c = self._batch.counters()
        return dict(lap_frames=c["lap_frames"], ocr_frames=c["ocr_frames"], oru=c["oru"], corrections=c["gallery_rows"])

    def state(self):
        if self._batch is None:
            z = np.zeros(0, dtype=np.int32)
            return dict(n=0, track_id=z, age=z, time_since_update=z, hits=z, hit_streak=z, observed=z, x=np.zeros((0, 9)),
                        P=np.zeros((0, 9, 9)), velocity=np.zeros((0, 4, 2)), last_observation=np.zeros((0, 5)), conf=np.zeros(0),
                        cls=np.zeros(0), det_ind=np.zeros(0), smooth_feat=np.zeros((0, 0), dtype=np.float32))
        st = self._batch.state(0)
        st["smooth_feat"] = st["smooth_feat"][:, :self._feat_dim]
        return st

    @property
    def trackers(self):
        st = self.state()
        return [_TrackView(id=int(st["track_id"][i]), age=int(st["age"][i]), hits=int(st["hits"][i]),
                           hit_streak=int(st["hit_streak"][i]), time_since_update=int(st["time_since_update"][i]),
                           conf=float(st["conf"][i]), cls=float(st["cls"][i]), det_ind=int(st["det_ind"][i]))
                for i in range(len(st["track_id"]))]
