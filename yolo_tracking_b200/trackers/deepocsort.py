"""DeepOCSORT with the reference's constructor and update() contract
(boxmot/trackers/deepocsort/deep_ocsort.py:308-520), backed by a one-stream context of the fused DeepOCSORT frame step
(csrc/deepocsort_step.cu): camera correction, predict, association with the adaptive appearance weight, the recovery
round, the Kalman update with the observation-centric re-update, the embedding blend, births and the output scan are ONE
kernel launch per frame; this class only hands the frame over.  For throughput track many streams with one
`BatchedTracker("deepocsort", ...)` instead.

Only the default `new_kf` filter is built (new_kf_off=True raises).  The ReID network and the camera-motion estimator are
out of scope (BASELINE.json): embeddings come from a `model` with get_features(xyxys, img) or from update(..., feats=...),
the warp from update(..., warp=...) (None = identity).
"""
from __future__ import annotations

import numpy as np

from .. import _lib
from ..batch import BatchedTracker
from .bytetrack import _SingleStreamTracker, _device_index


class _TrackView:
    """Read-only record of one live tracker (what tests and callers look at in `tracker.trackers`)."""
    __slots__ = ("id", "age", "hits", "hit_streak", "time_since_update", "conf", "cls", "det_ind")

    def __init__(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)


class DeepOCSort:
    def __init__(self, model_weights=None, device=0, fp16=False, per_class=True, det_thresh=0.3, max_age=30, min_hits=3,
                 iou_threshold=0.3, delta_t=3, asso_func="iou", inertia=0.2, w_association_emb=0.5, alpha_fixed_emb=0.95,
                 aw_param=0.5, embedding_off=False, cmc_off=False, aw_off=False, new_kf_off=False, model=None,
                 max_tracks=256, max_dets=256, **kwargs):
        if new_kf_off:
            raise NotImplementedError("only DeepOCSORT's default filter (new_kf) is built")
        if asso_func not in _lib.SIM:
            raise ValueError("Invalid function specified. Must be either '(g,d,c, )iou_batch' or 'centroid_batch'.")
        self.device = _device_index(device)
        self.max_age, self.min_hits, self.iou_threshold, self.det_thresh = max_age, min_hits, iou_threshold, det_thresh
        self.delta_t, self.asso_func, self.inertia = delta_t, asso_func, inertia
        self.w_association_emb, self.alpha_fixed_emb, self.aw_param = w_association_emb, alpha_fixed_emb, aw_param
        self.per_class, self.embedding_off, self.cmc_off, self.aw_off = per_class, embedding_off, cmc_off, aw_off
        self.model = model
        self.frame_count = 0
        self._max_tracks, self._max_dets = max_tracks, max_dets
        self._batch = None                                   # created on the first frame: the reference's attributes may be
        self._feat_dim = None                                # changed after construction (its own tests do), and the embedding
        self._pending_empty = 0
        _lib.load()                                          # size is only known then; fail loudly without the library

    # ------------------------------------------------------------------ the device context
    def _context(self, feat_dim):
        if self._batch is None:
            if self.asso_func not in _lib.SIM:
                raise ValueError("Invalid function specified. Must be either '(g,d,c, )iou_batch' or 'centroid_batch'.")
            self._feat_dim = 0 if self.embedding_off else int(feat_dim)
            pad = (-self._feat_dim) % 4                      # rows are padded with zeros to a multiple of 4 (dot products unchanged)
            self._feat_pad = pad
            self._batch = BatchedTracker("deepocsort", 1, max_tracks=self._max_tracks, max_dets=self._max_dets, device=self.device,
                                         feat_dim=self._feat_dim + pad, det_thresh=self.det_thresh, max_age=self.max_age,
                                         min_hits=self.min_hits, iou_threshold=self.iou_threshold, delta_t=self.delta_t,
                                         asso_func=self.asso_func, inertia=self.inertia, w_association_emb=self.w_association_emb,
                                         alpha_fixed_emb=self.alpha_fixed_emb, aw_param=self.aw_param,
                                         embedding_off=self.embedding_off, aw_off=self.aw_off)
            for _ in range(self._pending_empty):
                self._batch.update_frames([np.zeros((0, 6))], feats=None if self.embedding_off else [np.zeros((0, self._batch.feat_dim), dtype=np.float32)])
            self._pending_empty = 0
        return self._batch

    def update(self, dets, img, feats=None, warp=None):
        """`feats`: what self.model.get_features(dets[conf > det_thresh, :4], img) would return (deep_ocsort.py:382-390);
        `warp`: this frame's externally estimated 2x3 camera motion (self.cmc.apply in the reference, :393-396)."""
        _SingleStreamTracker._check(dets)
        assert isinstance(img, (np.ndarray, tuple)), f"Unsupported 'img' input type '{type(img)}', valid format is np.ndarray"
        self.frame_count += 1
        h, w = img.shape[:2] if isinstance(img, np.ndarray) else img
        dets = dets if dets.dtype == np.float32 else np.asarray(dets, dtype=np.float64)
        n = len(dets)
        if n > self._max_dets:
            raise ValueError(f"{n} detections exceed max_dets={self._max_dets}")
        keep = np.nonzero(dets[:, 4] > self.det_thresh)[0]
        rows = None
        if not self.embedding_off:
            if feats is not None:
                high = np.asarray(feats, dtype=np.float32)
                assert len(high) == len(keep), "feats must hold one row per detection with conf > det_thresh"
            elif len(keep):
                if self.model is None:
                    raise ValueError("DeepOCSort needs `feats` or a `model` with get_features(xyxys, img)")
                high = np.asarray(self.model.get_features(dets[keep, 0:4], img), dtype=np.float32)
            else:
                high = np.zeros((0, self._feat_dim or 4), dtype=np.float32)
            if self._batch is None and high.shape[0] == 0:
                # nothing to learn the embedding size from yet and nothing to track either: the frame is replayed (it
                # only advances the frame counter) once the context exists
                self._pending_empty += 1
                return np.array([])
            trk = self._context(high.shape[1])
            rows = np.zeros((n, trk.feat_dim), dtype=np.float32)       # one row per detection; only the kept ones are read
            if len(keep):
                rows[keep, :high.shape[1]] = high
        else:
            trk = self._context(0)
        out = trk.update_frames([dets], feats=None if rows is None else [rows],
                                warps=None if (warp is None or self.cmc_off) else np.asarray(warp, dtype=np.float64).reshape(1, 6),
                                img_hw=(h, w), dtype=dets.dtype)[0]
        return out if len(out) else np.array([])

    # ------------------------------------------------------------------ probes
    @property
    def stats(self):
        return self._batch.counters() if self._batch is not None else dict(lap_frames=0, ocr_frames=0, oru=0)

    def state(self):
        if self._batch is None:
            z = np.zeros(0, dtype=np.int32)
            return dict(track_id=z, age=z, time_since_update=z, hits=z, hit_streak=z, observed=z, frozen=z, x=np.zeros((0, 8)),
                        P=np.zeros((0, 8, 8)), velocity=np.zeros((0, 2)), last_observation=np.zeros((0, 5)), emb=np.zeros((0, 0)))
        st = self._batch.state(0)
        if not self.embedding_off and self._feat_pad:
            st["emb"] = st["emb"][:, :self._feat_dim]
        return st

    @property
    def trackers(self):
        st = self.state()
        return [_TrackView(id=int(st["track_id"][i]), age=int(st["age"][i]), hits=int(st["hits"][i]),
                           hit_streak=int(st["hit_streak"][i]), time_since_update=int(st["time_since_update"][i]),
                           conf=float(st["conf"][i]) if "conf" in st else 0.0, cls=float(st["cls"][i]) if "cls" in st else 0.0,
                           det_ind=int(st["det_ind"][i]) if "det_ind" in st else -1) for i in range(len(st["track_id"]))]
