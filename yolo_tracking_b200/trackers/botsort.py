"""BoTSORT with the reference's constructor and update() contract
(boxmot/trackers/botsort/bot_sort.py:184-420), backed by a one-stream device context.

Out of scope by BASELINE.json: the ReID network and the camera-motion estimator.  The
appearance seam of the reference (`self.model.get_features(xyxys, img)`, bot_sort.py:266) is
kept: pass any object with that method as `model`, or hand the per-detection embeddings to
`update(dets, img, feats=...)` directly.  Camera motion is the identity warp (the reference
ignores `cmc_method` as well and always builds SparseOptFlow, bot_sort.py:228)."""
from __future__ import annotations

import numpy as np

from .bytetrack import _SingleStreamTracker


class BoTSORT(_SingleStreamTracker):
    kind = "botsort"

    def __init__(self, model_weights=None, device=0, fp16=False, track_high_thresh=0.5, track_low_thresh=0.1,
                 new_track_thresh=0.6, track_buffer=30, match_thresh=0.8, proximity_thresh=0.5,
                 appearance_thresh=0.25, cmc_method="sparseOptFlow", frame_rate=30, fuse_first_associate=False,
                 with_reid=True, model=None, feat_dim=512, max_tracks=256, max_dets=256, camera_motion=False):
        self.fuse_first_associate = fuse_first_associate
        self.track_high_thresh, self.track_low_thresh, self.new_track_thresh = track_high_thresh, track_low_thresh, new_track_thresh
        self.match_thresh, self.proximity_thresh, self.appearance_thresh = match_thresh, proximity_thresh, appearance_thresh
        self.buffer_size = int(frame_rate / 30.0 * track_buffer)
        self.max_time_lost = self.buffer_size
        self.with_reid, self.model, self.feat_dim = with_reid, model, feat_dim if with_reid else 0
        self.camera_motion = bool(camera_motion)
        self._make(device, max_tracks, max_dets, feat_dim=self.feat_dim, track_high_thresh=track_high_thresh,
                   track_low_thresh=track_low_thresh, new_track_thresh=new_track_thresh, track_buffer=track_buffer,
                   match_thresh=match_thresh, proximity_thresh=proximity_thresh, appearance_thresh=appearance_thresh,
                   frame_rate=frame_rate, with_reid=with_reid, fuse_first_associate=fuse_first_associate,
                   camera_motion=camera_motion)

    def update(self, dets, img, feats=None, warp=None):
        """`feats`: [len(dets), feat_dim] embeddings per detection row as the ReID seam returns them; when
        omitted, `self.model.get_features(dets_first[:, :4], img)` is called for the first-round rows like the
        reference does.  `warp`: the frame's 2x3 camera-motion matrix (what `self.cmc.apply(img, dets_first)` returns in
        the reference, bot_sort.py:293-295); needs `camera_motion=True` at construction, None = identity."""
        self._check(dets)
        dets = self._as_rows(dets)
        n = len(dets)
        if n > self._max_dets:
            raise ValueError(f"{n} detections exceed max_dets={self._max_dets}")
        if warp is not None and not self.camera_motion:
            raise ValueError("construct BoTSORT(camera_motion=True) to apply camera-motion warps")
        rows_feats = None
        if self.with_reid:
            rows_feats = np.zeros((n, self.feat_dim), dtype=np.float32)
            if feats is not None:
                rows_feats[:] = np.asarray(feats, dtype=np.float32)
            else:
                first = np.nonzero(dets[:, 4] > self.track_high_thresh)[0]
                if len(first):
                    if self.model is None:
                        raise ValueError("with_reid=True needs `feats` or a `model` with get_features(xyxys, img)")
                    rows_feats[first] = np.asarray(self.model.get_features(dets[first, 0:4], img), dtype=np.float32)
        rows = self._step(dets, feats=rows_feats, warp=warp)
        return rows if len(rows) else np.asarray([])        # bot_sort.py:419: empty -> shape (0,)
