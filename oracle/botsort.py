"""ORACLE (test infrastructure): the reference's BoT-SORT frame step restated on numpy, with
the camera-motion warp fixed to identity and the ReID network replaced by caller-supplied
detection embeddings (BASELINE.json north_star).

Follows boxmot/trackers/botsort/bot_sort.py (reference file:line):
  STrack.__init__ :19-38, update_features :40-48 (fp32, in-place re-normalisation, the
  smooth/curr aliasing of a fresh detection), update_cls :50-67 (in-loop arg-max), multi_predict
  :76-92, activate :113-126, re_activate :128-143, update :145-169, xyxy :171-181,
  BoTSORT.__init__ :184-229, update :231-420, joint / sub / remove_duplicate :423-465
and boxmot/utils/matching.py embedding_distance :145-167 (fp32 cast, scipy cdist cosine in
double, clamp at 0), iou_distance :94-119, fuse_score :213-221, linear_assignment :56-71;
boxmot/motion/kalman_filters/botsort_kf.py (XYWH filter, oracle/kalman.py kind "xywh").

`feats[n_dets, F]` holds one embedding per detection ROW as the ReID seam returns it
(`get_features`, already divided by the Frobenius norm of the whole matrix); only the rows
of first-round detections are read, like bot_sort.py:266.

Parity pinned by tests/golden/botsort_*.npz, generated from the live reference.
"""
from __future__ import annotations

import numpy as np
from scipy.spatial.distance import cdist

from . import boxes, kalman
from .lap import assign_with_limit

NEW, TRACKED, LOST, LONG_LOST, REMOVED = 0, 1, 2, 3, 4     # botsort/basetrack.py:7-12
KIND = "xywh"


class _Trk:
    def __init__(self, xywh, score, cls, det_ind, feat):
        self.xywh = xywh
        self.score, self.cls, self.det_ind = score, cls, det_ind
        self.mean = self.cov = None
        self.state = NEW
        self.activated = False
        self.tid = 0
        self.frame_id = self.start_frame = 0
        self.tracklet_len = 0
        self.cls_hist = []
        self.vote(cls, score)
        self.smooth_feat = self.curr_feat = None
        if feat is not None:
            self.update_features(feat)

    def update_features(self, feat):                       # :40-48, float32 throughout
        feat /= np.linalg.norm(feat)
        self.curr_feat = feat
        if self.smooth_feat is None:
            self.smooth_feat = feat                        # alias: the next line renormalises feat too
        else:
            self.smooth_feat = 0.9 * self.smooth_feat + (1 - 0.9) * feat
        self.smooth_feat /= np.linalg.norm(self.smooth_feat)

    def vote(self, cls, score):                            # update_cls :50-67
        if self.cls_hist:
            best = 0
            found = False
            for c in self.cls_hist:
                if cls == c[0]:
                    c[1] += score
                    found = True
                if c[1] > best:
                    best = c[1]
                    self.cls = c[0]
            if not found:
                self.cls_hist.append([cls, score])
                self.cls = cls
        else:
            self.cls_hist.append([cls, score])
            self.cls = cls

    def box(self):                                         # :171-181
        return boxes.xywh_to_xyxy(self.xywh if self.mean is None else self.mean[:4])


def _stack_boxes(trks):
    return np.stack([t.box() for t in trks]) if trks else np.zeros((0, 4))


def _union(a, b):
    seen = {t.tid for t in a}
    out = list(a)
    for t in b:
        if t.tid not in seen:
            seen.add(t.tid)
            out.append(t)
    return out


def _iou_cost(trks, dets):                                 # matching.py:94-119
    if len(trks) == 0 or len(dets) == 0:
        return np.zeros((len(trks), len(dets)))
    return 1 - boxes.iou(_stack_boxes(trks), _stack_boxes(dets))


def _emb_cost(trks, dets):                                 # matching.py:145-167
    if len(trks) == 0 or len(dets) == 0:
        return np.zeros((len(trks), len(dets)), dtype=np.float32)
    df = np.asarray([d.curr_feat for d in dets], dtype=np.float32)
    tf = np.asarray([t.smooth_feat for t in trks], dtype=np.float32)
    return np.maximum(0.0, cdist(tf, df, "cosine"))


def _fuse(cost, dets):                                     # matching.py:213-221
    if cost.size == 0:
        return cost
    return 1 - (1 - cost) * np.array([d.score for d in dets])[None, :]


class BoTSORTOracle:
    def __init__(self, track_high_thresh=0.5, track_low_thresh=0.1, new_track_thresh=0.6, track_buffer=30,
                 match_thresh=0.8, proximity_thresh=0.5, appearance_thresh=0.25, frame_rate=30,
                 fuse_first_associate=False, with_reid=True):
        self.high, self.low, self.new_thresh = track_high_thresh, track_low_thresh, new_track_thresh
        self.match_thresh, self.prox, self.app = match_thresh, proximity_thresh, appearance_thresh
        self.max_time_lost = int(frame_rate / 30.0 * track_buffer)
        self.fuse_first, self.with_reid = fuse_first_associate, with_reid
        self.frame_id = 0
        self.tracked: list[_Trk] = []
        self.lost: list[_Trk] = []
        self.removed_ids: set[int] = set()
        self.next_id = 0
        self.track_updates = 0

    def _apply(self, trk, det):                            # update :145-169 / re_activate :128-143
        m, c = kalman.update(KIND, trk.mean, trk.cov, det.xywh)
        trk.mean, trk.cov = m[0], c[0]
        if det.curr_feat is not None:
            trk.update_features(det.curr_feat)
        trk.tracklet_len = trk.tracklet_len + 1 if trk.state == TRACKED else 0
        trk.state = TRACKED
        trk.activated = True
        trk.frame_id = self.frame_id
        trk.score, trk.cls, trk.det_ind = det.score, det.cls, det.det_ind
        trk.vote(det.cls, det.score)

    def _combined(self, trks, dets, fuse):
        iou_d = _iou_cost(trks, dets)
        mask = iou_d > self.prox
        if fuse:
            iou_d = _fuse(iou_d, dets)
        if not self.with_reid:
            return iou_d
        emb = _emb_cost(trks, dets) / 2.0
        emb[emb > self.app] = 1.0
        emb[mask] = 1.0
        return np.minimum(iou_d, emb)

    def update(self, dets, feats=None, warp=None):
        assert isinstance(dets, np.ndarray), "dets must be np.ndarray"
        assert dets.ndim == 2, "dets must be two-dimensional"
        assert dets.shape[1] == 6, "dets must have 6 columns"
        dets = np.asarray(dets, dtype=np.float64)
        self.frame_id += 1
        conf = dets[:, 4]
        lo = np.nonzero((conf > self.low) & (conf < self.high))[0]
        hi = np.nonzero(conf > self.high)[0]

        def mk(j, with_feat):
            f = None
            if with_feat and self.with_reid:
                f = np.array(feats[j], dtype=np.float32)       # a private copy: update_features works in place
            return _Trk(boxes.xyxy_to_xywh(dets[j, :4]), dets[j, 4], dets[j, 5], float(j), f)
        d1 = [mk(j, True) for j in hi]

        unconfirmed = [t for t in self.tracked if not t.activated]
        confirmed = [t for t in self.tracked if t.activated]
        pool = _union(confirmed, self.lost)
        self.track_updates += len(pool) + len(unconfirmed)
        if pool:                                               # multi_predict :76-92
            mean = np.stack([t.mean for t in pool])
            cov = np.stack([t.cov for t in pool])
            for k, t in enumerate(pool):
                if t.state != TRACKED:
                    mean[k, 6] = 0
                    mean[k, 7] = 0
            mean, cov = kalman.predict(KIND, mean, cov)
            for k, t in enumerate(pool):
                t.mean, t.cov = mean[k], cov[k]
        # multi_gmc (:94-111) on the pool, then on the unconfirmed tracks; with the identity warp an exact no-op
        if warp is not None:
            for group in (pool, unconfirmed):
                if group:
                    mean, cov = kalman.apply_warp(np.stack([t.mean for t in group]), np.stack([t.cov for t in group]), warp)
                    for k, t in enumerate(group):
                        t.mean, t.cov = mean[k], cov[k]

        activated, refound, newly_lost, newly_removed = [], [], [], []
        m1, ut1, ud1 = assign_with_limit(self._combined(pool, d1, self.fuse_first), self.match_thresh)
        for i, j in m1:
            was_tracked = pool[i].state == TRACKED
            self._apply(pool[i], d1[j])
            (activated if was_tracked else refound).append(pool[i])

        d2 = [mk(j, False) for j in lo]
        rest = [pool[i] for i in ut1 if pool[i].state == TRACKED]
        m2, ut2, _ = assign_with_limit(_iou_cost(rest, d2), 0.5)
        for i, j in m2:
            self._apply(rest[i], d2[j])
            activated.append(rest[i])
        for i in ut2:
            if rest[i].state != LOST:
                rest[i].state = LOST
                newly_lost.append(rest[i])

        left = [d1[j] for j in ud1]
        m3, uu3, ud3 = assign_with_limit(self._combined(unconfirmed, left, True), 0.7)
        for i, j in m3:
            self._apply(unconfirmed[i], left[j])
            activated.append(unconfirmed[i])
        for i in uu3:
            unconfirmed[i].state = REMOVED
            newly_removed.append(unconfirmed[i])

        for j in ud3:                                          # activate :113-126
            t = left[j]
            if t.score < self.new_thresh:
                continue
            self.next_id += 1
            t.tid = self.next_id
            m, c = kalman.initiate(KIND, t.xywh)
            t.mean, t.cov = m[0], c[0]
            t.tracklet_len = 0
            t.state = TRACKED
            t.activated = self.frame_id == 1
            t.frame_id = t.start_frame = self.frame_id
            activated.append(t)

        for t in self.lost:
            if self.frame_id - t.frame_id > self.max_time_lost:
                t.state = REMOVED
                newly_removed.append(t)

        self.tracked = [t for t in self.tracked if t.state == TRACKED]
        self.tracked = _union(self.tracked, activated)
        self.tracked = _union(self.tracked, refound)
        ids = {t.tid for t in self.tracked}
        self.lost = [t for t in self.lost if t.tid not in ids]
        self.lost.extend(newly_lost)
        self.lost = [t for t in self.lost if t.tid not in self.removed_ids]      # the OLD removed list
        self.removed_ids.update(t.tid for t in newly_removed)
        self._drop_duplicates()
        rows = [np.concatenate([t.box(), [t.tid, t.score, t.cls, t.det_ind]]) for t in self.tracked if t.activated]
        return np.asarray(rows)

    def _drop_duplicates(self):                                # :453-465
        a, b = self.tracked, self.lost
        if not a or not b:
            return
        pd = 1 - boxes.iou(_stack_boxes(a), _stack_boxes(b))
        da, db = set(), set()
        for p, q in zip(*np.nonzero(pd < 0.15)):
            if a[p].frame_id - a[p].start_frame > b[q].frame_id - b[q].start_frame:
                db.add(q)
            else:
                da.add(p)
        self.tracked = [t for i, t in enumerate(a) if i not in da]
        self.lost = [t for i, t in enumerate(b) if i not in db]

    def snapshot(self):
        trks = self.tracked + self.lost
        n = len(trks)
        F = next((len(t.smooth_feat) for t in trks if t.smooth_feat is not None), 0)
        return dict(
            n_tracked=np.int32(len(self.tracked)), n_lost=np.int32(len(self.lost)),
            track_id=np.array([t.tid for t in trks], dtype=np.int32),
            state=np.array([t.state for t in trks], dtype=np.int32),
            is_activated=np.array([t.activated for t in trks], dtype=np.int32),
            frame_id=np.array([t.frame_id for t in trks], dtype=np.int32),
            start_frame=np.array([t.start_frame for t in trks], dtype=np.int32),
            tracklet_len=np.array([t.tracklet_len for t in trks], dtype=np.int32),
            score=np.array([t.score for t in trks], dtype=np.float64),
            cls=np.array([t.cls for t in trks], dtype=np.float64),
            det_ind=np.array([t.det_ind for t in trks], dtype=np.float64),
            mean=np.stack([t.mean for t in trks]) if n else np.zeros((0, 8)),
            cov=np.stack([t.cov for t in trks]) if n else np.zeros((0, 8, 8)),
            smooth_feat=(np.stack([t.smooth_feat if t.smooth_feat is not None else np.zeros(F, np.float32) for t in trks])
                         if n and F else np.zeros((n, F), dtype=np.float32)),
        )
