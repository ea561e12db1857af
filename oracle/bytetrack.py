"""ORACLE (test infrastructure): the reference's ByteTrack frame step restated on numpy.

Follows boxmot/trackers/bytetrack/byte_tracker.py (reference file:line):
  STrack.__init__ :14-25, multi_predict :35-48, activate :50-62, re_activate :64-76,
  update :78-98, xyxy :100-111, BYTETracker.__init__ :115-130, update :132-281,
  joint_stracks :287-298, sub_stracks :301-309, remove_duplicate_stracks :312-325
and boxmot/utils/matching.py iou_distance :94-119, fuse_score :213-221,
linear_assignment :56-71; boxmot/trackers/bytetrack/basetrack.py :8-55.

List semantics (order of tracked/lost lists, first-wins joins, the "removed" list that is
consulted one frame late and never shrinks) are reproduced literally; numerics are batched
per association stage with the dense oracle Kalman filter.  IDs are per tracker instance
(1, 2, 3, ...) - the reference's class-global counter reset per stream.

Parity pinned by tests/golden/bytetrack_*.npz, generated from the live reference.
"""
from __future__ import annotations

import numpy as np

from . import boxes, kalman
from .lap import assign_with_limit

NEW, TRACKED, LOST, REMOVED = 0, 1, 2, 3


class _Trk:
    __slots__ = ("mean", "cov", "state", "activated", "tid", "frame_id", "start_frame",
                 "tracklet_len", "score", "cls", "det_ind")

    def box(self):
        m = self.mean[:4].copy()
        m[2] *= m[3]
        return boxes.xywh_to_xyxy(m)


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


def _minus(a, ids):
    return [t for t in a if t.tid not in ids]


class ByteTrackOracle:
    kind = "xyah"

    def __init__(self, track_thresh=0.45, match_thresh=0.8, track_buffer=25, frame_rate=30):
        self.track_thresh = track_thresh
        self.match_thresh = match_thresh
        self.det_thresh = track_thresh
        self.max_time_lost = int(frame_rate / 30.0 * track_buffer)
        self.frame_id = 0
        self.tracked: list[_Trk] = []
        self.lost: list[_Trk] = []
        self.removed_ids: set[int] = set()
        self.next_id = 0
        self.track_updates = 0
        self.last_matches = None

    # ------------------------------------------------------------------ helpers
    def _kf_update(self, pairs, z, score, cls, det_ind, reactivate_lost=True):
        """pairs: list of (track, det_row) ; batched update + bookkeeping (:64-98)."""
        if not pairs:
            return
        trks = [p[0] for p in pairs]
        rows = np.array([p[1] for p in pairs])
        mean = np.stack([t.mean for t in trks])
        cov = np.stack([t.cov for t in trks])
        mean, cov = kalman.update(self.kind, mean, cov, z[rows])
        for k, t in enumerate(trks):
            t.mean, t.cov = mean[k], cov[k]
            if t.state == TRACKED:
                t.tracklet_len += 1
            else:
                t.tracklet_len = 0
            t.state = TRACKED
            t.activated = True
            t.frame_id = self.frame_id
            t.score, t.cls, t.det_ind = score[rows[k]], cls[rows[k]], det_ind[rows[k]]

    # ------------------------------------------------------------------ frame step
    def update(self, dets, _img=None):
        assert isinstance(dets, np.ndarray), "dets must be np.ndarray"
        assert dets.ndim == 2, "dets must be two-dimensional"
        assert dets.shape[1] == 6, "dets must have 6 columns"
        dets = np.asarray(dets, dtype=np.float64)
        self.frame_id += 1
        conf = dets[:, 4]
        hi = np.nonzero(conf > self.track_thresh)[0]
        lo = np.nonzero((conf > 0.1) & (conf < self.track_thresh))[0]

        def prep(idx):
            raw = dets[idx, :4]
            xywh = boxes.xyxy_to_xywh(raw)
            return dict(xyxy=boxes.xywh_to_xyxy(xywh), z=boxes.tlwh_to_xyah(boxes.xywh_to_tlwh(xywh)),
                        score=dets[idx, 4], cls=dets[idx, 5], ind=idx.astype(np.float64))
        d1, d2 = prep(hi), prep(lo)

        unconfirmed = [t for t in self.tracked if not t.activated]
        confirmed = [t for t in self.tracked if t.activated]
        pool = _union(confirmed, self.lost)
        self.track_updates += len(pool) + len(unconfirmed)

        if pool:                                                     # multi_predict :35-48
            mean = np.stack([t.mean for t in pool])
            cov = np.stack([t.cov for t in pool])
            for k, t in enumerate(pool):
                if t.state != TRACKED:
                    mean[k, 7] = 0
            mean, cov = kalman.predict(self.kind, mean, cov)
            for k, t in enumerate(pool):
                t.mean, t.cov = mean[k], cov[k]

        def fused(cost, score):                                       # matching.py:213-221
            if cost.size == 0:
                return cost
            return 1 - (1 - cost) * score[None, :]

        def iou_cost(trks, det_xyxy):                                 # matching.py:94-119
            if len(trks) == 0 or len(det_xyxy) == 0:
                return np.zeros((len(trks), len(det_xyxy)), dtype=np.float32)
            return 1 - boxes.iou(_stack_boxes(trks), det_xyxy)

        # first association
        c1 = fused(iou_cost(pool, d1["xyxy"]), d1["score"])
        m1, ut1, ud1 = assign_with_limit(c1, self.match_thresh)
        refound = [pool[i] for i, _ in m1 if pool[i].state != TRACKED]
        self._kf_update([(pool[i], j) for i, j in m1], d1["z"], d1["score"], d1["cls"], d1["ind"])

        # second association: still-Tracked leftovers against low-score detections
        rest = [pool[i] for i in ut1 if pool[i].state == TRACKED]
        c2 = iou_cost(rest, d2["xyxy"])
        m2, ut2, _ = assign_with_limit(c2, 0.5)
        self._kf_update([(rest[i], j) for i, j in m2], d2["z"], d2["score"], d2["cls"], d2["ind"])
        newly_lost = []
        for i in ut2:
            if rest[i].state != LOST:
                rest[i].state = LOST
                newly_lost.append(rest[i])

        # unconfirmed tracks (never predicted) against the remaining high detections
        left = np.asarray(ud1, dtype=int)
        c3 = fused(iou_cost(unconfirmed, d1["xyxy"][left]), d1["score"][left])
        m3, uu3, ud3 = assign_with_limit(c3, 0.7)
        self._kf_update([(unconfirmed[i], left[j]) for i, j in m3], d1["z"], d1["score"], d1["cls"], d1["ind"])
        newly_removed = []
        for i in uu3:
            unconfirmed[i].state = REMOVED
            newly_removed.append(unconfirmed[i])

        # new tracks :242-248
        born = []
        for j in left[np.asarray(ud3, dtype=int)]:
            if d1["score"][j] < self.det_thresh:
                continue
            t = _Trk()
            self.next_id += 1
            t.tid = self.next_id
            m, c = kalman.initiate(self.kind, d1["z"][j])
            t.mean, t.cov = m[0], c[0]
            t.tracklet_len = 0
            t.state = TRACKED
            t.activated = self.frame_id == 1
            t.frame_id = t.start_frame = self.frame_id
            t.score, t.cls, t.det_ind = d1["score"][j], d1["cls"][j], d1["ind"][j]
            born.append(t)

        # age-out :250-253
        for t in self.lost:
            if self.frame_id - t.frame_id > self.max_time_lost:
                t.state = REMOVED
                newly_removed.append(t)

        # merge :257-268
        activated = [pool[i] for i, _ in m1 if pool[i] not in refound] + [rest[i] for i, _ in m2] \
            + [unconfirmed[i] for i, _ in m3] + born
        self.tracked = [t for t in self.tracked if t.state == TRACKED]
        self.tracked = _union(self.tracked, activated)
        self.tracked = _union(self.tracked, refound)
        self.lost = _minus(self.lost, {t.tid for t in self.tracked})
        self.lost.extend(newly_lost)
        self.lost = _minus(self.lost, self.removed_ids)          # consults the OLD removed list
        self.removed_ids.update(t.tid for t in newly_removed)
        self._drop_duplicates()
        self.last_matches = (m1, m2, m3)

        rows = [np.concatenate([t.box(), [t.tid, t.score, t.cls, t.det_ind]])
                for t in self.tracked if t.activated]
        return np.asarray(rows)

    def _drop_duplicates(self):                                       # :312-325
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

    # ------------------------------------------------------------------ parity probe
    def snapshot(self):
        """Track records in list order (tracked list, then lost list)."""
        trks = self.tracked + self.lost
        n = len(trks)
        out = dict(
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
        )
        return out
