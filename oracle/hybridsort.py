"""ORACLE (test infrastructure): the reference's HybridSORT frame step restated on numpy.

Follows (reference file:line):
  boxmot/trackers/hybridsort/hybridsort.py   k_previous_obs :22-30, convert_bbox_to_z :33-49 (five-vector
      [x, y, s, score, r]), convert_x_to_bbox :52-63, speed_direction_lt/rt/lb/rb :74-103, KalmanBoxTracker :106-334
      (update_features :188-205, update :220-297, predict :299-322), HybridSORT.__init__ :337-368,
      HybridSORT.update :373-570
  boxmot/motion/kalman_filters/hybridsort_kf.py  predict :339-379, freeze :383-387, unfreeze :390-436 (observation-centric
      re-update; it unpacks the five-vector as x, y, s, r, c - i.e. reads the SCORE as the aspect ratio and interpolates the
      aspect ratio linearly: kept), update :439-528 (Joseph form)
  boxmot/trackers/hybridsort/association.py  cal_score_dif_batch :44-54, linear_assignment :300-311, cost_vel :314-335,
      speed_direction_batch_lt/rt/lb/rb :338-383, associate_4_points_with_score_with_reid :495-581,
      embedding_distance :667-684
  boxmot/utils/iou.py  the similarity behind asso_func (oracle/boxes.py)

What the constructor fixes (hybridsort.py:337-364) and this restatement therefore takes as constants: TCM_first_step with
weight 0, EG_weight_high_score 1.3, long-term ReID weight 0 (the feature bank only enters the cost through that zero
weight, so it is not kept), long-term correction threshold 0.4, track_thresh 0.6 of the score clip, alpha 0.8, ECC off.
`use_byte` (off in hybridsort.yaml and never forwarded by tracker_zoo.py:100-115) is not restated: that branch calls
KalmanBoxTracker.update with an embedding row where the class goes (:470-474) and cannot produce a result row.

Reference quirks kept: dets0 (the rows the class and the last output column are read from) is NOT filtered by det_thresh
while the association rows are (:396-404), so a matched / new tracker takes cls = dets0[row, 5] and det_ind = dets0[row, 6] =
the SCORE of input detection number `row`; all embedding arithmetic is float32 with the in-place normalisations of
update_features (:189, :205 - a new tracker's row is normalised twice, a matched one once before the blend).

`update(dets, feats)`: feats are the rows get_features returns for the detections with conf > det_thresh, in order.
IDs are per tracker instance (KalmanBoxTracker.count is reset in HybridSORT.__init__, :364).

Parity pinned by tests/golden/hybridsort_*.npz, generated from the live reference.
"""
from __future__ import annotations

import numpy as np

from . import boxes
from .lap import assign_no_limit

F9 = np.eye(9)
F9[0, 5] = F9[1, 6] = F9[2, 7] = F9[3, 8] = 1.0
H59 = np.eye(5, 9)
R5 = np.diag([1.0, 1.0, 10.0, 10.0, 10.0])
Q9 = np.diag([1.0, 1.0, 1.0, 1.0, 1.0, 0.01, 0.01, 0.0001, 0.0001])
P0 = np.diag([10.0] * 5 + [1e4] * 4)
I9 = np.eye(9)
ALPHA = 0.8
TRACK_THRESH = 0.6
EG_WEIGHT = 1.3
CORRECTION_THRESH = 0.4


def bbox_to_z(b):
    """[x1, y1, x2, y2, score] -> [x, y, s, score, r] (hybridsort.py:33-49; score > 0 on every caller's path)."""
    w = b[2] - b[0]
    h = b[3] - b[1]
    return np.array([b[0] + w / 2.0, b[1] + h / 2.0, w * h, b[4], w / float(h + 1e-6)])


def x_to_bbox(x):
    with np.errstate(invalid="ignore", divide="ignore"):
        w = np.sqrt(x[2] * x[4])
        h = x[2] / w
    return np.array([x[0] - w / 2.0, x[1] - h / 2.0, x[0] + w / 2.0, x[1] + h / 2.0])


CORNERS = ((0, 1), (0, 3), (2, 1), (2, 3))          # lt, rt, lb, rb as the reference names them: (x index, y index)


def corner_direction(b1, b2, corner):
    ix, iy = corner
    dy, dx = b2[iy] - b1[iy], b2[ix] - b1[ix]
    return np.array([dy, dx]) / (np.sqrt(dy ** 2 + dx ** 2) + 1e-6)


class _KF:
    """9-d filter [u, v, s, c, r, du, dv, ds, dc] with the observation-centric re-update (hybridsort_kf.py)."""

    def __init__(self, z):
        self.x = np.zeros(9)
        self.x[:5] = z
        self.P = P0.copy()
        self.observed = False
        self.saved = None
        self.last_z = None
        self.gap = 0

    def predict(self):
        self.x = F9 @ self.x
        self.P = F9 @ self.P @ F9.T + Q9

    def _correct(self, z):
        y = z - H59 @ self.x
        pht = self.P @ H59.T
        s = H59 @ pht + R5
        k = pht @ np.linalg.inv(s)
        self.x = self.x + k @ y
        ikh = I9 - k @ H59
        self.P = ikh @ self.P @ ikh.T + k @ R5 @ k.T

    def update(self, z):
        self.gap += 1
        if z is None:
            if self.observed:
                self.saved = (self.x.copy(), self.P.copy())
            self.observed = False
            return
        virtual_last = None
        if not self.observed and self.saved is not None:
            self.x, self.P = self.saved
            self.saved = None
            # hybridsort_kf.py:401-413: `x1, y1, s1, r1, c1 = box1` on [x, y, s, score, r]
            x1, y1, s1, r1, c1 = self.last_z
            x2, y2, s2, r2, c2 = z
            with np.errstate(invalid="ignore", divide="ignore"):
                w1, h1 = np.sqrt(s1 * r1), np.sqrt(s1 / r1)
                w2, h2 = np.sqrt(s2 * r2), np.sqrt(s2 / r2)
            g = self.gap
            dx, dy, dw, dh, dc = (x2 - x1) / g, (y2 - y1) / g, (w2 - w1) / g, (h2 - h1) / g, (c2 - c1) / g
            for i in range(g):
                w, h = w1 + (i + 1) * dw, h1 + (i + 1) * dh
                virtual_last = np.array([x1 + (i + 1) * dx, y1 + (i + 1) * dy, w * h, w / float(h), c1 + (i + 1) * dc])
                self._correct(virtual_last)
                if i != g - 1:
                    self.predict()
        self.observed = True
        self._correct(z)
        self.last_z = virtual_last if virtual_last is not None else np.asarray(z, dtype=np.float64).copy()
        self.gap = 0


class _Trk:
    def __init__(self, bbox5, cls, det_ind, feat, tid):
        self.kf = _KF(bbox_to_z(bbox5))
        self.id = tid
        self.time_since_update = 0
        self.hits = 0
        self.hit_streak = 0
        self.age = 0
        self.conf = bbox5[4]
        self.cls = cls
        self.det_ind = det_ind
        self.last_observation = np.array([-1.0, -1, -1, -1, -1])
        self.observations = {}
        self.velocity = [None, None, None, None]
        self.smooth_feat = None
        self.update_features(feat)

    def update_features(self, feat):
        """hybridsort.py:188-205 with adapfs off; float32 throughout, `feat` is a private copy of the caller's row."""
        feat = feat / np.float32(np.linalg.norm(feat))
        if self.smooth_feat is None:
            self.smooth_feat = feat
        else:
            self.smooth_feat = np.float32(ALPHA) * self.smooth_feat + np.float32(1 - ALPHA) * feat
        self.smooth_feat = self.smooth_feat / np.float32(np.linalg.norm(self.smooth_feat))

    def k_previous(self, k):
        if not self.observations:
            return np.array([-1.0, -1, -1, -1, -1])
        for i in range(k):
            if self.age - (k - i) in self.observations:
                return self.observations[self.age - (k - i)]
        return self.observations[max(self.observations)]

    def predict(self):
        if self.kf.x[7] + self.kf.x[2] <= 0:
            self.kf.x[7] *= 0.0
        self.kf.predict()
        self.age += 1
        if self.time_since_update > 0:
            self.hit_streak = 0
        self.time_since_update += 1
        return x_to_bbox(self.kf.x), np.clip(self.kf.x[3], TRACK_THRESH, 1.0)

    def update(self, bbox5, cls, det_ind, feat, delta_t, update_feature=True):
        if bbox5 is None:
            self.kf.update(None)
            return
        self.conf = bbox5[4]
        self.cls = cls
        self.det_ind = det_ind
        if self.last_observation.sum() >= 0:
            vel = None
            for i in range(delta_t):                               # every observation of the window adds its direction
                prev = self.observations.get(self.age - i - 1)
                if prev is not None:
                    d = [corner_direction(prev, bbox5, c) for c in CORNERS]
                    vel = d if vel is None else [a + b for a, b in zip(vel, d)]
            if vel is None:
                vel = [corner_direction(self.last_observation, bbox5, c) for c in CORNERS]
            self.velocity = vel
        self.last_observation = bbox5
        self.observations[self.age] = bbox5
        self.time_since_update = 0
        self.hits += 1
        self.hit_streak += 1
        self.kf.update(bbox_to_z(bbox5))
        if update_feature:
            self.update_features(feat)


def embedding_distance(track_feats, det_feats):
    """association.py:667-684: scipy cdist 'cosine' in float64, clamped at 0 -> [T, D]."""
    from scipy.spatial.distance import cdist
    if len(track_feats) == 0 or len(det_feats) == 0:
        return np.zeros((len(track_feats), len(det_feats)))
    return np.maximum(0.0, cdist(np.asarray(track_feats, dtype=np.float64), np.asarray(det_feats, dtype=np.float64), "cosine"))


def associate(dets5, trks, asso, thr, vels, prev_obs, inertia, emb_cost, w, h):
    """associate_4_points_with_score_with_reid (association.py:495-581) with the constructor's constants.
    dets5 [D, 5], trks [T, 5] (box + clipped filter score), vels [4][T, 2], emb_cost [D, T]."""
    D, T = len(dets5), len(trks)
    if T == 0:
        return np.empty((0, 2), dtype=int), np.arange(D), np.empty((0,), dtype=int)
    valid = (prev_obs[:, 4] >= 0).astype(np.float64)[:, None]
    angle = None
    for (ix, iy), v in zip(CORNERS, vels):
        dx = dets5[None, :, ix] - prev_obs[:, None, ix]
        dy = dets5[None, :, iy] - prev_obs[:, None, iy]
        norm = np.sqrt(dx ** 2 + dy ** 2) + 1e-6
        X, Y = dx / norm, dy / norm                                    # [T, D]
        cosang = np.clip(v[:, 1:2] * X + v[:, 0:1] * Y, -1, 1)
        diff = (np.pi / 2.0 - np.abs(np.arccos(cosang))) / np.pi
        term = ((valid * diff) * inertia).T * dets5[:, 4:5]
        angle = term if angle is None else angle + term
    sim = boxes.similarity(asso, dets5[:, :4], trks[:, :4], w, h)      # [D, T]
    score_dif = np.abs(trks[None, :, 4] - dets5[:, None, 4])
    angle = angle - score_dif * 0.0                                     # TCM_first_step_weight = 0
    if min(sim.shape):
        m = assign_no_limit(1.0 * (-(sim + angle)) + EG_WEIGHT * emb_cost + 0.0)
    else:
        m = np.empty((0, 2), dtype=int)
    ud = [d for d in range(D) if d not in m[:, 0]]
    ut = [t for t in range(T) if t not in m[:, 1]]
    sim_thre = sim - score_dif
    keep = []
    for d, t in m:
        if emb_cost[d, t] > CORRECTION_THRESH and sim_thre[d, t] < thr:
            ud.append(d)
            ut.append(t)
        else:
            keep.append((d, t))
    return np.array(keep, dtype=int).reshape(-1, 2), np.array(ud, dtype=int), np.array(ut, dtype=int)


class HybridSortOracle:
    def __init__(self, det_thresh=0.0, max_age=30, min_hits=3, iou_threshold=0.3, delta_t=3, asso_func="iou", inertia=0.2,
                 use_byte=False):
        assert not use_byte, "the reference's use_byte branch cannot produce a result row (see the module docstring)"
        self.max_age, self.min_hits, self.iou_threshold = max_age, min_hits, iou_threshold
        self.det_thresh, self.delta_t, self.asso_func, self.inertia = det_thresh, delta_t, asso_func, inertia
        self.trackers: list[_Trk] = []
        self.frame_count = 0
        self.count = 0
        self.track_updates = 0
        self.stats = dict(ocr_frames=0, oru=0, corrections=0)

    def update(self, dets, feats, img=(1080, 1920)):
        """dets [n, 6]; feats [k, F] float32: get_features rows of the k detections with conf > det_thresh."""
        self.frame_count += 1
        h, w = img.shape[0:2] if hasattr(img, "shape") else img
        dets = np.asarray(dets, dtype=np.float64).reshape(-1, 6)
        dets0 = np.concatenate([dets, dets[:, 4:5]], axis=1)            # UNFILTERED: cls at [row, 5], score at [row, 6]
        keep = dets[:, 4] > self.det_thresh
        d1 = dets[keep][:, :5]
        feats = np.asarray(feats, dtype=np.float32).reshape(len(d1), -1) if len(d1) else np.zeros((0, 0), dtype=np.float32)

        T0 = len(self.trackers)
        trks = np.zeros((T0, 5))
        dead = []
        for t, trk in enumerate(self.trackers):
            pos, ks = trk.predict()
            trks[t, :4] = pos
            trks[t, 4] = ks
            if np.any(np.isnan(pos)):
                dead.append(t)
        trks = trks[[t for t in range(T0) if t not in dead]]
        for t in reversed(dead):
            self.trackers.pop(t)
        T = len(self.trackers)
        self.track_updates += T
        vels = [np.array([t.velocity[c] if t.velocity[c] is not None else np.zeros(2) for t in self.trackers]).reshape(T, 2)
                for c in range(4)]
        last = np.array([t.last_observation for t in self.trackers]).reshape(T, 5)
        kobs = np.array([t.k_previous(self.delta_t) for t in self.trackers]).reshape(T, 5)

        emb = embedding_distance([t.smooth_feat for t in self.trackers], feats).T          # [D, T]
        m, ud, ut = associate(d1, trks, self.asso_func, self.iou_threshold, vels, kobs, self.inertia, emb, w, h)
        self.stats["corrections"] += (min(len(d1), T) - len(m)) if T and len(d1) else 0
        for d, t in m:
            self._upd(self.trackers[t], d1[d], dets0[d, 5], dets0[d, 6], feats[d].copy(), True)

        if len(ud) > 0 and len(ut) > 0:                                        # OCR :513-545
            left = boxes.similarity(self.asso_func, d1[ud][:, :4], last[ut][:, :4], w, h)
            if left.max() > self.iou_threshold:
                self.stats["ocr_frames"] += 1
                gd, gt = [], []
                for a, b in assign_no_limit(-left):
                    if left[a, b] < self.iou_threshold:
                        continue
                    self._upd(self.trackers[ut[b]], d1[ud[a]], dets0[ud[a], 5], dets0[ud[a], 6], None, False)
                    gd.append(ud[a])
                    gt.append(ut[b])
                ud = np.setdiff1d(ud, np.array(gd))
                ut = np.setdiff1d(ut, np.array(gt))

        for t in ut:
            self.trackers[t].update(None, None, None, None, self.delta_t)
        for d in ud:
            self.trackers.append(_Trk(d1[d], dets0[d, 5], dets0[d, 6], feats[d].copy(), self.count))
            self.count += 1
        rows = []
        i = len(self.trackers)
        for trk in reversed(self.trackers):
            box = x_to_bbox(trk.kf.x) if trk.last_observation.sum() < 0 else trk.last_observation[:4]
            if trk.time_since_update < 1 and (trk.hit_streak >= self.min_hits or self.frame_count <= self.min_hits):
                rows.append(np.concatenate([box, [trk.id + 1, trk.conf, trk.cls, trk.det_ind]]))
            i -= 1
            if trk.time_since_update > self.max_age:
                self.trackers.pop(i)
        return np.stack(rows) if rows else np.empty((0, 7))

    def _upd(self, trk, bbox5, cls, det_ind, feat, update_feature):
        if not trk.kf.observed and trk.kf.saved is not None:
            self.stats["oru"] += 1
        trk.update(bbox5, cls, det_ind, feat, self.delta_t, update_feature)

    def snapshot(self):
        ts = self.trackers
        n = len(ts)
        F = len(ts[0].smooth_feat) if n else 0
        return dict(
            n=np.int32(n),
            track_id=np.array([t.id for t in ts], dtype=np.int32),
            age=np.array([t.age for t in ts], dtype=np.int32),
            time_since_update=np.array([t.time_since_update for t in ts], dtype=np.int32),
            hits=np.array([t.hits for t in ts], dtype=np.int32),
            hit_streak=np.array([t.hit_streak for t in ts], dtype=np.int32),
            observed=np.array([int(t.kf.observed) for t in ts], dtype=np.int32),
            x=np.stack([t.kf.x for t in ts]) if n else np.zeros((0, 9)),
            P=np.stack([t.kf.P for t in ts]) if n else np.zeros((0, 9, 9)),
            velocity=np.array([[v if v is not None else np.zeros(2) for v in t.velocity] for t in ts]).reshape(n, 4, 2),
            last_observation=np.array([t.last_observation for t in ts], dtype=np.float64).reshape(n, 5),
            smooth_feat=np.stack([t.smooth_feat for t in ts]).astype(np.float32) if n else np.zeros((0, F), dtype=np.float32),
        )
