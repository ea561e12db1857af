"""ORACLE (test infrastructure): the reference's OC-SORT frame step restated on numpy.

Follows (reference file:line):
  boxmot/trackers/ocsort/ocsort.py   k_previous_obs :14-22, convert_bbox_to_z :25-37,
      convert_x_to_bbox :40-54, speed_direction :57-62, KalmanBoxTracker :65-187
      (update :130-166, predict :168-181), OCSort.__init__ :191-216, OCSort.update :218-379
  boxmot/motion/kalman_filters/ocsort_kf.py  predict :339-379, freeze :383-387,
      unfreeze :390-434 (observation-centric re-update), update :437-526 (Joseph form)
  boxmot/utils/association.py  speed_direction_batch :8-17, linear_assignment :20-24,
      associate :111-201
The BYTE stage (`use_byte`, ocsort.py:293-317) is restated too.  IDs are per tracker
instance (the reference resets KalmanBoxTracker.count in OCSort.__init__, :216).

Parity pinned by tests/golden/ocsort_*.npz, generated from the live reference.
"""
from __future__ import annotations

import numpy as np

from . import boxes
from .lap import assign_no_limit

F7 = np.eye(7)
F7[0, 4] = F7[1, 5] = F7[2, 6] = 1.0
H47 = np.eye(4, 7)
R4 = np.diag([1.0, 1.0, 10.0, 10.0])
Q7 = np.diag([1.0, 1.0, 1.0, 1.0, 0.01, 0.01, 0.0001])
P0 = np.diag([10.0, 10.0, 10.0, 10.0, 1e4, 1e4, 1e4])
I7 = np.eye(7)


def bbox_to_z(b):
    w = b[2] - b[0]
    h = b[3] - b[1]
    return np.array([b[0] + w / 2.0, b[1] + h / 2.0, w * h, w / float(h + 1e-6)])


def x_to_bbox(x):
    with np.errstate(invalid="ignore", divide="ignore"):
        w = np.sqrt(x[2] * x[3])
        h = x[2] / w
    return np.array([x[0] - w / 2.0, x[1] - h / 2.0, x[0] + w / 2.0, x[1] + h / 2.0])


def direction(b1, b2):
    cx1, cy1 = (b1[0] + b1[2]) / 2.0, (b1[1] + b1[3]) / 2.0
    cx2, cy2 = (b2[0] + b2[2]) / 2.0, (b2[1] + b2[3]) / 2.0
    speed = np.array([cy2 - cy1, cx2 - cx1])
    return speed / (np.sqrt((cy2 - cy1) ** 2 + (cx2 - cx1) ** 2) + 1e-6)


class _KF:
    """7-d constant-velocity filter with observation-centric re-update (ocsort_kf.py)."""

    def __init__(self, z):
        self.x = np.zeros(7)
        self.x[:4] = z
        self.P = P0.copy()
        self.observed = False
        self.saved = None            # (x, P) frozen at the first missed frame
        self.last_z = None           # last real measurement (history_obs[index1])
        self.gap = 0                 # entries appended to history_obs since last_z

    def predict(self):
        self.x = F7 @ self.x
        self.P = F7 @ self.P @ F7.T + Q7

    def _correct(self, z):
        y = z - H47 @ self.x
        pht = self.P @ H47.T
        s = H47 @ pht + R4
        k = pht @ np.linalg.inv(s)
        self.x = self.x + k @ y
        ikh = I7 - k @ H47
        self.P = ikh @ self.P @ ikh.T + k @ R4 @ k.T

    def update(self, z):
        self.gap += 1
        if z is None:
            if self.observed:
                self.saved = (self.x.copy(), self.P.copy())
            self.observed = False
            return
        virtual_last = None
        if not self.observed and self.saved is not None:
            # unfreeze: replay a straight-line virtual trajectory over the gap
            self.x, self.P = self.saved
            self.saved = None
            x1, y1, s1, r1 = self.last_z
            w1, h1 = np.sqrt(s1 * r1), np.sqrt(s1 / r1)
            x2, y2, s2, r2 = z
            w2, h2 = np.sqrt(s2 * r2), np.sqrt(s2 / r2)
            g = self.gap
            dx, dy, dw, dh = (x2 - x1) / g, (y2 - y1) / g, (w2 - w1) / g, (h2 - h1) / g
            for i in range(g):
                w, h = w1 + (i + 1) * dw, h1 + (i + 1) * dh
                virtual_last = np.array([x1 + (i + 1) * dx, y1 + (i + 1) * dy, w * h, w / float(h)])
                self._correct(virtual_last)
                if i != g - 1:
                    self.predict()
        self.observed = True
        self._correct(z)                 # the real measurement is applied on top (double application)
        # the restored history ends with the last virtual box, not with z (ocsort_kf.py:391-395)
        self.last_z = virtual_last if virtual_last is not None else np.asarray(z, dtype=np.float64).copy()
        self.gap = 0


class _Trk:
    def __init__(self, bbox5, cls, det_ind, tid):
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
        self.velocity = None

    def k_previous(self, k):
        if not self.observations:
            return np.array([-1.0, -1, -1, -1, -1])
        for i in range(k):
            if self.age - (k - i) in self.observations:
                return self.observations[self.age - (k - i)]
        return self.observations[max(self.observations)]

    def predict(self):
        if self.kf.x[6] + self.kf.x[2] <= 0:
            self.kf.x[6] *= 0.0
        self.kf.predict()
        self.age += 1
        if self.time_since_update > 0:
            self.hit_streak = 0
        self.time_since_update += 1
        return x_to_bbox(self.kf.x)

    def update(self, bbox5, cls, det_ind, delta_t):
        self.det_ind = det_ind
        if bbox5 is None:
            self.kf.update(None)
            return
        self.conf = bbox5[4]
        self.cls = cls
        if self.last_observation.sum() >= 0:
            prev = None
            for i in range(delta_t):
                if self.age - (delta_t - i) in self.observations:
                    prev = self.observations[self.age - (delta_t - i)]
                    break
            if prev is None:
                prev = self.last_observation
            self.velocity = direction(prev, bbox5)
        self.last_observation = bbox5
        self.observations[self.age] = bbox5
        self.time_since_update = 0
        self.hits += 1
        self.hit_streak += 1
        self.kf.update(bbox_to_z(bbox5))


def associate(dets5, trks, asso, thr, velocities, prev_obs, inertia, w, h):
    """association.py:111-201 -> matches[k,2] (det, trk), unmatched_dets, unmatched_trks, used_lap."""
    D, T = len(dets5), len(trks)
    if T == 0:
        return np.empty((0, 2), dtype=int), np.arange(D), np.empty((0,), dtype=int), False
    cx_d, cy_d = (dets5[:, 0] + dets5[:, 2]) / 2.0, (dets5[:, 1] + dets5[:, 3]) / 2.0
    cx_p, cy_p = (prev_obs[:, 0] + prev_obs[:, 2]) / 2.0, (prev_obs[:, 1] + prev_obs[:, 3]) / 2.0
    dx = cx_d[None, :] - cx_p[:, None]
    dy = cy_d[None, :] - cy_p[:, None]
    norm = np.sqrt(dx ** 2 + dy ** 2) + 1e-6
    X, Y = dx / norm, dy / norm                                       # [T, D]
    cosang = np.clip(velocities[:, 1:2] * X + velocities[:, 0:1] * Y, -1, 1)
    diff = (np.pi / 2.0 - np.abs(np.arccos(cosang))) / np.pi
    valid = (prev_obs[:, 4] >= 0).astype(np.float64)[:, None]
    sim = boxes.similarity(asso, dets5[:, :4], trks[:, :4], w, h)     # [D, T]
    angle = ((valid * diff) * inertia).T * dets5[:, 4:5]
    used_lap = False
    if min(sim.shape):
        a = (sim > thr).astype(np.int32)
        if a.sum(1).max() == 1 and a.sum(0).max() == 1:
            m = np.stack(np.where(a), axis=1)
        else:
            m = assign_no_limit(-(sim + angle))
            used_lap = True
    else:
        m = np.empty((0, 2), dtype=int)
    ud = [d for d in range(D) if d not in m[:, 0]]
    ut = [t for t in range(T) if t not in m[:, 1]]
    keep = []
    for d, t in m:
        if sim[d, t] < thr:
            ud.append(d)
            ut.append(t)
        else:
            keep.append((d, t))
    return np.array(keep, dtype=int).reshape(-1, 2), np.array(ud, dtype=int), np.array(ut, dtype=int), used_lap


class OCSortOracle:
    def __init__(self, per_class=True, det_thresh=0.2, max_age=30, min_hits=3, asso_threshold=0.3, delta_t=3,
                 asso_func="iou", inertia=0.2, use_byte=False):
        self.max_age, self.min_hits, self.asso_threshold = max_age, min_hits, asso_threshold
        self.det_thresh, self.delta_t, self.asso_func, self.inertia, self.use_byte = det_thresh, delta_t, asso_func, inertia, use_byte
        self.trackers: list[_Trk] = []
        self.frame_count = 0
        self.count = 0
        self.track_updates = 0
        self.stats = dict(lap_frames=0, ocr_frames=0, oru=0, byte_matches=0)

    def update(self, dets, img):
        assert isinstance(dets, np.ndarray), "dets must be np.ndarray"
        assert dets.ndim == 2, "dets must be two-dimensional"
        assert dets.shape[1] == 6, "dets must have 6 columns"
        self.frame_count += 1
        h, w = img.shape[0:2] if hasattr(img, "shape") else img
        dets = np.asarray(dets, dtype=np.float64)
        ind = np.arange(len(dets), dtype=np.float64)
        conf = dets[:, 4]
        second = (conf > 0.1) & (conf < self.det_thresh)
        first = conf > self.det_thresh
        d2, ind2 = dets[second], ind[second]
        d1, ind1 = dets[first], ind[first]

        trks = np.zeros((len(self.trackers), 5))
        dead = []
        for t, trk in enumerate(self.trackers):
            pos = trk.predict()
            trks[t, :4] = pos
            if np.any(np.isnan(pos)):
                dead.append(t)
        trks = trks[[t for t in range(len(self.trackers)) if t not in dead]]
        for t in reversed(dead):
            self.trackers.pop(t)
        T = len(self.trackers)
        self.track_updates += T
        vel = np.array([t.velocity if t.velocity is not None else np.zeros(2) for t in self.trackers]).reshape(T, 2)
        last = np.array([t.last_observation for t in self.trackers]).reshape(T, 5)
        kobs = np.array([t.k_previous(self.delta_t) for t in self.trackers]).reshape(T, 5)

        m, ud, ut, used = associate(d1[:, :5], trks, self.asso_func, self.asso_threshold, vel, kobs, self.inertia, w, h)
        self.stats["lap_frames"] += int(used)
        for d, t in m:
            self._upd(self.trackers[t], d1[d, :5], d1[d, 5], ind1[d])

        if self.use_byte and len(d2) > 0 and len(ut) > 0:                    # BYTE stage :293-317
            left = boxes.ASSO[self.asso_func](d2[:, :4], trks[ut][:, :4])
            if left.max() > self.asso_threshold:
                gone = []
                for d, k in assign_no_limit(-left):
                    if left[d, k] < self.asso_threshold:
                        continue
                    self._upd(self.trackers[ut[k]], d2[d, :5], d2[d, 5], ind2[d])
                    self.stats["byte_matches"] += 1
                    gone.append(ut[k])
                ut = np.setdiff1d(ut, np.array(gone))

        if len(ud) > 0 and len(ut) > 0:                                        # OCR :319-345
            left = boxes.similarity(self.asso_func, d1[ud][:, :4], last[ut][:, :4], w, h)
            if left.max() > self.asso_threshold:
                self.stats["ocr_frames"] += 1
                gd, gt = [], []
                for a, b in assign_no_limit(-left):
                    if left[a, b] < self.asso_threshold:
                        continue
                    self._upd(self.trackers[ut[b]], d1[ud[a], :5], d1[ud[a], 5], ind1[ud[a]])
                    gd.append(ud[a])
                    gt.append(ut[b])
                ud = np.setdiff1d(ud, np.array(gd))
                ut = np.setdiff1d(ut, np.array(gt))

        for t in ut:
            self.trackers[t].update(None, None, None, self.delta_t)
        for d in ud:
            self.trackers.append(_Trk(d1[d, :5], d1[d, 5], ind1[d], self.count))
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
        return np.stack(rows) if rows else np.array([])

    def _upd(self, trk, bbox5, cls, det_ind):
        if not trk.kf.observed and trk.kf.saved is not None:
            self.stats["oru"] += 1
        trk.update(bbox5, cls, det_ind, self.delta_t)

    def snapshot(self):
        ts = self.trackers
        n = len(ts)
        return dict(
            n=np.int32(n),
            track_id=np.array([t.id for t in ts], dtype=np.int32),
            age=np.array([t.age for t in ts], dtype=np.int32),
            time_since_update=np.array([t.time_since_update for t in ts], dtype=np.int32),
            hits=np.array([t.hits for t in ts], dtype=np.int32),
            hit_streak=np.array([t.hit_streak for t in ts], dtype=np.int32),
            observed=np.array([int(t.kf.observed) for t in ts], dtype=np.int32),
            x=np.stack([t.kf.x for t in ts]) if n else np.zeros((0, 7)),
            P=np.stack([t.kf.P for t in ts]) if n else np.zeros((0, 7, 7)),
            velocity=np.array([t.velocity if t.velocity is not None else np.zeros(2) for t in ts]).reshape(n, 2),
            last_observation=np.array([t.last_observation for t in ts], dtype=np.float64).reshape(n, 5),
        )
