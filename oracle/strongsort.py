"""ORACLE (test infrastructure): the reference's StrongSORT frame step restated on numpy, with the camera-motion
estimator replaced by the identity warp and the ReID network by caller-supplied detection embeddings.

Follows (reference file:line):
  boxmot/trackers/strongsort/strong_sort.py           StrongSORT.__init__ :14-41, update :43-99
  boxmot/trackers/strongsort/sort/tracker.py          predict :59-66, update :73-102, _match :104-155, _initiate_track :157-168
  boxmot/trackers/strongsort/sort/track.py            Track.__init__ :72-99, to_tlwh / to_tlbr :101-127, camera_update :129-138,
                                                      predict :144-150, update :152-178, mark_missed :180-185
  boxmot/trackers/strongsort/sort/detection.py        to_xyah :34-41
  boxmot/trackers/strongsort/sort/linear_assignment.py min_cost_matching :14-79 (scipy linear_sum_assignment on the clipped
                                                      matrix), matching_cascade :82-141 (one round in this version; the
                                                      unmatched tracks come out of a Python set), gate_cost_matrix :144-200
  boxmot/trackers/strongsort/sort/iou_matching.py     iou :10-47, iou_cost :50-87
  boxmot/utils/matching.py                            _cosine_distance :247-267, _nn_cosine_distance :290-308,
                                                      NearestNeighborDistanceMetric :311-378 (float32 gallery arithmetic)
  boxmot/motion/kalman_filters/strongsort_kf.py       oracle/kalman.py kind "xyah_conf"

Two implementation-defined orders of the reference leak into its track ids and are kept literally: scipy's
linear_sum_assignment on exactly tied (clipped) costs, and the iteration order of a CPython set of ints.

`feats[n_dets, F]` holds one embedding per detection row as the ReID seam returns it (`get_features`, already divided by
the Frobenius norm of the whole matrix).  Parity pinned by tests/golden/strongsort_*.npz, generated from the live reference.
"""
from __future__ import annotations

import numpy as np
from scipy.optimize import linear_sum_assignment

from . import kalman

KIND = "xyah_conf"
TENTATIVE, CONFIRMED, DELETED = 1, 2, 3
INFTY_COST = 1e5
CHI2INV95_4 = 9.4877


class _Det:
    def __init__(self, tlwh, conf, cls, det_ind, feat):
        self.tlwh, self.conf, self.cls, self.det_ind, self.feat = tlwh, conf, cls, det_ind, feat

    def to_xyah(self):
        ret = self.tlwh.copy()
        ret[:2] += ret[2:] / 2
        ret[2] /= ret[3]
        return ret


class _Trk:
    def __init__(self, det, tid, n_init, max_age, ema_alpha):
        self.id = tid
        self.conf, self.cls, self.det_ind = det.conf, det.cls, det.det_ind
        self.hits, self.age, self.time_since_update = 1, 1, 0
        self.ema_alpha = ema_alpha
        self.state = TENTATIVE
        det.feat /= np.linalg.norm(det.feat)                    # in place, like track.py:88
        self.features = [det.feat]
        self._n_init, self._max_age = n_init, max_age
        m, c = kalman.initiate(KIND, det.to_xyah())
        self.mean, self.covariance = m[0], c[0]

    def to_tlwh(self):
        ret = self.mean[:4].copy()
        ret[2] *= ret[3]
        ret[:2] -= ret[2:] / 2
        return ret

    def to_tlbr(self):
        ret = self.to_tlwh()
        ret[2:] = ret[:2] + ret[2:]
        return ret

    def camera_update(self, warp=None):                         # track.py:129-138 (warp None = eye(2, 3))
        x1, y1, x2, y2 = self.to_tlbr()
        if warp is not None:
            m = np.array([warp[0], warp[1], [0, 0, 1]]).tolist()
            x1, y1, _ = m @ np.array([x1, y1, 1]).T
            x2, y2, _ = m @ np.array([x2, y2, 1]).T
        w, h = x2 - x1, y2 - y1
        cx, cy = x1 + w / 2, y1 + h / 2
        self.mean[:4] = [cx, cy, w / h, h]

    def predict(self):
        m, c = kalman.predict(KIND, self.mean, self.covariance)
        self.mean, self.covariance = m[0], c[0]
        self.age += 1
        self.time_since_update += 1

    def update(self, det):
        self.conf, self.cls, self.det_ind = det.conf, det.cls, det.det_ind
        m, c = kalman.update(KIND, self.mean, self.covariance, det.to_xyah(), self.conf)
        self.mean, self.covariance = m[0], c[0]
        feature = det.feat / np.linalg.norm(det.feat)
        smooth = self.ema_alpha * self.features[-1] + (1 - self.ema_alpha) * feature
        smooth /= np.linalg.norm(smooth)
        self.features = [smooth]
        self.hits += 1
        self.time_since_update = 0
        if self.state == TENTATIVE and self.hits >= self._n_init:
            self.state = CONFIRMED

    def mark_missed(self):
        if self.state == TENTATIVE:
            self.state = DELETED
        elif self.time_since_update > self._max_age:
            self.state = DELETED


def _nn_cosine(samples, feats):
    """matching.py:247-308 in float32: min over the gallery of 1 - a_hat . b_hat."""
    x = np.asarray(samples)
    y = np.asarray(feats)
    a = x / np.linalg.norm(x, axis=1, keepdims=True)
    b = y / np.linalg.norm(y, axis=1, keepdims=True)
    return (1.0 - np.dot(a, b.T)).min(axis=0)


def _iou_tlwh(bbox, cand):
    tl = np.c_[np.maximum(bbox[0], cand[:, 0])[:, None], np.maximum(bbox[1], cand[:, 1])[:, None]]
    br = np.c_[np.minimum(bbox[0] + bbox[2], cand[:, 0] + cand[:, 2])[:, None],
               np.minimum(bbox[1] + bbox[3], cand[:, 1] + cand[:, 3])[:, None]]
    wh = np.maximum(0.0, br - tl)
    inter = wh.prod(axis=1)
    return inter / (bbox[2:].prod() + cand[:, 2:].prod(axis=1) - inter)


class StrongSORTOracle:
    def __init__(self, max_dist=0.2, max_iou_dist=0.7, max_age=30, n_init=1, nn_budget=100, mc_lambda=0.995, ema_alpha=0.9):
        self.max_dist, self.max_iou_dist, self.max_age, self.n_init = max_dist, max_iou_dist, max_age, n_init
        self.budget, self.mc_lambda, self.ema_alpha = nn_budget, mc_lambda, ema_alpha
        self.tracks: list[_Trk] = []
        self.samples: dict = {}
        self.next_id = 1
        self.track_updates = 0

    # ------------------------------------------------------------------ matching
    def _min_cost_matching(self, metric, max_distance, dets, track_idx, det_idx):
        if len(det_idx) == 0 or len(track_idx) == 0:
            return [], track_idx, det_idx
        cost = metric(track_idx, det_idx)
        cost[cost > max_distance] = max_distance + 1e-5
        rows, cols = linear_sum_assignment(cost)
        matches, ut, ud = [], [], []
        for col, d in enumerate(det_idx):
            if col not in cols:
                ud.append(d)
        for row, t in enumerate(track_idx):
            if row not in rows:
                ut.append(t)
        for row, col in zip(rows, cols):
            t, d = track_idx[row], det_idx[col]
            if cost[row, col] > max_distance:
                ut.append(t)
                ud.append(d)
            else:
                matches.append((t, d))
        return matches, ut, ud

    def _gated_metric(self, dets):
        def metric(track_idx, det_idx):
            feats = np.array([dets[i].feat for i in det_idx])
            cost = np.zeros((len(track_idx), len(det_idx)))
            for r, k in enumerate(track_idx):
                cost[r, :] = _nn_cosine(self.samples[self.tracks[k].id], feats)
            meas = np.asarray([dets[i].to_xyah() for i in det_idx])
            for r, k in enumerate(track_idx):
                trk = self.tracks[k]
                gd = kalman.gating_distance(KIND, trk.mean, trk.covariance, meas, False)
                cost[r, gd > CHI2INV95_4] = INFTY_COST
                cost[r] = self.mc_lambda * cost[r] + (1 - self.mc_lambda) * gd
            return cost
        return metric

    def _iou_metric(self, dets):
        def metric(track_idx, det_idx):
            cost = np.zeros((len(track_idx), len(det_idx)))
            cand = np.asarray([dets[i].tlwh for i in det_idx])
            for r, k in enumerate(track_idx):
                if self.tracks[k].time_since_update > 1:
                    cost[r, :] = INFTY_COST
                    continue
                cost[r, :] = 1.0 - _iou_tlwh(self.tracks[k].to_tlwh(), cand)
            return cost
        return metric

    # ------------------------------------------------------------------ frame step
    def update(self, dets, feats, warp=None):
        assert isinstance(dets, np.ndarray), "dets must be np.ndarray"
        assert dets.ndim == 2, "dets must be two-dimensional"
        assert dets.shape[1] == 6, "dets must have 6 columns"
        dets = np.asarray(dets, dtype=np.float64)
        n = len(dets)
        if len(self.tracks) >= 1:
            for t in self.tracks:
                t.camera_update(warp)
        tlwh = dets[:, :4].copy()
        tlwh[:, 2] = dets[:, 2] - dets[:, 0]
        tlwh[:, 3] = dets[:, 3] - dets[:, 1]
        D = [_Det(tlwh[j], dets[j, 4], dets[j, 5], float(j), np.array(feats[j], dtype=np.float32)) for j in range(n)]
        self.track_updates += len(self.tracks)
        for t in self.tracks:
            t.predict()

        confirmed = [i for i, t in enumerate(self.tracks) if t.state == CONFIRMED]
        unconfirmed = [i for i, t in enumerate(self.tracks) if t.state != CONFIRMED]
        m_a, _, ud = self._min_cost_matching(self._gated_metric(D), self.max_dist, D, list(confirmed), list(range(n)))
        ut_a = list(set(confirmed) - set(k for k, _ in m_a))            # linear_assignment.py:141: CPython set order
        cand = unconfirmed + [k for k in ut_a if self.tracks[k].time_since_update == 1]
        ut_a = [k for k in ut_a if self.tracks[k].time_since_update != 1]
        m_b, ut_b, ud = self._min_cost_matching(self._iou_metric(D), self.max_iou_dist, D, cand, ud)
        matches = m_a + m_b
        unmatched_tracks = list(set(ut_a + ut_b))

        for k, d in matches:
            self.tracks[k].update(D[d])
        for k in unmatched_tracks:
            self.tracks[k].mark_missed()
        for d in ud:
            self.tracks.append(_Trk(D[d], self.next_id, self.n_init, self.max_age, self.ema_alpha))
            self.next_id += 1
        self.tracks = [t for t in self.tracks if t.state != DELETED]

        active = [t.id for t in self.tracks if t.state == CONFIRMED]
        for t in self.tracks:
            if t.state != CONFIRMED:
                continue
            for f in t.features:
                self.samples.setdefault(t.id, []).append(f)
                if self.budget is not None:
                    self.samples[t.id] = self.samples[t.id][-self.budget:]
        self.samples = {k: self.samples[k] for k in active}

        rows = []
        for t in self.tracks:
            if t.state != CONFIRMED or t.time_since_update >= 1:
                continue
            rows.append(np.concatenate((t.to_tlbr(), [t.id], [t.conf], [t.cls], [t.det_ind])).reshape(1, -1))
        return np.concatenate(rows) if rows else np.array([])

    def snapshot(self):
        ts = self.tracks
        n = len(ts)
        return dict(
            track_id=np.array([t.id for t in ts], dtype=np.int32), state=np.array([t.state for t in ts], dtype=np.int32),
            hits=np.array([t.hits for t in ts], dtype=np.int32), age=np.array([t.age for t in ts], dtype=np.int32),
            time_since_update=np.array([t.time_since_update for t in ts], dtype=np.int32),
            mean=np.stack([t.mean for t in ts]) if n else np.zeros((0, 8)),
            cov=np.stack([t.covariance for t in ts]) if n else np.zeros((0, 8, 8)),
            gallery=np.array([len(self.samples.get(t.id, [])) for t in ts], dtype=np.int32),
            feature=np.stack([t.features[-1] for t in ts]) if n else np.zeros((0, 0), dtype=np.float32),
        )
