"""Test infrastructure (CPU oracle): DeepOCSORT's adaptive appearance weight, restated from the reference
(boxmot/utils/association.py:79-108).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may import this.
Pinned by tests/golden/aux_ops.npz (outputs of the live reference, tests/golden/make_golden.py::gen_aux)."""
import numpy as np


def _weights(m, axis, bottom):
    """per row (axis=1) / column (axis=0): 1 - max(second / first - bottom, 0) / (1 - bottom), 0 if first == 0"""
    n = m.shape[axis]
    other = m.shape[1 - axis]
    if n < 2:
        return np.ones(other)
    part = -np.sort(-m, axis=axis)
    first = np.take(part, 0, axis=axis)
    second = np.take(part, 1, axis=axis)
    with np.errstate(divide="ignore", invalid="ignore"):
        w = 1 - np.maximum(second / first - bottom, 0) / (1 - bottom)
    return np.where(first == 0, 0.0, w)


def compute_aw_max_metric(emb_cost, w_association_emb, bottom=0.5):
    """association.py:79-108: w_emb = w_assoc * row weight * column weight; returns w_emb * emb_cost."""
    m = np.asarray(emb_cost, dtype=np.float64)
    w = np.full_like(m, w_association_emb)
    w *= _weights(m, 1, bottom)[:, None]
    w *= _weights(m, 0, bottom)[None, :]
    return w * m


# ======================================================================================================================
# The DeepOCSORT frame step (test infrastructure, like the rest of oracle/): restated on numpy from
#   boxmot/trackers/deepocsort/deep_ocsort.py  k_previous_obs :15-23, convert_bbox_to_z_new :41-46,
#       convert_x_to_bbox_new :49-51, speed_direction :68-73, new_kf_process_noise :76-80, new_kf_measurement_noise :83-87,
#       KalmanBoxTracker :90-305 (new_kf branch), DeepOCSort.update :357-520
#   boxmot/motion/kalman_filters/deepocsort_kf.py  predict :340-381, freeze :383-387, apply_affine_correction :389-405,
#       unfreeze :433-478, update :480-569
#   boxmot/utils/association.py  associate :111-201 with emb_cost, compute_aw_max_metric :79-108
# Only the default new_kf (8-d x, y, w, h filter) is restated; `new_kf_off=True` is refused.
# Pinned by tests/golden/deepocsort_*.npz (outputs of the live reference, tests/golden/make_golden.py::gen_deepocsort).
# Quirks kept on purpose: the virtual trajectory of unfreeze reads [x, y, w, h] as [x, y, s, r] and uses R = I, Q = I;
# after a re-update the filter's observation history ends with the last VIRTUAL box, which becomes `last_measurement`
# at the next freeze; R of a real update comes from the state before unfreeze; last_observation and observations[age]
# are one array, so a camera correction moves it twice while it is inside the delta_t window.
from . import boxes as _boxes            # noqa: E402
from .lap import assign_no_limit as _assign_no_limit          # noqa: E402

_F8 = np.eye(8)
_F8[0, 4] = _F8[1, 5] = _F8[2, 6] = _F8[3, 7] = 1.0
_H48 = np.eye(4, 8)
_I8 = np.eye(8)


def process_noise(w, h, p=1 / 20, v=1 / 160):
    return np.diag(((p * w) ** 2, (p * h) ** 2, (p * w) ** 2, (p * h) ** 2, (v * w) ** 2, (v * h) ** 2, (v * w) ** 2, (v * h) ** 2))


def measurement_noise(w, h, m=1 / 20):
    return np.diag(((m * w) ** 2, (m * h) ** 2, (m * w) ** 2, (m * h) ** 2))


def kf8_predict(x, P, Q):
    return _F8 @ x, _F8 @ P @ _F8.T + Q


def kf8_correct(x, P, z, R):
    """deepocsort_kf.py:549-563: Joseph form with an explicit inverse of S."""
    y = z - _H48 @ x
    PHT = P @ _H48.T
    S = _H48 @ PHT + R
    K = PHT @ np.linalg.inv(S)
    x = x + K @ y
    I_KH = _I8 - K @ _H48
    return x, I_KH @ P @ I_KH.T + K @ R @ K.T


def kf8_virtual_trajectory(x, P, box1, box2, gap):
    """deepocsort_kf.py:444-478 on the restored state -> x, P, the observation history entries it appended."""
    x1, y1, s1, r1 = box1
    w1, h1 = np.sqrt(s1 * r1), np.sqrt(s1 / r1)
    x2, y2, s2, r2 = box2
    w2, h2 = np.sqrt(s2 * r2), np.sqrt(s2 / r2)
    dx, dy, dw, dh = (x2 - x1) / gap, (y2 - y1) / gap, (w2 - w1) / gap, (h2 - h1) / gap
    hist = []
    for i in range(gap):
        xx, yy, w, h = x1 + (i + 1) * dx, y1 + (i + 1) * dy, w1 + (i + 1) * dw, h1 + (i + 1) * dh
        nb = np.array([xx, yy, w * h, w / float(h)])
        hist.append(nb)
        x, P = kf8_correct(x, P, nb, np.eye(4))
        if i != gap - 1:
            x, P = kf8_predict(x, P, _I8)
    return x, P, hist


class _KF8:
    def __init__(self, z):
        self.x = np.zeros(8)
        self.x[:4] = z
        self.P = process_noise(z[2], z[3])
        self.P[:4, :4] *= 4
        self.P[4:, 4:] *= 100
        self.observed = False
        self.saved = None                 # (x, P, history, last_measurement) at the first missed frame
        self.history = []
        self.last_measurement = None

    def affine(self, m, t):
        big = np.kron(np.eye(4), m)
        self.x = big @ self.x
        self.x[:2] += t
        self.P = big @ self.P @ big.T
        if not self.observed and self.saved is not None:
            sx, sP, sh, lm = self.saved
            sx = big @ sx
            sx[:2] += t
            lm = lm.copy()
            lm[:2] = m @ lm[:2] + t
            lm[2:] = m @ lm[2:]
            self.saved = (sx, big @ sP @ big.T, sh, lm)

    def update(self, z, R=None):
        self.history.append(z)
        if z is None:
            if self.observed:
                self.last_measurement = self.history[-2]
                self.saved = (self.x.copy(), self.P.copy(), list(self.history), self.last_measurement.copy())
            self.observed = False
            return False
        oru = False
        if not self.observed and self.saved is not None:
            new_hist = self.history
            self.x, self.P, hist, box1 = self.saved
            self.saved = None             # the snapshot's own attr_saved is never read again before the next freeze
            real = [k for k, d in enumerate(new_hist) if d is not None]
            gap = real[-1] - real[-2]
            self.x, self.P, virt = kf8_virtual_trajectory(self.x, self.P, box1, z, gap)
            self.history = hist[:-1] + virt
            self.last_measurement = box1
            oru = True
        self.observed = True
        self.x, self.P = kf8_correct(self.x, self.P, z, np.eye(4) if R is None else R)
        return oru


class _DTrk:
    def __init__(self, det7, tid, emb):
        b = det7[:5]
        self.conf, self.cls, self.det_ind = det7[4], det7[5], det7[6]
        w, h = b[2] - b[0], b[3] - b[1]
        self.kf = _KF8(np.array([b[0] + w / 2.0, b[1] + h / 2.0, w, h]))
        self.id = tid
        self.time_since_update = self.hits = self.hit_streak = self.age = 0
        self.last_observation = np.array([-1, -1, -1, -1, -1])
        self.observations = {}
        self.velocity = None
        self.emb = emb
        self.frozen = False

    def box(self):
        x, y, w, h = self.kf.x[:4]
        return np.array([x - w / 2, y - h / 2, x + w / 2, y + h / 2])

    def k_previous(self, k):
        if not self.observations:
            return [-1, -1, -1, -1, -1]
        for i in range(k):
            if self.age - (k - i) in self.observations:
                return self.observations[self.age - (k - i)]
        return self.observations[max(self.observations)]

    def affine(self, m, t, delta_t):
        if self.last_observation.sum() > 0:
            ps = m @ self.last_observation[:4].reshape(2, 2).T + t[:, None]
            self.last_observation[:4] = ps.T.reshape(-1)
        for dt in range(delta_t, -1, -1):
            if self.age - dt in self.observations:
                o = self.observations[self.age - dt]
                ps = m @ o[:4].reshape(2, 2).T + t[:, None]
                o[:4] = ps.T.reshape(-1)
        self.kf.affine(m, t)

    def predict(self):
        x = self.kf.x
        if x[2] + x[6] <= 0:
            x[6] = 0
        if x[3] + x[7] <= 0:
            x[7] = 0
        if self.frozen:
            x[6] = x[7] = 0
        self.kf.x, self.kf.P = kf8_predict(x, self.kf.P, process_noise(x[2], x[3]))
        self.age += 1
        if self.time_since_update > 0:
            self.hit_streak = 0
        self.time_since_update += 1
        return self.box()

    def update(self, det7, delta_t):
        if det7 is None:
            self.kf.update(None)
            self.frozen = True
            return False
        bbox = det7[:5]                  # a view: last_observation and observations[age] are this one array
        self.conf, self.cls, self.det_ind = det7[4], det7[5], det7[6]
        self.frozen = False
        if self.last_observation.sum() >= 0:
            prev = None
            for dt in range(delta_t, 0, -1):
                if self.age - dt in self.observations:
                    prev = self.observations[self.age - dt]
                    break
            if prev is None:
                prev = self.last_observation
            cx1, cy1 = (prev[0] + prev[2]) / 2.0, (prev[1] + prev[3]) / 2.0
            cx2, cy2 = (bbox[0] + bbox[2]) / 2.0, (bbox[1] + bbox[3]) / 2.0
            speed = np.array([cy2 - cy1, cx2 - cx1])
            self.velocity = speed / (np.sqrt((cy2 - cy1) ** 2 + (cx2 - cx1) ** 2) + 1e-6)
        self.last_observation = bbox
        self.observations[self.age] = bbox
        self.time_since_update = 0
        self.hits += 1
        self.hit_streak += 1
        R = measurement_noise(self.kf.x[2], self.kf.x[3])
        w, h = bbox[2] - bbox[0], bbox[3] - bbox[1]
        return self.kf.update(np.array([bbox[0] + w / 2.0, bbox[1] + h / 2.0, w, h]), R)

    def update_emb(self, emb, alpha):
        self.emb = alpha * self.emb + (1 - alpha) * emb
        self.emb /= np.linalg.norm(self.emb)


def associate_emb(dets5, trks, asso, thr, velocities, prev_obs, inertia, w, h, emb_cost, w_assoc_emb, aw_off, aw_param):
    """association.py:111-201 with the appearance term -> matches[k,2] (det, trk), unmatched_dets, unmatched_trks, used_lap."""
    D, T = len(dets5), len(trks)
    if T == 0:
        return np.empty((0, 2), dtype=int), np.arange(D), np.empty((0,), dtype=int), False
    cx_d, cy_d = (dets5[:, 0] + dets5[:, 2]) / 2.0, (dets5[:, 1] + dets5[:, 3]) / 2.0
    cx_p, cy_p = (prev_obs[:, 0] + prev_obs[:, 2]) / 2.0, (prev_obs[:, 1] + prev_obs[:, 3]) / 2.0
    dx = cx_d[None, :] - cx_p[:, None]
    dy = cy_d[None, :] - cy_p[:, None]
    norm = np.sqrt(dx ** 2 + dy ** 2) + 1e-6
    X, Y = dx / norm, dy / norm
    cosang = np.clip(velocities[:, 1:2] * X + velocities[:, 0:1] * Y, -1, 1)
    diff = (np.pi / 2.0 - np.abs(np.arccos(cosang))) / np.pi
    valid = (prev_obs[:, 4] >= 0).astype(np.float64)[:, None]
    sim = _boxes.similarity(asso, dets5[:, :4], trks[:, :4], w, h)
    angle = ((valid * diff) * inertia).T * dets5[:, 4:5]
    used_lap = False
    if min(sim.shape):
        a = (sim > thr).astype(np.int32)
        if a.sum(1).max() == 1 and a.sum(0).max() == 1:
            m = np.stack(np.where(a), axis=1)
        else:
            if emb_cost is None:
                emb = 0
            else:
                emb = emb_cost
                emb[sim <= 0] = 0
                emb = compute_aw_max_metric(emb, w_assoc_emb, aw_param) if not aw_off else emb * w_assoc_emb
            m = _assign_no_limit(-(sim + angle + emb))
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


class DeepOCSortOracle:
    def __init__(self, det_thresh=0.3, max_age=30, min_hits=3, iou_threshold=0.3, delta_t=3, asso_func="iou", inertia=0.2,
                 w_association_emb=0.5, alpha_fixed_emb=0.95, aw_param=0.5, embedding_off=False, cmc_off=False, aw_off=False,
                 new_kf_off=False):
        assert not new_kf_off, "only the default new_kf filter is restated"
        self.det_thresh, self.max_age, self.min_hits, self.iou_threshold = det_thresh, max_age, min_hits, iou_threshold
        self.delta_t, self.asso_func, self.inertia = delta_t, asso_func, inertia
        self.w_association_emb, self.alpha_fixed_emb, self.aw_param = w_association_emb, alpha_fixed_emb, aw_param
        self.embedding_off, self.cmc_off, self.aw_off = embedding_off, cmc_off, aw_off
        self.trackers: list[_DTrk] = []
        self.frame_count = 0
        self.count = 1                                             # deep_ocsort.py:347
        self.track_updates = 0
        self.stats = dict(lap_frames=0, ocr_frames=0, oru=0)

    def update(self, dets, feats, img_hw=(1080, 1920), warp=None):
        """`feats`: seam features of the detections with conf > det_thresh, in their order ([D', F] float32);
        `warp`: this frame's 2x3 camera motion (None = identity, exact no-op)."""
        assert isinstance(dets, np.ndarray) and dets.ndim == 2 and dets.shape[1] == 6
        self.frame_count += 1
        h, w = img_hw
        dets = np.hstack([dets, np.arange(len(dets)).reshape(-1, 1)])
        dets = dets[dets[:, 4] > self.det_thresh]
        dets_embs = np.ones((len(dets), 1)) if self.embedding_off or len(dets) == 0 else feats
        if not self.cmc_off and warp is not None:
            m, t = np.asarray(warp)[:, :2], np.asarray(warp)[:, 2]
            for trk in self.trackers:
                trk.affine(m, t, self.delta_t)
        trust = (dets[:, 4] - self.det_thresh) / (1 - self.det_thresh)
        af = self.alpha_fixed_emb
        dets_alpha = af + (1 - af) * (1 - trust)

        trks = np.zeros((len(self.trackers), 5))
        trk_embs, dead = [], []
        for t, trk in enumerate(self.trackers):
            pos = trk.predict()
            trks[t, :4] = pos
            if np.any(np.isnan(pos)):
                dead.append(t)
            else:
                trk_embs.append(trk.emb)
        trks = trks[[t for t in range(len(self.trackers)) if t not in dead]]
        trk_embs = np.vstack(trk_embs) if trk_embs else np.array(trk_embs)
        for t in reversed(dead):
            self.trackers.pop(t)
        T = len(self.trackers)
        self.track_updates += T
        vel = np.array([t.velocity if t.velocity is not None else np.zeros(2) for t in self.trackers]).reshape(T, 2)
        last = np.array([t.last_observation for t in self.trackers]).reshape(T, 5)
        kobs = np.array([t.k_previous(self.delta_t) for t in self.trackers]).reshape(T, 5)

        emb1 = None if self.embedding_off or len(dets) == 0 or T == 0 else dets_embs @ trk_embs.T
        m, ud, ut, used = associate_emb(dets[:, :5], trks, self.asso_func, self.iou_threshold, vel, kobs, self.inertia, w, h,
                                        emb1, self.w_association_emb, self.aw_off, self.aw_param)
        self.stats["lap_frames"] += int(used)
        for d, t in m:
            self._upd(self.trackers[t], dets[d], dets_embs[d], dets_alpha[d])
        if len(ud) > 0 and len(ut) > 0:
            left = _boxes.similarity(self.asso_func, dets[ud][:, :4], last[ut][:, :4], w, h)
            if left.max() > self.iou_threshold:
                self.stats["ocr_frames"] += 1
                gd, gt = [], []
                for a, b in _assign_no_limit(-left):
                    if left[a, b] < self.iou_threshold:
                        continue
                    self._upd(self.trackers[ut[b]], dets[ud[a]], dets_embs[ud[a]], dets_alpha[ud[a]])
                    gd.append(ud[a])
                    gt.append(ut[b])
                ud = np.setdiff1d(ud, np.array(gd))
                ut = np.setdiff1d(ut, np.array(gt))
        for t in ut:
            self.trackers[t].update(None, self.delta_t)
        for d in ud:
            self.trackers.append(_DTrk(dets[d], self.count, dets_embs[d]))
            self.count += 1
        rows = []
        i = len(self.trackers)
        for trk in reversed(self.trackers):
            box = trk.box() if trk.last_observation.sum() < 0 else trk.last_observation[:4]
            if trk.time_since_update < 1 and (trk.hit_streak >= self.min_hits or self.frame_count <= self.min_hits):
                rows.append(np.concatenate([box, [trk.id, trk.conf, trk.cls, trk.det_ind]]))
            i -= 1
            if trk.time_since_update > self.max_age:
                self.trackers.pop(i)
        return np.stack(rows) if rows else np.array([])

    def _upd(self, trk, det7, emb, alpha):
        self.stats["oru"] += int(trk.update(det7, self.delta_t))
        trk.update_emb(emb, alpha)

    def snapshot(self):
        ts = self.trackers
        n = len(ts)
        return dict(
            track_id=np.array([t.id for t in ts], dtype=np.int32), age=np.array([t.age for t in ts], dtype=np.int32),
            time_since_update=np.array([t.time_since_update for t in ts], dtype=np.int32),
            hits=np.array([t.hits for t in ts], dtype=np.int32), hit_streak=np.array([t.hit_streak for t in ts], dtype=np.int32),
            observed=np.array([int(t.kf.observed) for t in ts], dtype=np.int32),
            frozen=np.array([int(t.frozen) for t in ts], dtype=np.int32),
            x=np.stack([t.kf.x for t in ts]) if n else np.zeros((0, 8)),
            P=np.stack([t.kf.P for t in ts]) if n else np.zeros((0, 8, 8)),
            velocity=np.array([t.velocity if t.velocity is not None else np.zeros(2) for t in ts]).reshape(n, 2),
            last_observation=np.array([t.last_observation for t in ts], dtype=np.float64).reshape(n, 5),
            emb=np.stack([np.asarray(t.emb, dtype=np.float64) for t in ts]) if n else np.zeros((0, 0)))
