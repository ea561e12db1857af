"""DeepOCSORT with the reference's constructor and update() contract
(boxmot/trackers/deepocsort/deep_ocsort.py:308-520), one stream per object like the reference.

Like the StrongSORT drop-in this tracker is operator-backed, not one fused kernel: the list logic of DeepOCSort.update
(track order, counters, the observation dictionary, which detections become new tracks) runs in Python as in the reference,
and the numeric steps go through the CUDA operator kernels of the C-ABI, batched over the tracks of a frame:
  8-d filter predict with state-dependent Q          b200track_kf8_predict   (deep_ocsort.py:76-80, :246-270)
  Joseph-form update with state-dependent R          b200track_kf8_update    (deep_ocsort.py:83-87, :217-218; deepocsort_kf.py:549-563)
  observation-centric re-update (virtual trajectory) b200track_kf8_oru       (deepocsort_kf.py:433-478) - one launch for all
                                                                              re-found tracks of the frame
  camera correction of live and frozen states        b200track_kf_apply_warp (deepocsort_kf.py:389-405)
  IoU / GIoU / ... similarity                        b200track_box_similarity (iou.py)
  embedding similarity                               b200track_dot_matrix    (deep_ocsort.py:433)
  adaptive appearance weight                         b200track_aw_max_metric (association.py:79-108)
  velocity-direction + appearance cost               b200track_ocm_cost      (association.py:130-172)
  assignment without a cost limit                    b200track_lapjv         (association.py:20-24)
All Kalman updates of a frame (first round and OCR round touch disjoint tracks and nothing in between reads the filter)
are applied in one batch.  Only the default `new_kf` filter is built (new_kf_off=True raises).  The ReID network and the
camera-motion estimator are out of scope (BASELINE.json): embeddings come from a `model` with get_features(xyxys, img)
or from update(..., feats=...), the warp from update(..., warp=...) (None = identity).
"""
from __future__ import annotations

import numpy as np

from .. import _lib, _ops
from .bytetrack import _SingleStreamTracker, _device_index


class _Trk:
    __slots__ = ("id", "conf", "cls", "det_ind", "time_since_update", "hits", "hit_streak", "age", "last_observation",
                 "observations", "velocity", "emb", "frozen", "x", "P", "observed", "saved", "hist_last", "misses")

    def __init__(self, det7, tid, emb):
        b = det7[:5]
        self.conf, self.cls, self.det_ind = det7[4], det7[5], det7[6]
        self.id = tid
        self.time_since_update = self.hits = self.hit_streak = self.age = 0
        self.last_observation = np.array([-1, -1, -1, -1, -1])
        self.observations = {}
        self.velocity = None
        self.emb = emb
        self.frozen = False
        # KalmanFilter(dim_x=8, dim_z=4) as set up by deep_ocsort.py:103-138; P = 4 / 100 x the process noise of (w, h)
        w, h = b[2] - b[0], b[3] - b[1]
        self.x = np.array([b[0] + w / 2.0, b[1] + h / 2.0, w, h, 0.0, 0.0, 0.0, 0.0])
        self.P = np.diag(((w / 20) ** 2 * 4, (h / 20) ** 2 * 4, (w / 20) ** 2 * 4, (h / 20) ** 2 * 4,
                          (w / 160) ** 2 * 100, (h / 160) ** 2 * 100, (w / 160) ** 2 * 100, (h / 160) ** 2 * 100))
        self.observed = False
        self.saved = None            # [x, P, last_measurement] frozen at the first missed frame (deepocsort_kf.py:383-387, :507-514)
        self.hist_last = None        # last entry of the filter's observation history (a measurement or a virtual box)
        self.misses = 0              # None entries appended since the freeze

    def box(self):
        x, y, w, h = self.x[:4]
        return np.array([x - w / 2, y - h / 2, x + w / 2, y + h / 2])

    def k_previous(self, k):
        if not self.observations:
            return [-1, -1, -1, -1, -1]
        for i in range(k):
            if self.age - (k - i) in self.observations:
                return self.observations[self.age - (k - i)]
        return self.observations[max(self.observations)]

    def move_observations(self, m, t, delta_t):
        """deep_ocsort.py:226-241; last_observation and observations[age of that frame] are ONE array in the reference,
        so it moves twice while inside the delta_t window - kept (same aliasing here)."""
        if self.last_observation.sum() > 0:
            ps = m @ self.last_observation[:4].reshape(2, 2).T + t[:, None]
            self.last_observation[:4] = ps.T.reshape(-1)
        for dt in range(delta_t, -1, -1):
            if self.age - dt in self.observations:
                o = self.observations[self.age - dt]
                ps = m @ o[:4].reshape(2, 2).T + t[:, None]
                o[:4] = ps.T.reshape(-1)

    def observe(self, det7, delta_t):
        """The bookkeeping half of KalmanBoxTracker.update (deep_ocsort.py:183-216); the filter half is batched."""
        bbox = det7[:5]
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

    def miss(self):
        """kf.update(None) (deepocsort_kf.py:506-521) + frozen = True."""
        if self.observed:
            self.saved = [self.x.copy(), self.P.copy(), np.array(self.hist_last, dtype=np.float64)]
            self.misses = 0
        self.observed = False
        self.misses += 1
        self.hist_last = None
        self.frozen = True


class DeepOCSort:
    def __init__(self, model_weights=None, device=0, fp16=False, per_class=True, det_thresh=0.3, max_age=30, min_hits=3,
                 iou_threshold=0.3, delta_t=3, asso_func="iou", inertia=0.2, w_association_emb=0.5, alpha_fixed_emb=0.95,
                 aw_param=0.5, embedding_off=False, cmc_off=False, aw_off=False, new_kf_off=False, model=None, **kwargs):
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
        self.trackers: list[_Trk] = []
        self.frame_count = 0
        self._next_id = 1                                  # KalmanBoxTracker.count = 1 (deep_ocsort.py:347), per instance here
        self.stats = dict(lap_frames=0, ocr_frames=0, oru=0)
        _lib.load()
        _ops._torch()                                      # fail loudly without a CUDA device: there is no CPU path

    # ------------------------------------------------------------------ association.py:111-201
    def _associate(self, dets, trks, vel, kobs, dets_embs, w, h):
        D, T = len(dets), len(trks)
        if T == 0:
            return [], list(range(D)), []
        if D == 0:
            return [], [], list(range(T))
        sim = _ops.box_similarity(self.asso_func, dets[:, :4], trks[:, :4], w, h)
        a = sim > self.iou_threshold
        if a.sum(1).max() == 1 and a.sum(0).max() == 1:
            pairs = [(int(d), int(t)) for d, t in zip(*np.where(a))]
        else:
            emb = None
            if not self.embedding_off:
                emb = _ops.dot_matrix(dets_embs, np.vstack([t.emb for t in self.trackers]))
                emb[sim <= 0] = 0
                emb = _ops.aw_max_metric(emb, self.w_association_emb, self.aw_param) if not self.aw_off else emb * self.w_association_emb
            x, _ = _ops.lapjv(_ops.ocm_cost(sim, dets[:, :5], vel, kobs, self.inertia, emb))
            pairs = [(d, int(x[d])) for d in range(D) if x[d] >= 0]
            self.stats["lap_frames"] += 1
        md, mt = {d for d, _ in pairs}, {t for _, t in pairs}
        ud = [d for d in range(D) if d not in md]
        ut = [t for t in range(T) if t not in mt]
        keep = []
        for d, t in pairs:
            if sim[d, t] < self.iou_threshold:
                ud.append(d)
                ut.append(t)
            else:
                keep.append((d, t))
        return keep, ud, ut

    def update(self, dets, img, feats=None, warp=None):
        """`feats`: what self.model.get_features(dets[conf > det_thresh, :4], img) would return (deep_ocsort.py:382-390);
        `warp`: this frame's externally estimated 2x3 camera motion (self.cmc.apply in the reference, :393-396)."""
        _SingleStreamTracker._check(dets)
        assert isinstance(img, (np.ndarray, tuple)), f"Unsupported 'img' input type '{type(img)}', valid format is np.ndarray"
        self.frame_count += 1
        h, w = img.shape[:2] if isinstance(img, np.ndarray) else img
        dets = np.hstack([np.asarray(dets, dtype=np.float64), np.arange(len(dets), dtype=np.float64).reshape(-1, 1)])
        dets = dets[dets[:, 4] > self.det_thresh]
        D = len(dets)
        if self.embedding_off or D == 0:
            dets_embs = np.ones((D, 1))
        elif feats is not None:
            dets_embs = np.asarray(feats)
            assert len(dets_embs) == D, "feats must hold one row per detection with conf > det_thresh"
        else:
            if self.model is None:
                raise ValueError("DeepOCSort needs `feats` or a `model` with get_features(xyxys, img)")
            dets_embs = self.model.get_features(dets[:, 0:4], img)
        trs = self.trackers
        if not self.cmc_off and warp is not None and trs:
            wm = np.asarray(warp, dtype=np.float64)
            m, t = wm[:, :2], wm[:, 2]
            frozen = [k for k in trs if not k.observed and k.saved is not None]
            for k in trs:
                k.move_observations(m, t, self.delta_t)
            mean, cov = _ops.kf_apply_warp(np.stack([k.x for k in trs] + [k.saved[0] for k in frozen]),
                                           np.stack([k.P for k in trs] + [k.saved[1] for k in frozen]), wm)
            for i, k in enumerate(trs):
                k.x, k.P = mean[i], cov[i]
            for i, k in enumerate(frozen):
                lm = k.saved[2]
                lm[:2] = m @ lm[:2] + t
                lm[2:] = m @ lm[2:]
                k.saved[0], k.saved[1] = mean[len(trs) + i], cov[len(trs) + i]
        trust = (dets[:, 4] - self.det_thresh) / (1 - self.det_thresh)
        af = self.alpha_fixed_emb
        dets_alpha = af + (1 - af) * (1 - trust)

        # KalmanBoxTracker.predict for every track (deep_ocsort.py:246-270), one launch
        if trs:
            for k in trs:
                if k.x[2] + k.x[6] <= 0:
                    k.x[6] = 0
                if k.x[3] + k.x[7] <= 0:
                    k.x[7] = 0
                if k.frozen:
                    k.x[6] = k.x[7] = 0
            mean, cov = _ops.kf8_predict(np.stack([k.x for k in trs]), np.stack([k.P for k in trs]))
            for i, k in enumerate(trs):
                k.x, k.P = mean[i], cov[i]
                k.age += 1
                if k.time_since_update > 0:
                    k.hit_streak = 0
                k.time_since_update += 1
            self.trackers = trs = [k for k in trs if not np.any(np.isnan(k.box()))]
        T = len(trs)
        trks = np.stack([k.box() for k in trs]) if T else np.zeros((0, 4))
        vel = np.array([k.velocity if k.velocity is not None else np.zeros(2) for k in trs]).reshape(T, 2)
        last = np.array([k.last_observation for k in trs], dtype=np.float64).reshape(T, 5)
        kobs = np.array([k.k_previous(self.delta_t) for k in trs], dtype=np.float64).reshape(T, 5)

        matched, ud, ut = self._associate(dets, trks, vel, kobs, dets_embs, w, h)
        # second round: OCR on the last observations (deep_ocsort.py:456-491)
        if len(ud) > 0 and len(ut) > 0:
            left = _ops.box_similarity(self.asso_func, dets[ud][:, :4], last[ut][:, :4], w, h)
            if left.max() > self.iou_threshold:
                self.stats["ocr_frames"] += 1
                x, _ = _ops.lapjv(_ops.ocm_cost(left))
                gd, gt = [], []
                for a in range(len(ud)):
                    b = int(x[a])
                    if b < 0 or left[a, b] < self.iou_threshold:
                        continue
                    matched.append((ud[a], ut[b]))
                    gd.append(ud[a])
                    gt.append(ut[b])
                ud = np.setdiff1d(ud, np.array(gd)).astype(int).tolist()
                ut = np.setdiff1d(ut, np.array(gt)).astype(int).tolist()

        # every Kalman update of the frame in one batch: re-update of the re-found tracks first, then the measurement
        if matched:
            ks = [trs[t] for _, t in matched]
            z = np.empty((len(matched), 4))
            wh = np.stack([k.x[2:4] for k in ks])                     # R from the state BEFORE a possible unfreeze
            for i, (d, _) in enumerate(matched):
                b = dets[d]
                bw, bh = b[2] - b[0], b[3] - b[1]
                z[i] = (b[0] + bw / 2.0, b[1] + bh / 2.0, bw, bh)
            oru = [i for i, k in enumerate(ks) if not k.observed and k.saved is not None]
            if oru:
                self.stats["oru"] += len(oru)
                mean, cov, virt = _ops.kf8_oru(np.stack([ks[i].saved[0] for i in oru]), np.stack([ks[i].saved[1] for i in oru]),
                                               np.stack([ks[i].saved[2] for i in oru]), z[oru], [ks[i].misses + 1 for i in oru])    # gap = index2 - index1
                for j, i in enumerate(oru):
                    ks[i].x, ks[i].P, ks[i].hist_last, ks[i].saved = mean[j], cov[j], virt[j], None
            mean, cov = _ops.kf8_update(np.stack([k.x for k in ks]), np.stack([k.P for k in ks]), z, wh)
            oru = set(oru)
            for i, ((d, _), k) in enumerate(zip(matched, ks)):
                k.x, k.P = mean[i], cov[i]
                if i not in oru:
                    k.hist_last = z[i].copy()
                k.observed = True
                k.observe(dets[d], self.delta_t)
                k.emb = dets_alpha[d] * k.emb + (1 - dets_alpha[d]) * dets_embs[d]        # update_emb (deep_ocsort.py:222-224)
                k.emb /= np.linalg.norm(k.emb)
        for t in ut:
            trs[t].miss()
        for d in ud:
            trs.append(_Trk(dets[d], self._next_id, dets_embs[d]))
            self._next_id += 1
        rows = []
        i = len(trs)
        for k in reversed(trs):
            box = k.box() if k.last_observation.sum() < 0 else k.last_observation[:4]
            if k.time_since_update < 1 and (k.hit_streak >= self.min_hits or self.frame_count <= self.min_hits):
                rows.append(np.concatenate((box, [k.id], [k.conf], [k.cls], [k.det_ind])).reshape(1, -1))
            i -= 1
            if k.time_since_update > self.max_age:
                trs.pop(i)
        return np.concatenate(rows) if rows else np.array([])

    def state(self):
        ts = self.trackers
        n = len(ts)
        return dict(
            track_id=np.array([t.id for t in ts], dtype=np.int32), age=np.array([t.age for t in ts], dtype=np.int32),
            time_since_update=np.array([t.time_since_update for t in ts], dtype=np.int32),
            hits=np.array([t.hits for t in ts], dtype=np.int32), hit_streak=np.array([t.hit_streak for t in ts], dtype=np.int32),
            observed=np.array([int(t.observed) for t in ts], dtype=np.int32), frozen=np.array([int(t.frozen) for t in ts], dtype=np.int32),
            x=np.stack([t.x for t in ts]) if n else np.zeros((0, 8)), P=np.stack([t.P for t in ts]) if n else np.zeros((0, 8, 8)),
            velocity=np.array([t.velocity if t.velocity is not None else np.zeros(2) for t in ts]).reshape(n, 2),
            last_observation=np.array([t.last_observation for t in ts], dtype=np.float64).reshape(n, 5),
            emb=np.stack([np.asarray(t.emb, dtype=np.float64) for t in ts]) if n else np.zeros((0, 0)))
