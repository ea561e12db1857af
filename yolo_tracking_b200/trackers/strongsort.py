"""StrongSORT with the reference's constructor and update() contract
(boxmot/trackers/strongsort/strong_sort.py:13-99), one stream per object like the reference.

Default path: a one-stream context of the batched StrongSORT frame step (csrc/strongsort_step.cu, kind "strongsort" of
`BatchedTracker`): camera correction, Kalman predict, the tensor-core gallery distance, Mahalanobis gate + motion fusion,
both assignment rounds (scipy's linear_sum_assignment restated bit-faithfully including ties, the unmatched set in CPython's
set order), Kalman update, feature smoothing, lifecycle, gallery ring and the result rows all run on the device - eight
launches per frame, the host only hands buffers over.  Track many streams with one `BatchedTracker("strongsort", S, ...)`.

Shapes the tensor-core gallery distance does not take (feature size not a multiple of 64, nn_budget above 128 or None, more
than 256 tracks / detections, mc_lambda <= 0) fall back to the operator-backed form: the list logic of
strongsort/sort/tracker.py in Python, every numeric step a CUDA operator of the C-ABI
  Kalman predict / update with confidence-scaled noise  b200track_kf_predict / _kf_update (strongsort_kf.py:88-189)
  gallery cosine distance                               b200track_nn_cosine_distance     (matching.py:247-378)
  feature smoothing, first features, camera correction  b200track_ema_unit_features / _unit_features / _camera_update_xyah
  Mahalanobis gate + motion fusion                      b200track_gate_cost              (linear_assignment.py:144-200)
  IoU cost                                              b200track_iou_distance           (iou_matching.py:50-87)
  assignment on the clipped matrix                      b200track_linear_sum_assignment  (linear_assignment.py:59-61)
The ReID network and the ECC camera-motion estimator are out of scope (BASELINE.json): embeddings come from a
`model` object with get_features(xyxys, img) or from `update(..., feats=...)`; the warp is passed in (None = identity).
"""
from __future__ import annotations

import numpy as np

from .. import _lib, _ops
from ..batch import BatchedTracker
from .bytetrack import _SingleStreamTracker, _device_index

TENTATIVE, CONFIRMED, DELETED = 1, 2, 3
KIND = _lib.KF_XYAH_CONF
INFTY_COST = 1e5


class _Track:
    __slots__ = ("id", "conf", "cls", "det_ind", "hits", "age", "time_since_update", "state", "feature", "mean", "covariance")

    def to_tlwh(self):
        ret = self.mean[:4].copy()
        ret[2] *= ret[3]
        ret[:2] -= ret[2:] / 2
        return ret

    def to_tlbr(self):
        ret = self.to_tlwh()
        ret[2:] = ret[:2] + ret[2:]
        return ret


class _FusedStream:
    """One-stream context of the batched StrongSORT step, device buffers held as torch tensors."""

    def __init__(self, device, max_tracks, max_dets, dim, **cfg):
        torch = _ops._torch()
        self.torch = torch
        self.D, self.T, self.F = max_dets, max_tracks, dim
        self.batch = BatchedTracker("strongsort", 1, max_tracks=max_tracks, max_dets=max_dets, device=device, feat_dim=dim, **cfg)
        dev = f"cuda:{device}"
        self.h_dets = torch.zeros((1, max_dets, 6), dtype=torch.float64, pin_memory=True)
        self.h_feats = torch.zeros((1, max_dets, dim), dtype=torch.float32, pin_memory=True)
        self.d_dets = torch.zeros((1, max_dets, 6), dtype=torch.float64, device=dev)
        self.d_feats = torch.zeros((1, max_dets, dim), dtype=torch.float32, device=dev)
        self.d_nd = torch.zeros((1,), dtype=torch.int32, device=dev)
        self.d_warp = torch.zeros((1, 6), dtype=torch.float64, device=dev)
        self.d_out = torch.zeros((1, max_tracks, 8), dtype=torch.float64, device=dev)
        self.d_nout = torch.zeros((2,), dtype=torch.int32, device=dev)

    def step(self, dets, feats, warp):
        torch = self.torch
        n = len(dets)
        if n > self.D:
            raise ValueError(f"{n} detections exceed max_dets={self.D}")
        self.h_dets[0, :n] = torch.from_numpy(np.ascontiguousarray(dets, dtype=np.float64))
        self.d_dets.copy_(self.h_dets, non_blocking=True)
        if n:
            self.h_feats[0, :n] = torch.from_numpy(np.ascontiguousarray(feats, dtype=np.float32))
            self.d_feats.copy_(self.h_feats, non_blocking=True)
        self.d_nd.fill_(n)
        if warp is not None:
            self.d_warp.copy_(torch.from_numpy(np.asarray(warp, dtype=np.float64).reshape(1, 6)))
        stream = torch.cuda.current_stream().cuda_stream
        self.batch.step_device(self.d_dets, self.d_nd, self.d_out, self.d_nout, d_feats=self.d_feats, stream=stream,
                               d_warps=self.d_warp if warp is not None else None)
        m = int(self.d_nout[0].item())
        self.batch.sync()                                   # capacity overflows of this step raise here
        return self.d_out[0, :m].cpu().numpy() if m else np.array([])


class StrongSORT:
    def __init__(self, model_weights=None, device=0, fp16=False, max_dist=0.2, max_iou_dist=0.7, max_age=30, n_init=1,
                 nn_budget=100, mc_lambda=0.995, ema_alpha=0.9, model=None, max_tracks=256, max_dets=256, fused=True, **_capacity):
        self.device = _device_index(device)
        self._fused = None                                 # one-stream context of the batched step (default path)
        self._want_fused, self._max_dets = bool(fused), max_dets
        self.max_dist, self.max_iou_dist, self.max_age, self.n_init = max_dist, max_iou_dist, max_age, n_init
        self.nn_budget, self.mc_lambda, self.ema_alpha = nn_budget, mc_lambda, ema_alpha
        self.model = model
        self.tracks: list[_Track] = []
        self.samples: dict = {}                            # plain path only (see _gallery_ok)
        self._store = None                                 # device-resident gallery of the tensor-core path
        self._max_tracks = max_tracks
        self._next_id = 1
        _lib.load()
        _ops._torch()                                   # fail loudly without a CUDA device: there is no CPU path

    # ------------------------------------------------------------------ matching (linear_assignment.py:14-79)
    def _min_cost_matching(self, cost_fn, max_distance, track_idx, det_idx):
        if len(det_idx) == 0 or len(track_idx) == 0:
            return [], track_idx, det_idx
        cost = cost_fn(track_idx, det_idx)
        cost[cost > max_distance] = max_distance + 1e-5
        rows, cols = _ops.linear_sum_assignment(cost)
        rows, cols = rows.tolist(), cols.tolist()
        matches, ut, ud = [], [], []
        colset, rowset = set(cols), set(rows)
        for col, d in enumerate(det_idx):
            if col not in colset:
                ud.append(d)
        for row, t in enumerate(track_idx):
            if row not in rowset:
                ut.append(t)
        for row, col in zip(rows, cols):
            t, d = track_idx[row], det_idx[col]
            if cost[row, col] > max_distance:
                ut.append(t)
                ud.append(d)
            else:
                matches.append((t, d))
        return matches, ut, ud

    def _gallery_ok(self, dim):
        """The tensor-core gallery distance needs budget <= 128 rows per track, dim % 64 == 0 and a positive mc_lambda (its
        threshold is max_dist / mc_lambda: a larger cosine distance cannot survive the fused, clipped cost)."""
        return self.nn_budget is not None and 0 < self.nn_budget <= 128 and dim % 64 == 0 and self.mc_lambda > 0

    def update(self, dets, img, feats=None, warp=None):
        with _ops.on_device(self.device):
            return self._update(dets, img, feats, warp)

    def _update(self, dets, img, feats=None, warp=None):
        """`warp`: externally estimated 2x3 camera-motion matrix of this frame (what self.cmc.apply(img, xyxy) returns in the
        reference, strong_sort.py:63-65); None = identity.  Estimation itself (OpenCV ECC) is out of scope."""
        _SingleStreamTracker._check(dets)
        assert isinstance(img, np.ndarray) or img is None or isinstance(img, tuple), "Unsupported 'img' input format"
        dets = np.asarray(dets, dtype=np.float64)
        n = len(dets)
        if feats is None:
            if n and self.model is None:
                raise ValueError("StrongSORT needs `feats` or a `model` with get_features(xyxys, img)")
            feats = self.model.get_features(dets[:, 0:4], img) if n else np.zeros((0, 1), dtype=np.float32)
        # private copy: a new track normalises its row in place
        feats = np.array(feats, dtype=np.float32).reshape(n, -1) if n else np.zeros((0, 1), dtype=np.float32)
        tracks = self.tracks
        if self._fused is None and self._want_fused and not tracks and not self.samples and self._store is None:
            if n == 0:
                return np.array([])                         # nothing tracked, nothing seen
            if self._gallery_ok(feats.shape[1]) and self._max_tracks <= 256 and self._max_dets <= 256:
                self._fused = _FusedStream(self.device, self._max_tracks, self._max_dets, feats.shape[1], max_dist=self.max_dist,
                                           max_iou_dist=self.max_iou_dist, max_age=self.max_age, n_init=self.n_init,
                                           nn_budget=self.nn_budget, mc_lambda=self.mc_lambda, ema_alpha=self.ema_alpha)
        if self._fused is not None:
            return self._fused.step(dets, feats, warp)
        # Track.camera_update (track.py:129-138) - with the identity warp still not an exact no-op in floating point
        if tracks:
            moved = _ops.camera_update_xyah(np.stack([t.mean for t in tracks]), warp)
            for k, t in enumerate(tracks):
                t.mean = moved[k]
        use_store = n > 0 and n <= 256 and self._gallery_ok(feats.shape[1])
        if use_store and self._store is None and not self.samples:
            self._store = _ops.GalleryStore(self._max_tracks, self.nn_budget, feats.shape[1])
        use_store = use_store and self._store is not None
        tlwh = dets[:, :4].copy()
        tlwh[:, 2] = dets[:, 2] - dets[:, 0]
        tlwh[:, 3] = dets[:, 3] - dets[:, 1]
        xyah = tlwh.copy()
        xyah[:, :2] += xyah[:, 2:] / 2
        xyah[:, 2] /= xyah[:, 3]
        # Tracker.predict (tracker.py:59-66)
        if tracks:
            mean, cov = _ops.kf_predict(KIND, np.stack([t.mean for t in tracks]), np.stack([t.covariance for t in tracks]))
            for k, t in enumerate(tracks):
                t.mean, t.covariance = mean[k], cov[k]
                t.age += 1
                t.time_since_update += 1

        def gated_metric(track_idx, det_idx):
            if self._store is not None and use_store:
                # a cosine distance above max_dist / mc_lambda cannot survive the fused cost's clip at max_dist
                thr = self.max_dist / self.mc_lambda * (1.0 + 1e-12)
                cost = self._store.distance([tracks[k].id for k in track_idx], feats[det_idx], thr, thr + 1e-5)
            elif self._store is not None:
                raise RuntimeError("StrongSORT: a frame with more than 256 detections after the device-resident gallery was set up")
            else:
                cost = _ops.nn_cosine_distance([self.samples[tracks[k].id] for k in track_idx], feats[det_idx])
            cost = _ops.gate_cost(KIND, cost, np.stack([tracks[k].mean for k in track_idx]),
                                  np.stack([tracks[k].covariance for k in track_idx]), xyah[det_idx], False, fuse=True,
                                  lambda_=self.mc_lambda)
            cost[np.isinf(cost)] = INFTY_COST               # the reference gates with 1e5, not inf; both clip to the same value
            return cost

        def iou_metric(track_idx, det_idx):
            boxes_t = np.stack([tracks[k].to_tlbr() for k in track_idx])
            boxes_d = np.concatenate([tlwh[det_idx, :2], tlwh[det_idx, :2] + tlwh[det_idx, 2:]], axis=1)
            cost = _ops.iou_distance(boxes_t, boxes_d)
            for r, k in enumerate(track_idx):
                if tracks[k].time_since_update > 1:
                    cost[r, :] = INFTY_COST
            return cost

        # Tracker._match (tracker.py:104-155)
        confirmed = [i for i, t in enumerate(tracks) if t.state == CONFIRMED]
        unconfirmed = [i for i, t in enumerate(tracks) if t.state != CONFIRMED]
        m_a, _, ud = self._min_cost_matching(gated_metric, self.max_dist, list(confirmed), list(range(n)))
        ut_a = list(set(confirmed) - set(k for k, _ in m_a))            # linear_assignment.py:141 (CPython set order)
        cand = unconfirmed + [k for k in ut_a if tracks[k].time_since_update == 1]
        ut_a = [k for k in ut_a if tracks[k].time_since_update != 1]
        m_b, ut_b, ud = self._min_cost_matching(iou_metric, self.max_iou_dist, cand, ud)
        matches = m_a + m_b
        unmatched_tracks = list(set(ut_a + ut_b))

        # Tracker.update (tracker.py:73-102)
        if matches:
            ks = [k for k, _ in matches]
            ds = [d for _, d in matches]
            mean, cov = _ops.kf_update(KIND, np.stack([tracks[k].mean for k in ks]), np.stack([tracks[k].covariance for k in ks]),
                                       xyah[ds], dets[ds, 4])
            smooth = _ops.ema_unit_features(np.stack([tracks[k].feature for k in ks]), feats[ds], self.ema_alpha)   # track.py:166-172
            for i, (k, d) in enumerate(matches):
                t = tracks[k]
                t.mean, t.covariance = mean[i], cov[i]
                t.conf, t.cls, t.det_ind = dets[d, 4], dets[d, 5], float(d)
                t.feature = smooth[i]
                t.hits += 1
                t.time_since_update = 0
                if t.state == TENTATIVE and t.hits >= self.n_init:
                    t.state = CONFIRMED
        for k in unmatched_tracks:
            t = tracks[k]
            if t.state == TENTATIVE or t.time_since_update > self.max_age:
                t.state = DELETED
        if ud:
            mean, cov = _ops.kf_initiate(KIND, xyah[ud])
            first = _ops.unit_features(feats[ud])
            for i, d in enumerate(ud):
                t = _Track()
                t.id = self._next_id
                self._next_id += 1
                t.conf, t.cls, t.det_ind = dets[d, 4], dets[d, 5], float(d)
                t.hits, t.age, t.time_since_update, t.state = 1, 1, 0, TENTATIVE
                t.feature = first[i]
                t.mean, t.covariance = mean[i], cov[i]
                tracks.append(t)
        self.tracks = tracks = [t for t in tracks if t.state != DELETED]
        # NearestNeighborDistanceMetric.partial_fit (matching.py:343-358)
        active = [t.id for t in tracks if t.state == CONFIRMED]
        if self._store is None and active and self._gallery_ok(len(tracks[0].feature)) and not self.samples:
            self._store = _ops.GalleryStore(self._max_tracks, self.nn_budget, len(tracks[0].feature))
        if self._store is not None:
            self._store.append(active, np.stack([t.feature for t in tracks if t.state == CONFIRMED]) if active else np.zeros((0, 1)))
            self._store.keep_only(active)
        else:
            for t in tracks:
                if t.state == CONFIRMED:
                    g = self.samples.setdefault(t.id, [])
                    g.append(t.feature)
                    if self.nn_budget is not None:
                        self.samples[t.id] = g[-self.nn_budget:]
            self.samples = {k: self.samples[k] for k in active}
        rows = [np.concatenate((t.to_tlbr(), [t.id], [t.conf], [t.cls], [t.det_ind])).reshape(1, -1)
                for t in tracks if t.state == CONFIRMED and t.time_since_update < 1]
        return np.concatenate(rows) if rows else np.array([])

    def state(self):
        if self._fused is not None:
            return self._fused.batch.state(0)
        ts = self.tracks
        n = len(ts)
        return dict(track_id=np.array([t.id for t in ts], dtype=np.int32), state=np.array([t.state for t in ts], dtype=np.int32),
                    hits=np.array([t.hits for t in ts], dtype=np.int32), age=np.array([t.age for t in ts], dtype=np.int32),
                    time_since_update=np.array([t.time_since_update for t in ts], dtype=np.int32),
                    mean=np.stack([t.mean for t in ts]) if n else np.zeros((0, 8)),
                    cov=np.stack([t.covariance for t in ts]) if n else np.zeros((0, 8, 8)),
                    gallery=np.array([self._store.count(t.id) if self._store is not None else len(self.samples.get(t.id, []))
                                      for t in ts], dtype=np.int32),
                    feature=np.stack([t.feature for t in ts]) if n else np.zeros((0, 0), dtype=np.float32))
