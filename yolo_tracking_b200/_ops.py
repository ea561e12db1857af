"""numpy-in / numpy-out wrappers over the operator-level C-ABI (b200track_kf_*, _box_similarity,
_iou_distance, _embedding_distance, _lapjv).  torch only provides the device buffers."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("yolo_tracking_b200 needs a CUDA device (there is no CPU fallback)")
    return torch


def _dev(a, dtype):
    """Host array -> tensor on the CURRENT CUDA device: the operator entry points launch on the current device, so callers
    that own a device (the drop-in trackers) run their operator calls inside `on_device(index)`."""
    torch = _torch()
    arr = np.ascontiguousarray(a, dtype=dtype)
    if not arr.flags.writeable:
        arr = arr.copy()
    return torch.from_numpy(arr).to("cuda")


def on_device(index):
    """Context manager: operator calls inside run on (and allocate on) CUDA device `index`."""
    return _torch().cuda.device(int(index))


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _sync_check(rc):
    _lib.check(rc)
    _torch().cuda.synchronize()


def kf_initiate(kind, z):
    lib = _lib.load()
    torch = _torch()
    z = np.asarray(z, dtype=np.float64).reshape(-1, 4)
    n = len(z)
    dz = _dev(z, np.float64)
    mean = torch.empty((n, 8), dtype=torch.float64, device=dz.device)
    cov = torch.empty((n, 8, 8), dtype=torch.float64, device=dz.device)
    _sync_check(lib.b200track_kf_initiate(kind, n, _p(dz), _p(mean), _p(cov), None))
    return mean.cpu().numpy(), cov.cpu().numpy()


def kf_predict(kind, mean, cov):
    lib = _lib.load()
    m = _dev(np.asarray(mean).reshape(-1, 8), np.float64)
    c = _dev(np.asarray(cov).reshape(-1, 8, 8), np.float64)
    _sync_check(lib.b200track_kf_predict(kind, m.shape[0], _p(m), _p(c), None))
    return m.cpu().numpy(), c.cpu().numpy()


def kf_apply_warp(mean, cov, warp, warp_index=None):
    """STrack.multi_gmc (bot_sort.py:95-111): warp [2, 3] (or [W, 2, 3] with warp_index [n]) applied to dense states."""
    lib = _lib.load()
    m = _dev(np.asarray(mean).reshape(-1, 8), np.float64)
    c = _dev(np.asarray(cov).reshape(-1, 8, 8), np.float64)
    w = _dev(np.asarray(warp, dtype=np.float64).reshape(-1, 6), np.float64)
    wi = _dev(np.asarray(warp_index, dtype=np.int32).reshape(-1), np.int32) if warp_index is not None else None
    _sync_check(lib.b200track_kf_apply_warp(m.shape[0], _p(m), _p(c), _p(w), _p(wi) if wi is not None else None, None))
    return m.cpu().numpy(), c.cpu().numpy()


def aw_max_metric(emb_cost, w_association_emb, bottom=0.5):
    """compute_aw_max_metric (association.py:79-108): emb_cost [R, C] or [B, R, C] float64."""
    lib = _lib.load()
    torch = _torch()
    e = np.asarray(emb_cost, dtype=np.float64)
    e3 = e.reshape((-1,) + e.shape[-2:])
    if e3.size == 0:
        return e.copy()
    de = _dev(e3, np.float64)
    out = torch.empty_like(de)
    _sync_check(lib.b200track_aw_max_metric(e3.shape[0], e3.shape[1], e3.shape[2], _p(de), float(w_association_emb), float(bottom), _p(out), None))
    return out.cpu().numpy().reshape(e.shape)


def kf_project(kind, mean, cov, conf=None):
    lib = _lib.load()
    torch = _torch()
    m = _dev(np.asarray(mean).reshape(-1, 8), np.float64)
    c = _dev(np.asarray(cov).reshape(-1, 8, 8), np.float64)
    n = m.shape[0]
    cf = _dev(np.broadcast_to(np.asarray(conf, dtype=np.float64), (n,)), np.float64) if conf is not None else None
    pm = torch.empty((n, 4), dtype=torch.float64, device=m.device)
    pc = torch.empty((n, 4, 4), dtype=torch.float64, device=m.device)
    _sync_check(lib.b200track_kf_project(kind, n, _p(m), _p(c), _p(cf), _p(pm), _p(pc), None))
    return pm.cpu().numpy(), pc.cpu().numpy()


def kf_update(kind, mean, cov, z, conf=None):
    lib = _lib.load()
    m = _dev(np.asarray(mean).reshape(-1, 8), np.float64)
    c = _dev(np.asarray(cov).reshape(-1, 8, 8), np.float64)
    n = m.shape[0]
    dz = _dev(np.asarray(z).reshape(-1, 4), np.float64)
    cf = _dev(np.broadcast_to(np.asarray(conf, dtype=np.float64), (n,)), np.float64) if conf is not None else None
    _sync_check(lib.b200track_kf_update(kind, n, _p(m), _p(c), _p(dz), _p(cf), None))
    return m.cpu().numpy(), c.cpu().numpy()


def kf_gating_distance(kind, mean, cov, meas, only_position=False, metric="maha", conf=None):
    if metric not in ("maha", "gaussian"):
        raise ValueError("invalid distance metric")
    lib = _lib.load()
    torch = _torch()
    m = _dev(np.asarray(mean).reshape(-1, 8), np.float64)
    c = _dev(np.asarray(cov).reshape(-1, 8, 8), np.float64)
    z = _dev(np.asarray(meas).reshape(-1, 4), np.float64)
    T, D = m.shape[0], z.shape[0]
    cf = _dev(np.broadcast_to(np.asarray(conf, dtype=np.float64), (T,)), np.float64) if conf is not None else None
    out = torch.empty((T, D), dtype=torch.float64, device=m.device)
    _sync_check(lib.b200track_kf_gating_distance(kind, T, D, _p(m), _p(c), _p(z), int(bool(only_position)),
                                                 0 if metric == "maha" else 1, _p(cf), _p(out), None))
    return out.cpu().numpy()


def gate_cost(kind, cost, mean, cov, meas, only_position=False, fuse=False, lambda_=0.98, conf=None):
    """cost [T, D] or [B, T, D] gated (and optionally fused) with the squared Mahalanobis distance, matching.py:170-196."""
    lib = _lib.load()
    torch = _torch()
    cost = np.asarray(cost, dtype=np.float64)
    single = cost.ndim == 2
    c3 = cost[None] if single else cost
    B, T, D = c3.shape
    m = _dev(np.asarray(mean).reshape(B, T, 8), np.float64)
    c = _dev(np.asarray(cov).reshape(B, T, 8, 8), np.float64)
    z = _dev(np.asarray(meas).reshape(B, D, 4), np.float64)
    cf = _dev(np.broadcast_to(np.asarray(conf, dtype=np.float64), (B, T)), np.float64) if conf is not None else None
    dc = _dev(c3, np.float64).clone()
    _sync_check(lib.b200track_gate_cost(kind, B, T, D, _p(m), _p(c), _p(z), int(bool(only_position)), int(bool(fuse)), float(lambda_),
                                        _p(cf), _p(dc), None))
    out = dc.cpu().numpy()
    return out[0] if single else out


def box_similarity(name, a, b, w=0.0, h=0.0):
    if name not in _lib.SIM:
        raise ValueError("Invalid function specified. Must be either '(g,d,c, )iou_batch' or 'centroid_batch'.")
    lib = _lib.load()
    torch = _torch()
    da = _dev(np.asarray(a, dtype=np.float64).reshape(-1, 4), np.float64)
    db = _dev(np.asarray(b, dtype=np.float64).reshape(-1, 4), np.float64)
    out = torch.empty((da.shape[0], db.shape[0]), dtype=torch.float64, device=da.device)
    _sync_check(lib.b200track_box_similarity(_lib.SIM[name], da.shape[0], db.shape[0], _p(da), _p(db), float(w), float(h),
                                             _p(out), None))
    return out.cpu().numpy()


def iou_distance(a, b, score=None):
    lib = _lib.load()
    torch = _torch()
    da = _dev(np.asarray(a, dtype=np.float64).reshape(-1, 4), np.float64)
    db = _dev(np.asarray(b, dtype=np.float64).reshape(-1, 4), np.float64)
    ds = _dev(score, np.float64) if score is not None else None
    out = torch.empty((da.shape[0], db.shape[0]), dtype=torch.float64, device=da.device)
    _sync_check(lib.b200track_iou_distance(da.shape[0], db.shape[0], _p(da), _p(db), _p(ds), _p(out), None))
    return out.cpu().numpy()


def embedding_distance(a, b):
    lib = _lib.load()
    torch = _torch()
    da = _dev(np.asarray(a), np.float32)
    db = _dev(np.asarray(b), np.float32)
    out = torch.empty((da.shape[0], db.shape[0]), dtype=torch.float64, device=da.device)
    _sync_check(lib.b200track_embedding_distance(da.shape[0], db.shape[0], da.shape[1], _p(da), _p(db), _p(out), None))
    return out.cpu().numpy()


def appearance_cost(trk, det, scale=0.5, thresh=0.25, fill=1.0, gate=None, return_stats=False):
    """Thresholded cosine cost of `batch` streams: trk [B, T, F], det [B, D, F] (fp32) -> out [B, T, D] fp64, where
    out = fill if gated or scale * max(0, cosine distance) > thresh, else that value exactly.  tcgen05 bf16
    pre-filter + fp64 re-evaluation of the survivors (include/b200track.h: b200track_appearance_cost)."""
    lib = _lib.load()
    torch = _torch()
    trk = np.asarray(trk, dtype=np.float32)
    det = np.asarray(det, dtype=np.float32)
    single = trk.ndim == 2
    if single:
        trk, det = trk[None], det[None]
        gate = None if gate is None else np.asarray(gate)[None]
    B, T, F = trk.shape
    D = det.shape[1]
    dt, dd = _dev(trk, np.float32), _dev(det, np.float32)
    dg = _dev(np.asarray(gate) != 0, np.uint8) if gate is not None else None
    nbytes = C.c_uint64()
    _lib.check(lib.b200track_appearance_cost_workspace(B, T, D, F, C.byref(nbytes)))
    ws = torch.empty((max(int(nbytes.value), 1),), dtype=torch.uint8, device=dt.device)
    stats = torch.zeros((2,), dtype=torch.int64, device=dt.device)
    out = torch.full((B, T, D), float("nan"), dtype=torch.float64, device=dt.device)
    _sync_check(lib.b200track_appearance_cost(B, T, D, F, _p(dt), _p(dd), _p(dg), float(scale), float(thresh), float(fill),
                                              _p(out), _p(ws), int(nbytes.value), _p(stats), None))
    st = stats.cpu().numpy()
    if st[1]:
        raise RuntimeError(f"appearance_cost: {int(st[1])} tile(s) hit an internal pipeline error")
    res = out.cpu().numpy()
    res = res[0] if single else res
    return (res, int(st[0])) if return_stats else res


def gallery_cost(gallery, count, det, thresh=0.2, fill=None, return_stats=False, resident_bf16=False):
    """Thresholded gallery distance of StrongSORT for a batch of streams (matching.py:247-378 + linear_assignment.py:59-78):
    gallery [B, T, G, F] float32, count [B, T], det [B, D, F] -> cost [B, T, D] float64 (tensor-core pre-filter, exact values)."""
    lib = _lib.load()
    torch = _torch()
    gal = np.asarray(gallery, dtype=np.float32)
    dt = np.asarray(det, dtype=np.float32)
    B, T, G, F = gal.shape
    D = dt.shape[1]
    fill = thresh + 1e-5 if fill is None else fill
    if B * T * D == 0:
        return np.zeros((B, T, D))
    dg, dc, dd = _dev(gal, np.float32), _dev(np.asarray(count).reshape(B, T), np.int32), _dev(dt, np.float32)
    # resident_bf16: the unit-norm bf16 copy of the gallery is built once by the caller (tracker state) instead of per call
    g16 = None
    if resident_bf16:
        g16 = torch.empty((B, T, G, F), dtype=torch.bfloat16, device=dg.device)
        _lib.check(lib.b200track_unit_bf16(B * T * G, F, _p(dg), _p(g16), None))
    need = C.c_uint64()
    _lib.check(lib.b200track_gallery_cost_workspace(B, T, G, D, F, 0 if resident_bf16 else 1, C.byref(need)))
    ws = torch.empty((max(int(need.value), 1),), dtype=torch.uint8, device=dg.device)
    out = torch.empty((B, T, D), dtype=torch.float64, device=dg.device)
    st = torch.zeros((3,), dtype=torch.int64, device=dg.device)
    _sync_check(lib.b200track_gallery_cost(B, T, G, D, F, _p(dg), _p(g16), _p(dc), _p(dd), float(thresh), float(fill), _p(out), _p(ws),
                                           int(need.value), _p(st), None))
    st = st.cpu().numpy()
    if st[1]:
        raise RuntimeError("gallery_cost: tensor-core pipeline protocol error")
    res = out.cpu().numpy()
    return (res, st) if return_stats else res


def nn_cosine_distance(galleries, det_feats):
    """galleries: list (one per track) of [n_t, F] float32 arrays; det_feats [D, F] -> cost [T, D] float64
    (NearestNeighborDistanceMetric.distance, matching.py:360-378)."""
    lib = _lib.load()
    torch = _torch()
    T, det = len(galleries), np.asarray(det_feats, dtype=np.float32)
    D = det.shape[0]
    if T == 0 or D == 0:
        return np.zeros((T, D))
    seg = np.zeros(T + 1, dtype=np.int32)
    seg[1:] = np.cumsum([len(g) for g in galleries])
    gal = _dev(np.concatenate([np.asarray(g, dtype=np.float32).reshape(-1, det.shape[1]) for g in galleries], axis=0), np.float32)
    dseg, ddet = _dev(seg, np.int32), _dev(det, np.float32)
    out = torch.empty((T, D), dtype=torch.float64, device=ddet.device)
    _sync_check(lib.b200track_nn_cosine_distance(T, D, det.shape[1], _p(gal), _p(dseg), _p(ddet), _p(out), None))
    return out.cpu().numpy()


def linear_sum_assignment(cost):
    """scipy.optimize.linear_sum_assignment (minimisation), bit-faithful including ties: cost [R, C] -> (row_ind, col_ind),
    or [B, R, C] -> list of such pairs."""
    lib = _lib.load()
    torch = _torch()
    cost = np.asarray(cost, dtype=np.float64)
    single = cost.ndim == 2
    c3 = cost[None] if single else cost
    B, R, Cc = c3.shape
    n = min(R, Cc)
    res = []
    if n == 0:
        res = [(np.empty(0, dtype=np.int64), np.empty(0, dtype=np.int64)) for _ in range(B)]
    else:
        dc = _dev(c3, np.float64)
        out = torch.full((B, n), -1, dtype=torch.int32, device=dc.device)
        err = torch.zeros((1,), dtype=torch.int32, device=dc.device)
        _sync_check(lib.b200track_linear_sum_assignment(B, R, Cc, _p(dc), _p(out), _p(err), None))
        if int(err.item()):
            raise ValueError("cost matrix is infeasible")
        c4r = out.cpu().numpy().astype(np.int64)
        for b in range(B):
            if Cc < R:                                    # scipy solved the transposed problem
                order = np.argsort(c4r[b], kind="stable")
                res.append((c4r[b][order], order))
            else:
                res.append((np.arange(R, dtype=np.int64), c4r[b]))
    return res[0] if single else res


def lapjv(cost, cost_limit=np.inf):
    """cost [R, C] or [B, R, C] -> x [.., R], y [.., C] (int32, -1 = unmatched)."""
    lib = _lib.load()
    torch = _torch()
    cost = np.asarray(cost, dtype=np.float64)
    single = cost.ndim == 2
    c3 = cost[None] if single else cost
    B, R, Cc = c3.shape
    dc = _dev(c3, np.float64)
    x = torch.full((B, max(R, 1)), -1, dtype=torch.int32, device=dc.device)
    y = torch.full((B, max(Cc, 1)), -1, dtype=torch.int32, device=dc.device)
    _sync_check(lib.b200track_lapjv(B, R, Cc, _p(dc) if R * Cc else None, float(cost_limit), _p(x), _p(y), None))
    x, y = x.cpu().numpy()[:, :R], y.cpu().numpy()[:, :Cc]
    return (x[0], y[0]) if single else (x, y)


# ---------------------------------------------------------------------------------- DeepOCSORT operators (csrc/kf8.cu)
def kf8_predict(mean, cov, unit_q=False):
    """kf.predict(Q=new_kf_process_noise(w, h)) of DeepOCSORT's 8-d filter (deep_ocsort.py:76-80, :263-266)."""
    lib = _lib.load()
    m = _dev(np.asarray(mean).reshape(-1, 8), np.float64)
    c = _dev(np.asarray(cov).reshape(-1, 8, 8), np.float64)
    _sync_check(lib.b200track_kf8_predict(m.shape[0], _p(m), _p(c), int(bool(unit_q)), None))
    return m.cpu().numpy(), c.cpu().numpy()


def kf8_update(mean, cov, z, wh=None):
    """kf.update(z, R=new_kf_measurement_noise(w, h)) (deepocsort_kf.py:549-563); wh [n, 2] or None (R = I)."""
    lib = _lib.load()
    m = _dev(np.asarray(mean).reshape(-1, 8), np.float64)
    c = _dev(np.asarray(cov).reshape(-1, 8, 8), np.float64)
    dz = _dev(np.asarray(z).reshape(-1, 4), np.float64)
    dwh = _dev(np.asarray(wh).reshape(-1, 2), np.float64) if wh is not None else None
    _sync_check(lib.b200track_kf8_update(m.shape[0], _p(m), _p(c), _p(dz), _p(dwh), None))
    return m.cpu().numpy(), c.cpu().numpy()


def kf8_oru(mean, cov, box1, box2, gap):
    """KalmanFilter.unfreeze's virtual trajectory (deepocsort_kf.py:433-478) -> mean, cov, last virtual box [n, 4]."""
    lib = _lib.load()
    torch = _torch()
    m = _dev(np.asarray(mean).reshape(-1, 8), np.float64)
    c = _dev(np.asarray(cov).reshape(-1, 8, 8), np.float64)
    b1 = _dev(np.asarray(box1).reshape(-1, 4), np.float64)
    b2 = _dev(np.asarray(box2).reshape(-1, 4), np.float64)
    g = _dev(np.asarray(gap).reshape(-1), np.int32)
    last = torch.empty((m.shape[0], 4), dtype=torch.float64, device=m.device)
    _sync_check(lib.b200track_kf8_oru(m.shape[0], _p(m), _p(c), _p(b1), _p(b2), _p(g), _p(last), None))
    return m.cpu().numpy(), c.cpu().numpy(), last.cpu().numpy()


def ocm_cost(sim, dets5=None, vel=None, prev5=None, inertia=0.0, emb=None):
    """-(sim + velocity-direction term + emb) + the canonical tie-break (association.py:130-172); sim [D, T]."""
    lib = _lib.load()
    torch = _torch()
    sim = np.asarray(sim, dtype=np.float64)
    D, T = sim.shape
    if D == 0 or T == 0:
        return np.zeros((D, T))
    ds = _dev(sim, np.float64)
    dd = _dev(np.asarray(dets5).reshape(D, 5), np.float64) if dets5 is not None else None
    dv = _dev(np.asarray(vel).reshape(T, 2), np.float64) if dets5 is not None else None
    dp = _dev(np.asarray(prev5).reshape(T, 5), np.float64) if dets5 is not None else None
    de = _dev(np.asarray(emb).reshape(D, T), np.float64) if emb is not None else None
    out = torch.empty_like(ds)
    _sync_check(lib.b200track_ocm_cost(D, T, _p(dd), _p(dv), _p(dp), float(inertia), _p(ds), _p(de), _p(out), None))
    return out.cpu().numpy()


def dot_matrix(a, b):
    """a @ b.T with fp64 accumulation (deep_ocsort.py:433): a [n, dim], b [m, dim] -> [n, m]."""
    lib = _lib.load()
    torch = _torch()
    da = _dev(np.asarray(a, dtype=np.float64), np.float64)
    db = _dev(np.asarray(b, dtype=np.float64), np.float64)
    out = torch.empty((da.shape[0], db.shape[0]), dtype=torch.float64, device=da.device)
    if out.numel():
        _sync_check(lib.b200track_dot_matrix(da.shape[0], db.shape[0], da.shape[1], _p(da), _p(db), _p(out), None))
    return out.cpu().numpy()


def ema_unit_features(trk, det, alpha):
    """Track.update's feature smoothing (strongsort/sort/track.py:166-172) for n rows, float32:
    unit(alpha * trk + (1 - alpha) * unit(det))."""
    lib = _lib.load()
    a = _dev(np.asarray(trk, dtype=np.float32), np.float32)
    b = _dev(np.asarray(det, dtype=np.float32), np.float32)
    n, F = a.shape
    _sync_check(lib.b200track_ema_unit_features(n, F, _p(a), _p(b), float(alpha), None))
    return a.cpu().numpy()


def unit_features(rows):
    """rows / |row| in float32 (a new StrongSORT track's first feature)."""
    lib = _lib.load()
    a = _dev(np.asarray(rows, dtype=np.float32), np.float32)
    _sync_check(lib.b200track_unit_features(a.shape[0], a.shape[1], _p(a), None))
    return a.cpu().numpy()


def camera_update_xyah(mean, warp=None):
    """Track.camera_update (strongsort/sort/track.py:129-138) on xyah means [n, 8]; warp [2, 3] or None (identity)."""
    lib = _lib.load()
    m = _dev(np.asarray(mean, dtype=np.float64).reshape(-1, 8), np.float64)
    w = _dev(np.asarray(warp, dtype=np.float64).reshape(6), np.float64) if warp is not None else None
    _sync_check(lib.b200track_camera_update_xyah(m.shape[0], _p(m), _p(w), None))
    return m.cpu().numpy()


class GalleryStore:
    """Device-resident StrongSORT gallery (NearestNeighborDistanceMetric.samples, matching.py:311-378): per slot the last
    `budget` features of a track as fp32 rows and as unit-norm bf16 rows, the operand formats of the tensor-core gallery
    distance (b200track_gallery_cost: tcgen05 pre-filter, exact float32 values).  The host keeps which slot belongs to which
    track and how many rows it holds; features are appended on the device (b200track_gallery_append)."""

    def __init__(self, n_slots, budget, dim):
        torch = _torch()
        self.lib = _lib.load()
        self.n_slots, self.budget, self.dim = int(n_slots), int(budget), int(dim)
        self.gal32 = torch.zeros((1, self.n_slots, self.budget, self.dim), dtype=torch.float32, device="cuda")
        self.gal16 = torch.zeros((1, self.n_slots, self.budget, self.dim), dtype=torch.bfloat16, device="cuda")
        self.appended = np.zeros(self.n_slots, dtype=np.int64)         # rows ever appended per slot
        self.slot_of = {}                                              # track id -> slot
        self.free = list(range(self.n_slots - 1, -1, -1))
        self._ws = None

    def count(self, track_id):
        s = self.slot_of.get(track_id)
        return 0 if s is None else int(min(self.appended[s], self.budget))

    def append(self, track_ids, rows):
        """One new feature row per listed track (a slot is allocated on a track's first row)."""
        if not len(track_ids):
            return
        slots, pos = [], []
        for tid in track_ids:
            s = self.slot_of.get(tid)
            if s is None:
                if not self.free:
                    raise RuntimeError(f"more than {self.n_slots} confirmed tracks (max_tracks)")
                s = self.slot_of[tid] = self.free.pop()
                self.appended[s] = 0
            slots.append(s)
            pos.append(int(self.appended[s] % self.budget))
            self.appended[s] += 1
        d_rows = _dev(np.asarray(rows, dtype=np.float32).reshape(len(slots), self.dim), np.float32)
        d_slot, d_pos = _dev(slots, np.int32), _dev(pos, np.int32)
        _sync_check(self.lib.b200track_gallery_append(len(slots), self.dim, self.budget, _p(d_rows), _p(d_slot), _p(d_pos),
                                                      _p(self.gal32), _p(self.gal16), None))

    def keep_only(self, track_ids):
        """samples = {k: samples[k] for k in active_targets}: slots of every other track are released."""
        keep = set(track_ids)
        for tid in [t for t in self.slot_of if t not in keep]:
            self.free.append(self.slot_of.pop(tid))

    def distance(self, track_ids, det, thresh, fill):
        """min over a track's rows of 1 - cos(row, det) where that is <= thresh, else `fill`: [len(track_ids), D] float64."""
        torch = _torch()
        det = np.asarray(det, dtype=np.float32)
        D = det.shape[0]
        if not len(track_ids) or D == 0:
            return np.zeros((len(track_ids), D))
        cnt = np.zeros(self.n_slots, dtype=np.int32)
        for tid, s in self.slot_of.items():
            cnt[s] = min(self.appended[s], self.budget)
        dc, dd = _dev(cnt.reshape(1, -1), np.int32), _dev(det.reshape(1, D, self.dim), np.float32)
        need = C.c_uint64()
        _lib.check(self.lib.b200track_gallery_cost_workspace(1, self.n_slots, self.budget, D, self.dim, 0, C.byref(need)))
        if self._ws is None or self._ws.numel() < int(need.value):
            self._ws = torch.empty((max(int(need.value), 1),), dtype=torch.uint8, device="cuda")
        out = torch.empty((1, self.n_slots, D), dtype=torch.float64, device="cuda")
        st = torch.zeros((3,), dtype=torch.int64, device="cuda")
        _sync_check(self.lib.b200track_gallery_cost(1, self.n_slots, self.budget, D, self.dim, _p(self.gal32), _p(self.gal16), _p(dc),
                                                    _p(dd), float(thresh), float(fill), _p(out), _p(self._ws), int(need.value),
                                                    _p(st), None))
        if int(st[1]):
            raise RuntimeError("gallery_cost: tensor-core pipeline protocol error")
        full = out[0].cpu().numpy()
        return full[[self.slot_of[t] for t in track_ids]]


# ---- OC-SORT's XYSR filter at operator level (csrc/kf_xysr.cu) -------------------------------------------------------
def _xysr(mode, x, P, z=None, last_z=None, gap=None):
    lib = _lib.load()
    torch = _torch()
    dx = _dev(np.asarray(x, dtype=np.float64).reshape(-1, 7), np.float64)
    dP = _dev(np.asarray(P, dtype=np.float64).reshape(-1, 7, 7), np.float64)
    n = dx.shape[0]
    err = torch.zeros((1,), dtype=torch.int32, device=dx.device)
    vl = None
    if mode == 0:
        _sync_check(lib.b200track_kf_xysr_predict(n, _p(dx), _p(dP), _p(err), None))
    else:
        dz = _dev(np.asarray(z, dtype=np.float64).reshape(n, 4), np.float64)
        if mode == 1:
            _sync_check(lib.b200track_kf_xysr_update(n, _p(dx), _p(dP), _p(dz), _p(err), None))
        else:
            dl = _dev(np.asarray(last_z, dtype=np.float64).reshape(n, 4), np.float64)
            dg = _dev(np.asarray(gap).reshape(n), np.int32)
            vl = torch.zeros((n, 4), dtype=torch.float64, device=dx.device)
            _sync_check(lib.b200track_kf_xysr_unfreeze_update(n, _p(dx), _p(dP), _p(dl), _p(dg), _p(dz), _p(vl), _p(err), None))
    if int(err.item()):
        raise ValueError("covariance does not have the structure of OC-SORT's XYSR filter")
    out = (dx.cpu().numpy(), dP.cpu().numpy())
    return out + (vl.cpu().numpy(),) if vl is not None else out


def kf_xysr_predict(x, P):
    """KalmanFilter.predict of OC-SORT's 7-d filter (ocsort_kf.py:339-379, ocsort.py:79-106): x [n, 7], P [n, 7, 7]."""
    return _xysr(0, x, P)


def kf_xysr_update(x, P, z):
    """KalmanFilter.update(z) (ocsort_kf.py:437-526), z [n, 4] = [x, y, s, r]."""
    return _xysr(1, x, P, z)


def kf_xysr_unfreeze_update(x_saved, P_saved, last_z, gap, z):
    """update(z) on a frozen filter: the virtual trajectory of unfreeze() (ocsort_kf.py:383-434) from the saved state, then
    the real measurement; returns (x, P, last virtual box)."""
    return _xysr(2, x_saved, P_saved, z, last_z, gap)


# ---- HybridSORT's score-carrying filter at operator level (csrc/kf_hybrid.cu) ------------------------------------------
def _xyscr(mode, x, P, z=None, last_z=None, gap=None):
    lib = _lib.load()
    torch = _torch()
    dx = _dev(np.asarray(x, dtype=np.float64).reshape(-1, 9), np.float64)
    dP = _dev(np.asarray(P, dtype=np.float64).reshape(-1, 9, 9), np.float64)
    n = dx.shape[0]
    err = torch.zeros((1,), dtype=torch.int32, device=dx.device)
    vl = None
    if mode == 0:
        _sync_check(lib.b200track_kf_xyscr_predict(n, _p(dx), _p(dP), _p(err), None))
    else:
        dz = _dev(np.asarray(z, dtype=np.float64).reshape(n, 5), np.float64)
        if mode == 1:
            _sync_check(lib.b200track_kf_xyscr_update(n, _p(dx), _p(dP), _p(dz), _p(err), None))
        else:
            dl = _dev(np.asarray(last_z, dtype=np.float64).reshape(n, 5), np.float64)
            dg = _dev(np.asarray(gap).reshape(n), np.int32)
            vl = torch.zeros((n, 5), dtype=torch.float64, device=dx.device)
            _sync_check(lib.b200track_kf_xyscr_unfreeze_update(n, _p(dx), _p(dP), _p(dl), _p(dg), _p(dz), _p(vl), _p(err), None))
    if int(err.item()):
        raise ValueError("covariance does not have the structure of HybridSORT's filter")
    out = (dx.cpu().numpy(), dP.cpu().numpy())
    return out + (vl.cpu().numpy(),) if vl is not None else out


def kf_xyscr_predict(x, P):
    """KalmanFilter.predict of HybridSORT's 9-d filter (hybridsort_kf.py:339-379, hybridsort.py:126-150): x [n, 9], P [n, 9, 9]."""
    return _xyscr(0, x, P)


def kf_xyscr_update(x, P, z):
    """KalmanFilter.update(z) (hybridsort_kf.py:439-528), z [n, 5] = [x, y, s, score, r]."""
    return _xyscr(1, x, P, z)


def kf_xyscr_unfreeze_update(x_saved, P_saved, last_z, gap, z):
    """update(z) on a frozen filter: the virtual trajectory of unfreeze() (hybridsort_kf.py:390-436) from the saved state,
    then the real measurement; returns (x, P, last virtual box)."""
    return _xyscr(2, x_saved, P_saved, z, last_z, gap)
