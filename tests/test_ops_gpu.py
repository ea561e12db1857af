"""GPU parity of the operator-level C-ABI (Kalman filters, box costs, cosine distance, lapjv)
against the golden vectors of the live reference and against the oracle."""
import numpy as np
import pytest

from _util import assert_close, load_golden

pytestmark = pytest.mark.gpu

KINDS = {"xyah": 0, "xywh": 1, "xyah_conf": 2}


@pytest.mark.parametrize("kind", ["xyah", "xywh", "xyah_conf"])
def test_kalman_ops_match_reference(kind):
    from yolo_tracking_b200 import _ops
    g = load_golden("kf_" + kind)
    k = KINDS[kind]
    m, c = _ops.kf_initiate(k, g["z0"])
    assert_close(m, g["init_mean"], what="initiate mean")
    assert_close(c, g["init_cov"], what="initiate cov")
    for s in range(g["z"].shape[0]):
        m, c = _ops.kf_predict(k, m, c)
        assert_close(m, g["pred_mean"][s], what=f"predict mean {s}")
        assert_close(c, g["pred_cov"][s], what=f"predict cov {s}")
        conf = g["conf"][s] if kind == "xyah_conf" else None
        pm, pc = _ops.kf_project(k, m, c, conf)
        assert_close(pm, g["proj_mean"][s], what="project mean")
        assert_close(pc, g["proj_cov"][s], what="project cov")
        m, c = _ops.kf_update(k, m, c, g["z"][s], conf)
        assert_close(m, g["upd_mean"][s], what=f"update mean {s}")
        assert_close(c, g["upd_cov"][s], abs_=1e-10, what=f"update cov {s}")
    pm_, pc_ = g["pred_mean"][-1], g["pred_cov"][-1]
    conf0 = np.zeros(len(pm_)) if kind == "xyah_conf" else None
    assert_close(_ops.kf_gating_distance(k, pm_, pc_, g["gate_meas"], False, "maha", conf0), g["gate_maha4"], what="maha4")
    assert_close(_ops.kf_gating_distance(k, pm_, pc_, g["gate_meas"], True, "maha", conf0), g["gate_maha2"], what="maha2")
    if kind != "xyah_conf":
        assert_close(_ops.kf_gating_distance(k, pm_, pc_, g["gate_meas"], False, "gaussian"), g["gate_gauss"], what="gauss")


def test_kalman_ops_dense_covariance_vs_oracle():
    """A dense (not block-sparse) SPD covariance, as after an external camera-motion warp."""
    from oracle import kalman
    from yolo_tracking_b200 import _ops
    rng = np.random.default_rng(3)
    n = 200
    A = rng.normal(size=(n, 8, 8))
    cov = A @ np.transpose(A, (0, 2, 1)) + 8 * np.eye(8)
    mean = rng.normal(size=(n, 8)) * 10 + np.array([500, 400, 0.5, 120, 0, 0, 0, 0])
    z = mean[:, :4] + rng.normal(size=(n, 4))
    for kind in ("xyah", "xywh"):
        k = KINDS[kind]
        pm, pc = _ops.kf_predict(k, mean, cov)
        om, oc = kalman.predict(kind, mean, cov)
        assert_close(pm, om); assert_close(pc, oc)
        um, uc = _ops.kf_update(k, om, oc, z)
        qm, qc = kalman.update(kind, om, oc, z)
        assert_close(um, qm, rel=1e-9, abs_=1e-9); assert_close(uc, qc, rel=1e-9, abs_=1e-9)
        d = _ops.kf_gating_distance(k, om[:5], oc[:5], z[:50])
        for i in range(5):
            assert_close(d[i], kalman.gating_distance(kind, om[i], oc[i], z[:50]), rel=1e-9, abs_=1e-9)


def test_box_costs_bit_exact():
    from yolo_tracking_b200 import _ops
    g = load_golden("costs")
    a, b = g["a"], g["b"]
    for name in ("iou", "giou", "diou"):
        assert np.array_equal(_ops.box_similarity(name, a, b), g[name]), name      # same op order -> same bits
    assert_close(_ops.box_similarity("ciou", a, b), g["ciou"], rel=1e-12)          # atan differs by <= 1-2 ulp
    assert np.array_equal(_ops.box_similarity("centroid", a, b, 640, 480), g["centroid"])
    assert np.array_equal(_ops.iou_distance(a, b), g["iou_distance"])
    assert np.array_equal(_ops.iou_distance(a, b, g["score"]), g["fuse_score"])
    with pytest.raises(ValueError):
        _ops.box_similarity("nosuch", a, b)


def test_embedding_distance():
    from yolo_tracking_b200 import _ops
    g = load_golden("costs")
    got = _ops.embedding_distance(g["feat_a"], g["feat_b"])
    assert_close(got, g["embedding_distance"], rel=1e-9, abs_=1e-12)


def test_lapjv_known_answer_and_shapes():
    from yolo_tracking_b200 import _ops
    from yolo_tracking_b200.utils import association, matching
    x, y = _ops.lapjv(np.array([[0.1, 0.79], [0.2, 2.0]]), 0.8)
    assert x.tolist() == [0, -1] and y.tolist() == [0, -1]
    m, ua, ub = matching.linear_assignment(np.array([[0.1, 0.79], [0.2, 2.0]]), 0.8)
    assert m.tolist() == [[0, 0]] and list(ua) == [1] and list(ub) == [1]
    m, ua, ub = matching.linear_assignment(np.zeros((0, 3)), 0.8)
    assert m.shape == (0, 2) and list(ua) == [] and list(ub) == [0, 1, 2]
    assert association.linear_assignment(np.zeros((0, 3))).shape == (0, 2)


@pytest.mark.parametrize("rows,cols,limit,density", [
    (60, 50, 0.8, 0.05), (200, 200, 0.8, 0.02), (200, 180, 0.5, 0.2), (37, 91, 0.7, 1.0),
    (64, 64, np.inf, 1.0), (200, 150, np.inf, 1.0), (90, 200, np.inf, 1.0), (1, 1, 0.5, 1.0), (5, 1, np.inf, 1.0),
])
def test_lapjv_random_vs_oracle(rows, cols, limit, density):
    from oracle.lap import lapjv_extended
    from yolo_tracking_b200 import _ops
    rng = np.random.default_rng(rows * 1000 + cols)
    B = 6
    cost = rng.random((B, rows, cols))
    if density < 1.0:                       # most pairs are far apart: cost 1.0 > limit, like iou_distance
        cost = np.where(rng.random((B, rows, cols)) < density, cost, 1.0)
    x, y = _ops.lapjv(cost, limit)
    for b in range(B):
        _, ox, oy = lapjv_extended(cost[b], limit)
        assert np.array_equal(x[b], ox), f"problem {b}: x"
        assert np.array_equal(y[b], oy), f"problem {b}: y"


def test_lapjv_negative_costs_no_limit():
    """OCSORT calls lapjv on -(iou + angle) without a limit (association.py:171-172)."""
    from oracle.lap import lapjv_extended
    from yolo_tracking_b200 import _ops
    rng = np.random.default_rng(77)
    cost = -(rng.random((4, 120, 130)) * 1.5)
    x, y = _ops.lapjv(cost)
    for b in range(4):
        _, ox, oy = lapjv_extended(cost[b])
        assert np.array_equal(x[b], ox) and np.array_equal(y[b], oy)


@pytest.mark.parametrize("rows,cols", [(27, 34), (34, 27), (150, 200), (200, 150)])
def test_lapjv_no_limit_mostly_zero_costs(rows, cols):
    """OC-SORT-shaped matrices: exact zeros (disjoint boxes) almost everywhere, one strong pair per
    object plus a few weak ones, canonical tie-break (oracle/lap.py "Ties").  Free columns must end
    with a zero dual, which a column-reduction start gets wrong."""
    from oracle.lap import lapjv_extended, tie_break
    from yolo_tracking_b200 import _ops
    rng = np.random.default_rng(rows * 7 + cols)
    B = 5
    cost = np.zeros((B, rows, cols))
    for b in range(B):
        k = min(rows, cols) - 3
        rr, cc = rng.permutation(rows)[:k], rng.permutation(cols)[:k]
        cost[b, rr, cc] = -rng.uniform(0.3, 1.0, k)
        extra = rng.random((rows, cols)) < 0.03
        cost[b] = np.where(extra & (cost[b] == 0), -rng.uniform(0.0, 0.6, (rows, cols)), cost[b])
        cost[b] = tie_break(cost[b])
    x, y = _ops.lapjv(cost)
    for b in range(B):
        _, ox, oy = lapjv_extended(cost[b])
        # the tie-break is separable in (row, column): it fixes WHICH rows / columns stay unmatched and
        # the real pairs, while exactly tied zero-cost pairs may be permuted among themselves
        assert np.array_equal(x[b] >= 0, ox >= 0) and np.array_equal(y[b] >= 0, oy >= 0), f"problem {b}: matched sets"
        r = np.nonzero(ox >= 0)[0]
        assert all(y[b][x[b][i]] == i for i in r)
        strong = cost[b][r, ox[r]] < -1e-6
        assert np.array_equal(x[b][r][strong], ox[r][strong]), f"problem {b}: real pairs"
        assert abs(cost[b][r, x[b][r]].sum() - cost[b][r, ox[r]].sum()) < 1e-12, f"problem {b}: objective"


def _exact_appearance(trk, det, scale, thresh, fill, gate=None):
    from scipy.spatial.distance import cdist
    out = np.empty((trk.shape[0], trk.shape[1], det.shape[1]))
    for b in range(trk.shape[0]):
        e = scale * np.maximum(0.0, cdist(trk[b].astype(np.float32), det[b].astype(np.float32), "cosine"))   # matching.py:156-166
        e[e > thresh] = fill
        if gate is not None:
            e[gate[b] != 0] = fill
        out[b] = e
    return out


@pytest.mark.parametrize("B,T,D,F,scale,thresh,fill", [
    (3, 100, 90, 512, 0.5, 0.25, 1.0),            # BoT-SORT default appearance_thresh
    (2, 200, 200, 512, 0.5, 0.4818211117541298, 1.0),   # botsort.yaml: the threshold sits inside the bulk of unrelated pairs
    (2, 130, 37, 128, 1.0, 0.2, 0.2 + 1e-5),      # StrongSORT max_dist = 0.2 -> max_dist + 1e-5; ragged tile edges
    (1, 300, 260, 256, 0.5, 0.3, 1.0),            # more than one tile in both directions
    (4, 1, 1, 64, 0.5, 0.25, 1.0),
])
def test_appearance_cost_tensor_core_prefilter_is_exact(B, T, D, F, scale, thresh, fill):
    """tcgen05 bf16 pre-filter + fp64 re-check == the reference's thresholded cosine cost (1e-9 relative)."""
    from yolo_tracking_b200 import _ops
    rng = np.random.default_rng(B * 1000 + T)
    proto = rng.standard_normal((B, max(T, D), F))
    trk = (proto[:, :T] + 0.3 * rng.standard_normal((B, T, F))).astype(np.float32) * rng.uniform(0.5, 2.0, (B, T, 1)).astype(np.float32)
    det = (proto[:, :D] + 0.3 * rng.standard_normal((B, D, F))).astype(np.float32)
    gate = rng.random((B, T, D)) < 0.2
    for gt in (None, gate):
        out, n_exact = _ops.appearance_cost(trk, det, scale, thresh, fill, gate=gt, return_stats=True)
        ref = _exact_appearance(trk, det, scale, thresh, fill, gt)
        assert np.array_equal(out == fill, ref == fill), "threshold decisions"
        assert_close(out, ref, what="appearance cost")
        assert n_exact < B * T * D or T * D <= 4           # the pre-filter did discard work
        assert (ref != fill).sum() <= n_exact


def test_gating_distance_batched_equals_per_stream_calls():
    """b200track_kf_gating_distance_batched == one b200track_kf_gating_distance call per stream == the oracle."""
    import ctypes as C
    import torch
    from oracle import kalman
    from yolo_tracking_b200 import _lib, _ops
    lib = _lib.load()
    rng = np.random.default_rng(5)
    B, T, D = 3, 21, 17
    z = np.stack([rng.uniform(100, 1800, B * T), rng.uniform(100, 1000, B * T), rng.uniform(0.3, 0.8, B * T),
                  rng.uniform(60, 220, B * T)], axis=1)
    mean, cov = kalman.initiate("xyah", z)
    mean, cov = kalman.predict("xyah", mean, cov)
    meas = mean.reshape(B, T, 8)[:, rng.integers(0, T, D), :4] + rng.normal(0, 3, (B, D, 4))
    dm, dc, dz = (torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (mean, cov, meas))
    out = torch.empty((B, T, D), dtype=torch.float64, device="cuda")
    p = lambda t: C.c_void_p(t.data_ptr())
    _lib.check(lib.b200track_kf_gating_distance_batched(_lib.KF_XYAH, B, T, D, p(dm), p(dc), p(dz), 0, 0, None, p(out), None))
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    for b in range(B):
        one = _ops.kf_gating_distance(_lib.KF_XYAH, mean[b * T:(b + 1) * T], cov[b * T:(b + 1) * T], meas[b])
        assert np.array_equal(got[b], one)
        for t in range(T):
            assert_close(got[b, t], kalman.gating_distance("xyah", mean[b * T + t], cov[b * T + t], meas[b]), what="maha")


@pytest.mark.parametrize("only_position", [False, True])
def test_gate_cost_matrix_and_fuse_motion(only_position):
    """matching.py:170-196 (gate_cost_matrix, fuse_motion) through the fused kernel vs the oracle's gating_distance."""
    from oracle import kalman
    from yolo_tracking_b200.motion.kalman_filters import KalmanFilterXYAH, chi2inv95
    from yolo_tracking_b200.utils import matching
    rng = np.random.default_rng(9)
    T, D = 37, 23
    z = np.stack([rng.uniform(100, 1800, T), rng.uniform(100, 1000, T), rng.uniform(0.3, 0.8, T), rng.uniform(60, 220, T)], axis=1)
    mean, cov = kalman.initiate("xyah", z)
    mean, cov = kalman.predict("xyah", mean, cov)
    meas = mean[rng.integers(0, T, D), :4] + rng.normal(0, 4, (D, 4)) * np.array([1, 1, 0.005, 1])
    cost = rng.random((T, D))
    gd = np.stack([kalman.gating_distance("xyah", mean[t], cov[t], meas, only_position) for t in range(T)])
    thr = chi2inv95[2 if only_position else 4]
    ref_gate = np.where(gd > thr, np.inf, cost)
    ref_fuse = 0.98 * ref_gate + (1 - 0.98) * gd

    class Trk:
        def __init__(self, m, c): self.mean, self.covariance = m, c

    class Det:
        def __init__(self, v): self.v = v
        def to_xyah(self): return self.v
    kf = KalmanFilterXYAH()
    trks, dets = [Trk(mean[t], cov[t]) for t in range(T)], [Det(m) for m in meas]
    got = matching.gate_cost_matrix(kf, cost.copy(), trks, dets, only_position)
    assert np.array_equal(np.isinf(got), np.isinf(ref_gate)) and 0 < np.isinf(ref_gate).sum() < T * D
    assert np.array_equal(got[np.isfinite(got)], ref_gate[np.isfinite(ref_gate)])
    got = matching.fuse_motion(kf, cost.copy(), (mean, cov), meas, only_position)
    assert np.array_equal(np.isinf(got), np.isinf(ref_fuse))
    assert_close(got[np.isfinite(got)], ref_fuse[np.isfinite(ref_fuse)], what="fuse_motion")
    assert matching.gate_cost_matrix(kf, np.zeros((0, 3)), [], dets[:3]).shape == (0, 3)


def test_linear_sum_assignment_is_bit_faithful_to_scipy_including_ties():
    """StrongSORT's solver (strongsort/sort/linear_assignment.py:59-61): clipped matrices are mostly one repeated value,
    and scipy's tie behaviour decides the order of the unmatched lists.  Same (row_ind, col_ind) as scipy on tie-heavy,
    constant, integer, tall and wide matrices."""
    from scipy.optimize import linear_sum_assignment as scipy_lsa
    from yolo_tracking_b200 import _ops
    rng = np.random.default_rng(0)
    for shape in [(1, 1), (3, 7), (7, 3), (12, 12), (40, 55), (55, 40), (130, 200), (200, 130)]:
        mats = []
        for kind in range(4):
            c = rng.random((3,) + shape)
            if kind == 1:
                c[c > 0.25] = 0.25 + 1e-5                  # min_cost_matching's clipping
            elif kind == 2:
                c = rng.integers(0, 3, (3,) + shape).astype(np.float64)
            elif kind == 3:
                c = np.full((3,) + shape, 0.7)
            mats.append(c)
        cost = np.concatenate(mats, axis=0)
        got = _ops.linear_sum_assignment(cost)
        for b in range(len(cost)):
            a, c = scipy_lsa(cost[b])
            assert np.array_equal(got[b][0], a) and np.array_equal(got[b][1], c), (shape, b)
    a, c = _ops.linear_sum_assignment(np.zeros((0, 4)))
    assert len(a) == 0 and len(c) == 0
    with pytest.raises(ValueError):
        _ops.linear_sum_assignment(np.full((2, 2), np.inf))


def test_camera_warp_operator_matches_reference():
    """b200track_kf_apply_warp against STrack.multi_gmc of the live reference (golden), one warp for all tracks and a
    per-track choice of warp."""
    from yolo_tracking_b200 import _ops
    g = load_golden("aux_ops")
    for k, H in enumerate(g["warps"]):
        m, c = _ops.kf_apply_warp(g["mean"], g["cov"], H)
        assert_close(m, g[f"gmc_mean{k}"], what=f"warp {k} mean")
        assert_close(c, g[f"gmc_cov{k}"], abs_=1e-10, what=f"warp {k} cov")
    idx = np.arange(len(g["mean"])) % 3
    m, c = _ops.kf_apply_warp(g["mean"], g["cov"], g["warps"], warp_index=idx)
    for k in range(3):
        assert_close(m[idx == k], g[f"gmc_mean{k}"][idx == k], what="indexed warp mean")
        assert_close(c[idx == k], g[f"gmc_cov{k}"][idx == k], abs_=1e-10, what="indexed warp cov")
    m0, c0 = _ops.kf_apply_warp(g["mean"], g["cov"], np.eye(2, 3))
    assert np.array_equal(m0, g["mean"]) and np.array_equal(c0, g["cov"])                  # identity warp is exact


def test_adaptive_weight_operator_matches_reference():
    """b200track_aw_max_metric against compute_aw_max_metric of the live reference: zero rows / columns, tied top-2,
    single row / column, and a batch."""
    from yolo_tracking_b200 import _ops
    g = load_golden("aux_ops")
    for k in range(4):
        assert_close(_ops.aw_max_metric(g[f"aw_in{k}"], 0.75, 0.5), g[f"aw_out{k}"], what=f"aw {k}")
    assert_close(_ops.aw_max_metric(g["aw_in0"], 0.4, 0.3), g["aw_out0_b"], what="aw 0 b")
    batch = np.stack([g["aw_in3"], g["aw_in3"][::-1].copy()])
    out = _ops.aw_max_metric(batch, 0.75, 0.5)
    assert_close(out[0], g["aw_out3"], what="batched aw")
    from oracle.deepocsort import compute_aw_max_metric
    assert_close(out[1], compute_aw_max_metric(batch[1], 0.75, 0.5), what="batched aw 1")


@pytest.mark.parametrize("B, T, G, D, F", [(2, 5, 100, 37, 128), (3, 9, 128, 200, 512), (1, 3, 7, 256, 64)])
def test_gallery_cost_tensor_core_operator(B, T, G, D, F):
    """b200track_gallery_cost (tcgen05 pre-filter + exact float32 re-evaluation) against the reference arithmetic of
    NearestNeighborDistanceMetric (matching.py:247-267, float32 numpy) followed by the clip of min_cost_matching: the same
    entries survive the threshold, surviving values agree to float32 rounding; ragged / empty galleries."""
    from yolo_tracking_b200 import _ops
    rng = np.random.default_rng(B * 1000 + T)
    proto = rng.standard_normal((B, T, F)).astype(np.float32)
    gal = (proto[:, :, None, :] + 0.25 * rng.standard_normal((B, T, G, F))).astype(np.float32)
    count = rng.integers(1, G + 1, (B, T)).astype(np.int32)
    count[0, 0] = 0                                         # a track without stored features: nothing matches
    count[-1, -1] = G
    owner = rng.integers(0, T, (B, D))
    det = (proto[np.arange(B)[:, None], owner] + 0.45 * rng.standard_normal((B, D, F))).astype(np.float32)
    thresh = 0.2
    got, st = _ops.gallery_cost(gal, count, det, thresh, return_stats=True)
    ref = np.full((B, T, D), thresh + 1e-5)
    raw = np.full((B, T, D), np.inf)
    for b in range(B):
        bn = det[b] / np.linalg.norm(det[b], axis=1, keepdims=True)
        for t in range(T):
            if count[b, t] == 0:
                continue
            a = gal[b, t, :count[b, t]]
            a = a / np.linalg.norm(a, axis=1, keepdims=True)
            raw[b, t] = (1.0 - np.dot(a, bn.T)).min(axis=0)
    keep = raw <= thresh
    ref[keep] = raw[keep]
    assert np.abs(raw - thresh).min() > 1e-5, "test data too close to the threshold for a float32 comparison"
    assert np.array_equal(got <= thresh, keep)
    assert keep.sum() > 0 and (~keep).sum() > 0
    assert np.allclose(got, ref, rtol=0, atol=2e-6)
    assert st[2] >= keep.sum()                              # every surviving pair went through the exact path
    # the caller-resident bf16 gallery (tracker state, converted once) gives the same matrix
    assert np.array_equal(_ops.gallery_cost(gal, count, det, thresh, resident_bf16=True), got)


def test_xysr_filter_operators_match_the_oracle_filter():
    """b200track_kf_xysr_predict / _update / _unfreeze_update (OC-SORT's 7-d filter at operator level, ocsort_kf.py:339-526)
    against oracle/ocsort.py's filter object through occlusion gaps: every filter gets the same measurements as its twin."""
    from oracle.ocsort import _KF
    from yolo_tracking_b200 import _ops
    rng = np.random.default_rng(11)
    n, frames = 40, 30
    z0 = np.stack([rng.uniform(100, 1800, n), rng.uniform(100, 900, n), rng.uniform(2000, 20000, n), rng.uniform(0.3, 0.8, n)], axis=1)
    kfs = [_KF(z0[i]) for i in range(n)]
    x = np.stack([k.x for k in kfs]); P = np.stack([k.P for k in kfs])
    observed = np.zeros(n, dtype=bool)
    saved = [None] * n
    last_z = z0.copy()
    gap = np.zeros(n, dtype=np.int64)
    n_oru = 0
    for f in range(frames):
        for k in kfs:
            k.predict()
        x, P = _ops.kf_xysr_predict(x, P)
        seen = rng.random(n) > 0.35
        z = np.stack([x[:, 0] + rng.normal(0, 3, n), x[:, 1] + rng.normal(0, 3, n), np.abs(x[:, 2]) * rng.uniform(0.9, 1.1, n) + 50,
                      rng.uniform(0.3, 0.8, n)], axis=1)
        for i, k in enumerate(kfs):
            k.update(z[i] if seen[i] else None)
        gap += 1
        # the operator calls a tracker would make: freeze = keep a copy; frozen + seen = unfreeze_update; else update
        for i in np.nonzero(~seen)[0]:
            if observed[i]:
                saved[i] = (x[i].copy(), P[i].copy())
            observed[i] = False
        thaw = np.array([i for i in np.nonzero(seen)[0] if not observed[i] and saved[i] is not None], dtype=np.int64)
        plain = np.array([i for i in np.nonzero(seen)[0] if i not in set(thaw.tolist())], dtype=np.int64)
        if len(plain):
            x[plain], P[plain] = _ops.kf_xysr_update(x[plain], P[plain], z[plain])
            last_z[plain] = z[plain]
        if len(thaw):
            xs = np.stack([saved[i][0] for i in thaw]); Ps = np.stack([saved[i][1] for i in thaw])
            x[thaw], P[thaw], vl = _ops.kf_xysr_unfreeze_update(xs, Ps, last_z[thaw], gap[thaw], z[thaw])
            last_z[thaw] = vl
            n_oru += len(thaw)
            for i in thaw:
                saved[i] = None
        observed[seen] = True
        gap[seen] = 0
        assert_close(x, np.stack([k.x for k in kfs]), what=f"frame {f} x")
        assert_close(P.reshape(n, 49), np.stack([k.P for k in kfs]).reshape(n, 49), abs_=1e-9, what=f"frame {f} P")
        assert_close(last_z, np.stack([k.last_z if k.last_z is not None else z0[i] for i, k in enumerate(kfs)]), what=f"frame {f} last z")
    assert n_oru > 20
    bad = P.copy(); bad[0, 0, 1] = 1.0
    with pytest.raises(ValueError):
        _ops.kf_xysr_predict(x, bad)


def test_hybrid_filter_operators_match_the_oracle_filter():
    """b200track_kf_xyscr_predict / _update / _unfreeze_update (HybridSORT's 9-d score-carrying filter at operator level,
    hybridsort_kf.py:339-528) against oracle/hybridsort.py's filter object through occlusion gaps: every filter gets the same
    measurements as its twin (the unfreeze reads the score as the aspect ratio, like the reference)."""
    from oracle.hybridsort import _KF
    from yolo_tracking_b200 import _ops
    rng = np.random.default_rng(12)
    n, frames = 40, 30
    z0 = np.stack([rng.uniform(100, 1800, n), rng.uniform(100, 900, n), rng.uniform(2000, 20000, n), rng.uniform(0.3, 0.95, n),
                   rng.uniform(0.3, 0.8, n)], axis=1)
    kfs = [_KF(z0[i]) for i in range(n)]
    x = np.stack([k.x for k in kfs]); P = np.stack([k.P for k in kfs])
    observed = np.zeros(n, dtype=bool)
    saved = [None] * n
    last_z = z0.copy()
    gap = np.zeros(n, dtype=np.int64)
    n_oru = 0
    for f in range(frames):
        for k in kfs:
            k.predict()
        x, P = _ops.kf_xyscr_predict(x, P)
        seen = rng.random(n) > 0.35
        z = np.stack([x[:, 0] + rng.normal(0, 3, n), x[:, 1] + rng.normal(0, 3, n), np.abs(x[:, 2]) * rng.uniform(0.9, 1.1, n) + 50,
                      rng.uniform(0.3, 0.95, n), rng.uniform(0.3, 0.8, n)], axis=1)
        for i, k in enumerate(kfs):
            k.update(z[i] if seen[i] else None)
        gap += 1
        for i in np.nonzero(~seen)[0]:
            if observed[i]:
                saved[i] = (x[i].copy(), P[i].copy())
            observed[i] = False
        thaw = np.array([i for i in np.nonzero(seen)[0] if not observed[i] and saved[i] is not None], dtype=np.int64)
        plain = np.array([i for i in np.nonzero(seen)[0] if i not in set(thaw.tolist())], dtype=np.int64)
        if len(plain):
            x[plain], P[plain] = _ops.kf_xyscr_update(x[plain], P[plain], z[plain])
            last_z[plain] = z[plain]
        if len(thaw):
            xs = np.stack([saved[i][0] for i in thaw]); Ps = np.stack([saved[i][1] for i in thaw])
            x[thaw], P[thaw], vl = _ops.kf_xyscr_unfreeze_update(xs, Ps, last_z[thaw], gap[thaw], z[thaw])
            last_z[thaw] = vl
            n_oru += len(thaw)
            for i in thaw:
                saved[i] = None
        observed[seen] = True
        gap[seen] = 0
        assert_close(x, np.stack([k.x for k in kfs]), what=f"frame {f} x")
        assert_close(P.reshape(n, 81), np.stack([k.P for k in kfs]).reshape(n, 81), abs_=1e-9, what=f"frame {f} P")
        assert_close(last_z, np.stack([k.last_z if k.last_z is not None else z0[i] for i, k in enumerate(kfs)]), what=f"frame {f} last z")
    assert n_oru > 20
    bad = P.copy(); bad[0, 0, 1] = 1.0
    with pytest.raises(ValueError):
        _ops.kf_xyscr_predict(x, bad)
