import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def assert_close(a, b, rel=1e-9, abs_=1e-12, what=""):
    """Floating-point parity bar from BASELINE.json north_star: 1e-9 relative (abs floor for zeros)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    if a.size == 0:
        return
    err = np.abs(a - b)
    tol = abs_ + rel * np.maximum(np.abs(a), np.abs(b))
    bad = err > tol
    assert not bad.any(), f"{what}: max err {err.max():.3e} at {np.argwhere(bad)[:3].tolist()}"


def botsort_scenario(name):
    """(scenario dict, reference params, dets, ndets, seam features) of a BoT-SORT golden; the inputs are
    re-generated (tests/golden/scenarios.py) and checked against the checksum stored in the fixture."""
    import sys
    if GOLDEN not in sys.path:
        sys.path.insert(0, GOLDEN)
    from scenarios import BOTSORT_SCENARIOS, BOTSORT_YAML, botsort_inputs, camera_warps
    sc = dict(BOTSORT_SCENARIOS[name])
    sc["warps"] = camera_warps(sc) if sc.get("camera") else None
    cfg = dict(BOTSORT_YAML)
    cfg.update(sc["params"])
    dets, nd, _, feats = botsort_inputs(sc)
    g = load_golden(name)
    assert np.array_equal(nd, g["ndets"]) and np.allclose([dets.sum(), float(np.abs(feats).sum())], g["dets_sum"], rtol=1e-12), \
        "synthetic inputs drifted from the ones the golden was generated on"
    return sc, cfg, dets, nd, feats, g


def strongsort_scenario(name):
    import sys
    if GOLDEN not in sys.path:
        sys.path.insert(0, GOLDEN)
    from scenarios import STRONGSORT_SCENARIOS, STRONGSORT_YAML, camera_warps, strongsort_inputs
    sc = dict(STRONGSORT_SCENARIOS[name])
    sc["warps"] = camera_warps(sc) if sc.get("camera") else None
    cfg = dict(STRONGSORT_YAML)
    cfg.update(sc["params"])
    dets, nd, _, feats = strongsort_inputs(sc)
    g = load_golden(name)
    assert np.array_equal(nd, g["ndets"]) and np.allclose([dets.sum(), float(np.abs(feats).sum())], g["dets_sum"], rtol=1e-12), \
        "synthetic inputs drifted from the ones the golden was generated on"
    return sc, cfg, dets, nd, feats, g


def deepocsort_scenario(name):
    """(scenario, params, dets, ndets, per-frame seam features of the detections above det_thresh, golden)."""
    import sys
    if GOLDEN not in sys.path:
        sys.path.insert(0, GOLDEN)
    from scenarios import DEEPOCSORT_SCENARIOS, DEEPOCSORT_YAML, camera_warps, deepocsort_inputs
    sc = dict(DEEPOCSORT_SCENARIOS[name])
    sc["warps"] = camera_warps(sc) if sc.get("camera") else None
    cfg = dict(DEEPOCSORT_YAML)
    cfg.update(sc["params"])
    dets, nd, _, feats = deepocsort_inputs(sc, cfg["det_thresh"])
    g = load_golden(name)
    assert np.array_equal(nd, g["ndets"]) and np.allclose([dets.sum(), float(sum(np.abs(f).sum() for f in feats))], g["dets_sum"],
                                                          rtol=1e-12), "synthetic inputs drifted from the ones the golden was generated on"
    return sc, cfg, dets, nd, feats, g


def check_deepocsort_frame(name, f, out, s, g, heavy):
    """One frame of a DeepOCSORT replay against the live reference's golden: output rows, track records, filter state."""
    ref = g["out"][g["out_offs"][f]:g["out_offs"][f + 1]]
    assert out.reshape(-1, 8).shape == ref.shape, f"{name} frame {f}"
    if ref.size:
        assert np.array_equal(out[:, 4:], ref[:, 4:]), f"{name} frame {f}: id/conf/cls/det_ind"
        assert_close(out[:, :4], ref[:, :4], what=f"{name} frame {f} boxes")
    lo, hi = g["rec_offs"][f], g["rec_offs"][f + 1]
    mine = np.stack([s["track_id"], s["age"], s["time_since_update"], s["hits"], s["hit_streak"], s["observed"], s["frozen"]],
                    axis=1).reshape(-1, 7)
    assert np.array_equal(mine, g["rec"][lo:hi]), f"{name} frame {f}: track records"
    assert_close(s["x"], g["x"][lo:hi], what=f"{name} frame {f} x")
    assert_close(s["velocity"], g["vel"][lo:hi], what=f"{name} frame {f} velocity")
    assert_close(s["last_observation"], g["last"][lo:hi], what=f"{name} frame {f} last observation")
    if f in heavy:
        a, b = heavy[f]
        assert_close(s["P"].reshape(-1, 64), g["P"][a:b], abs_=1e-10, what=f"{name} frame {f} P")


def hybridsort_scenario(name, full=False):
    """(scenario, params, dets, ndets, per-frame seam features of the detections above det_thresh - `full`: of every
    detection, as the device path takes them -, golden)."""
    import sys
    if GOLDEN not in sys.path:
        sys.path.insert(0, GOLDEN)
    from scenarios import HYBRIDSORT_SCENARIOS, HYBRIDSORT_YAML, hybridsort_inputs
    sc = dict(HYBRIDSORT_SCENARIOS[name])
    cfg = dict(HYBRIDSORT_YAML)
    cfg.update(sc["params"])
    dets, nd, _, feats = hybridsort_inputs(sc, cfg["det_thresh"])
    g = load_golden(name)
    assert np.array_equal(nd, g["ndets"]) and np.allclose([dets.sum(), float(sum(np.abs(f).sum() for f in feats))], g["dets_sum"],
                                                          rtol=1e-12), "synthetic inputs drifted from the ones the golden was generated on"
    if full:
        feats = hybridsort_inputs(sc, cfg["det_thresh"], full=True)[3]
    return sc, cfg, dets, nd, feats, g


def per_class_calls(dets, tracker_classes):
    """The calls the reference's PerClassDecorator makes for one frame (boxmot/utils/__init__.py:33-52), as index arrays into
    `dets` in call order: one per class that has detections or live trackers, in the iteration order of the same set / dict
    expressions."""
    by_cls = {class_id: np.array([i for i, det in enumerate(dets) if det[5] == class_id], dtype=np.int64)
              for class_id in set(det[5] for det in dets)}
    relevant = set([np.float64(c) for c in tracker_classes]).union(set(by_cls.keys()))
    return [by_cls.get(int(class_id), np.zeros(0, dtype=np.int64)) for class_id in relevant]


def check_hybridsort_frame(name, f, out, s, g, heavy):
    """One frame of a HybridSORT replay against the live reference's golden: output rows (the last column is the SCORE of
    the input row the reference indexes, hybridsort.py:396), track records, 9-d filter state, four corner velocities."""
    ref = g["out"][g["out_offs"][f]:g["out_offs"][f + 1]]
    assert out.reshape(-1, 8).shape == ref.shape, f"{name} frame {f}"
    if ref.size:
        assert np.array_equal(out[:, 4:], ref[:, 4:]), f"{name} frame {f}: id/conf/cls/last column"
        assert_close(out[:, :4], ref[:, :4], what=f"{name} frame {f} boxes")
    lo, hi = g["rec_offs"][f], g["rec_offs"][f + 1]
    mine = np.stack([s["track_id"], s["age"], s["time_since_update"], s["hits"], s["hit_streak"], s["observed"]], axis=1).reshape(-1, 6)
    assert np.array_equal(mine, g["rec"][lo:hi]), f"{name} frame {f}: track records"
    assert_close(s["x"], g["x"][lo:hi], what=f"{name} frame {f} x")
    assert_close(s["velocity"].reshape(-1, 8), g["vel"][lo:hi], what=f"{name} frame {f} velocity")
    assert_close(s["last_observation"], g["last"][lo:hi], what=f"{name} frame {f} last observation")
    if f in heavy:
        a, b = heavy[f]
        assert_close(s["P"].reshape(-1, 81), g["P"][a:b], abs_=1e-10, what=f"{name} frame {f} P")


def heavy_offsets(g):
    offs, pos = {}, 0
    for f in g["heavy_frames"]:
        n = int(g["rec_offs"][f + 1] - g["rec_offs"][f])
        offs[int(f)] = (pos, pos + n)
        pos += n
    return offs


class OracleOps:
    """The operator namespace of yolo_tracking_b200._ops restated with the CPU oracle - TEST ONLY: it lets the `not gpu`
    suite replay the host-side list logic of an operator-backed tracker (monkeypatched in, never importable by the
    product) against the goldens; the GPU suite runs the same replay through the CUDA operators."""
    @staticmethod
    def _torch():
        return None

    @staticmethod
    def box_similarity(name, a, b, w=0.0, h=0.0):
        from oracle import boxes
        return boxes.similarity(name, np.asarray(a, dtype=np.float64).reshape(-1, 4), np.asarray(b, dtype=np.float64).reshape(-1, 4), w, h)

    @staticmethod
    def dot_matrix(a, b):
        return np.asarray(a, dtype=np.float64) @ np.asarray(b, dtype=np.float64).T

    @staticmethod
    def aw_max_metric(emb, w, bottom=0.5):
        from oracle.deepocsort import compute_aw_max_metric
        return compute_aw_max_metric(emb, w, bottom)

    @staticmethod
    def ocm_cost(sim, dets5=None, vel=None, prev5=None, inertia=0.0, emb=None):
        from oracle.lap import tie_break
        s = np.array(sim, dtype=np.float64)
        if dets5 is not None:
            cx_d, cy_d = (dets5[:, 0] + dets5[:, 2]) / 2.0, (dets5[:, 1] + dets5[:, 3]) / 2.0
            cx_p, cy_p = (prev5[:, 0] + prev5[:, 2]) / 2.0, (prev5[:, 1] + prev5[:, 3]) / 2.0
            dx, dy = cx_d[None, :] - cx_p[:, None], cy_d[None, :] - cy_p[:, None]
            norm = np.sqrt(dx ** 2 + dy ** 2) + 1e-6
            cosang = np.clip(vel[:, 1:2] * (dx / norm) + vel[:, 0:1] * (dy / norm), -1, 1)
            diff = (np.pi / 2.0 - np.abs(np.arccos(cosang))) / np.pi
            valid = (prev5[:, 4] >= 0).astype(np.float64)[:, None]
            s = s + ((valid * diff) * inertia).T * dets5[:, 4:5]
        if emb is not None:
            s = s + emb
        return tie_break(-s)

    @staticmethod
    def lapjv(cost, cost_limit=np.inf):
        from oracle.lap import lapjv_extended
        _, x, y = lapjv_extended(cost, cost_limit)
        return x, y

    @staticmethod
    def kf8_predict(mean, cov, unit_q=False):
        from oracle import deepocsort as o
        out = [o.kf8_predict(m, c, np.eye(8) if unit_q else o.process_noise(m[2], m[3])) for m, c in zip(mean, cov)]
        return np.stack([a for a, _ in out]), np.stack([b for _, b in out])

    @staticmethod
    def kf8_update(mean, cov, z, wh=None):
        from oracle import deepocsort as o
        out = [o.kf8_correct(m, c, zz, np.eye(4) if wh is None else o.measurement_noise(wh[i][0], wh[i][1]))
               for i, (m, c, zz) in enumerate(zip(mean, cov, z))]
        return np.stack([a for a, _ in out]), np.stack([b for _, b in out])

    @staticmethod
    def kf8_oru(mean, cov, box1, box2, gap):
        from oracle import deepocsort as o
        out = [o.kf8_virtual_trajectory(m, c, b1, b2, int(g)) for m, c, b1, b2, g in zip(mean, cov, box1, box2, gap)]
        return np.stack([a for a, _, _ in out]), np.stack([b for _, b, _ in out]), np.stack([v[-1] for _, _, v in out])

    @staticmethod
    def kf_apply_warp(mean, cov, warp, warp_index=None):
        from oracle import kalman
        return kalman.apply_warp(np.asarray(mean), np.asarray(cov), np.asarray(warp))


def deepocsort_known_answer(make_tracker):
    """The reference's own DeepOCSORT test (tests/test_python.py:51-95): two boxes in, two rows out in reversed order
    within atol=1 / rtol=7e-3; nothing before min_hits consecutive hits once frame_count > min_hits."""
    from numpy.testing import assert_allclose
    rgb = np.random.default_rng(0).integers(0, 255, size=(640, 640, 3), dtype=np.uint8)
    det = np.array([[144, 212, 578, 480, 0.82, 0], [425, 281, 576, 472, 0.56, 65]], dtype=np.float64)
    trk = make_tracker()
    trk.asso_func = "centroid"
    out = trk.update(det, rgb)
    assert out.shape == (2, 8)
    assert_allclose(det, np.flip(np.delete(out, [4, 7], axis=1), axis=0), atol=1, rtol=7e-3)
    trk = make_tracker()
    trk.min_hits = 2
    sizes = [trk.update(d, rgb).size for d in (np.empty((0, 6)), np.empty((0, 6)), det, det)]
    assert sizes == [0, 0, 0, 0]
    out = trk.update(det, rgb)
    assert out.shape == (2, 8)
    out = trk.update(det, rgb)
    assert out.shape == (2, 8)
    assert_allclose(det, np.flip(np.delete(out, [4, 7], axis=1), axis=0), atol=1, rtol=7e-3)


class RandomReID:
    """get_features seam with seeded random embeddings through the reference's whole-matrix normalisation."""
    def __init__(self, dim=32):
        self.rng, self.dim = np.random.default_rng(5), dim

    def get_features(self, xyxys, img):
        f = self.rng.normal(0, 1, (len(xyxys), self.dim)).astype(np.float32)
        return f / np.linalg.norm(f)


def deepocsort_edge_replay(make_tracker, make_oracle):
    """Edges of DeepOCSort.update against the oracle: appearance switched off, empty frames in the middle of a run,
    frames whose detections all fall under det_thresh, tracks that age out."""
    from yolo_tracking_b200.synth import make_stream
    dets, nd, embs = make_stream(4, 90, 9, 60, emb_dim=16, occlusion=True, miss_prob=0.1)
    for cfg in (dict(embedding_off=True), dict(det_thresh=0.45, max_age=4, min_hits=2, asso_func="diou"), dict(aw_off=True, asso_func="ciou")):
        base = dict(det_thresh=0, max_age=30, min_hits=1, iou_threshold=0.3, delta_t=3, asso_func="giou", inertia=0.2)
        base.update(cfg)
        trk, orc = make_tracker(**base), make_oracle(**base)
        for f in range(60):
            d = dets[f, :nd[f]].copy()
            if f in (7, 8, 30):
                d = np.empty((0, 6))
            if f in (15, 16, 17, 18, 19, 20):
                d[:, 4] = 0.01 * (1 + np.arange(len(d)))           # nothing passes a threshold of 0.45
            keep = d[:, 4] > base["det_thresh"]
            raw = embs[f, :len(d)][keep].astype(np.float32)
            feats = raw / np.linalg.norm(raw) if len(raw) else np.zeros((0, 16), dtype=np.float32)
            got = trk.update(d, (1080, 1920), feats=feats).reshape(-1, 8)
            ref = orc.update(d, feats).reshape(-1, 8)
            assert got.shape == ref.shape, (cfg, f)
            assert np.array_equal(got[:, 4:], ref[:, 4:]), (cfg, f)
            assert_close(got[:, :4], ref[:, :4], what=f"{cfg} frame {f}")
        assert [t.id for t in trk.trackers] == [t.id for t in orc.trackers]
