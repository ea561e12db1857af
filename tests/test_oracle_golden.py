"""The oracle (oracle/) against golden vectors produced by the live reference
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from _util import assert_close, load_golden
from oracle import boxes, kalman
from oracle.bytetrack import ByteTrackOracle
from oracle.lap import assign_no_limit, assign_with_limit, lapjv_extended


@pytest.mark.parametrize("kind", ["xyah", "xywh", "xyah_conf"])
def test_kalman_matches_reference(kind):
    g = load_golden("kf_" + kind)
    m, c = kalman.initiate(kind, g["z0"])
    assert_close(m, g["init_mean"], what="initiate mean")
    assert_close(c, g["init_cov"], what="initiate cov")
    steps = g["z"].shape[0]
    for s in range(steps):
        m, c = kalman.predict(kind, m, c)
        assert_close(m, g["pred_mean"][s], what=f"predict mean {s}")
        assert_close(c, g["pred_cov"][s], what=f"predict cov {s}")
        conf = g["conf"][s] if kind == "xyah_conf" else 0.0
        pm, pc = kalman.project(kind, m, c, conf)
        assert_close(pm, g["proj_mean"][s], what="project mean")
        assert_close(pc, g["proj_cov"][s], what="project cov")
        m, c = kalman.update(kind, m, c, g["z"][s], conf)
        assert_close(m, g["upd_mean"][s], what=f"update mean {s}")
        assert_close(c, g["upd_cov"][s], abs_=1e-10, what=f"update cov {s}")
    pm_, pc_ = g["pred_mean"][-1], g["pred_cov"][-1]
    for i in range(pm_.shape[0]):
        assert_close(kalman.gating_distance(kind, pm_[i], pc_[i], g["gate_meas"]), g["gate_maha4"][i], what="maha4")
        assert_close(kalman.gating_distance(kind, pm_[i], pc_[i], g["gate_meas"], True), g["gate_maha2"][i], what="maha2")
        if kind != "xyah_conf":
            assert_close(kalman.gating_distance(kind, pm_[i], pc_[i], g["gate_meas"], False, "gaussian"),
                         g["gate_gauss"][i], what="gauss")


def test_pairwise_costs_match_reference():
    g = load_golden("costs")
    a, b = g["a"], g["b"]
    for name in ("iou", "giou", "diou", "ciou"):
        assert np.array_equal(boxes.ASSO[name](a, b), g[name]), name      # same op order -> same bits
    assert np.array_equal(boxes.centroid(a, b, 640, 480), g["centroid"])
    cost = 1 - boxes.iou(a, b)
    assert np.array_equal(cost, g["iou_distance"])
    assert np.array_equal(1 - (1 - cost) * g["score"][None, :], g["fuse_score"])


def test_lapjv_known_answer():
    # distinguishes lapjv's extended-matrix optimum from "Hungarian then threshold"
    opt, x, y = lapjv_extended(np.array([[0.1, 0.79], [0.2, 2.0]]), 0.8)
    assert x.tolist() == [0, -1] and y.tolist() == [0, -1]
    m, ua, ub = assign_with_limit(np.array([[0.1, 0.79], [0.2, 2.0]]), 0.8)
    assert m.tolist() == [[0, 0]] and ua.tolist() == [1] and ub.tolist() == [1]
    m, ua, ub = assign_with_limit(np.zeros((0, 3)), 0.8)
    assert m.shape == (0, 2) and list(ua) == [] and list(ub) == [0, 1, 2]
    # no limit: every min(R, C) row is matched
    rng = np.random.default_rng(0)
    c = rng.random((5, 3))
    assert len(assign_no_limit(c)) == 3


def _replay(name):
    g = load_golden(name)
    p = g["params"]
    trk = ByteTrackOracle(p[0], p[1], int(p[2]), int(p[3]))
    dets, nd = g["dets"], g["ndets"]
    cov_frames = {int(f): k for k, f in enumerate(g["cov_frames"])}
    cov_pos = 0
    cov_offs = [0]
    for f in g["cov_frames"]:
        cov_offs.append(cov_offs[-1] + int(g["counts"][f].sum()))
    for f in range(dets.shape[0]):
        assert trk.track_updates == int(g["pool"][:f].sum())
        out = trk.update(dets[f, :nd[f]], None)
        ref = g["out"][g["out_offs"][f]:g["out_offs"][f + 1]]
        assert out.reshape(-1, 8).shape == ref.shape, f"frame {f}"
        if ref.size:
            assert np.array_equal(out[:, 4:], ref[:, 4:]), f"frame {f}: id/conf/cls/det_ind"
            assert_close(out[:, :4], ref[:, :4], what=f"frame {f} boxes")
        s = trk.snapshot()
        assert (int(s["n_tracked"]), int(s["n_lost"])) == tuple(g["counts"][f])
        rec = g["rec"][g["rec_offs"][f]:g["rec_offs"][f + 1]]
        mine = np.stack([s["track_id"], s["state"], s["is_activated"], s["frame_id"],
                         s["start_frame"], s["tracklet_len"]], axis=1).reshape(-1, 6)
        assert np.array_equal(mine, rec), f"frame {f}: lifecycle records"
        assert_close(s["mean"], g["mean"][g["rec_offs"][f]:g["rec_offs"][f + 1]], what=f"frame {f} mean")
        if f in cov_frames:
            k = cov_frames[f]
            assert_close(s["cov"].reshape(-1, 64), g["cov"][cov_offs[k]:cov_offs[k + 1]], abs_=1e-10,
                         what=f"frame {f} cov")


@pytest.mark.parametrize("name", ["bytetrack_c1", "bytetrack_churn"])
def test_bytetrack_oracle_replays_reference(name):
    _replay(name)


def test_bytetrack_reference_known_answer():
    # the reference's own test input (tests/test_python.py:165-185)
    g = load_golden("bytetrack_2box")
    trk = ByteTrackOracle(0.5, 0.8, 30, 30)
    for k in range(3):
        out = trk.update(g["det"], None)
        assert out.shape == (2, 8)
        assert_close(out, g["out"][k])
    assert_close(np.delete(out, [4, 7], axis=1), g["det"], rel=7e-3, abs_=1.0)


def test_bytetrack_empty_and_asserts():
    trk = ByteTrackOracle(0.5, 0.8, 30, 30)
    assert trk.update(np.empty((0, 6))).shape == (0,)
    with pytest.raises(AssertionError):
        trk.update(np.zeros((2, 5)))
    with pytest.raises(AssertionError):
        trk.update([[0, 0, 1, 1, 0.9, 0]])


# ----------------------------------------------------------------------------- OC-SORT
def _ocsort_from_params(p, **kw):
    from oracle.ocsort import OCSortOracle
    return OCSortOracle(False, det_thresh=p[0], max_age=int(p[1]), min_hits=int(p[2]), asso_threshold=p[3],
                        delta_t=int(p[4]), asso_func="giou", inertia=p[5], use_byte=bool(p[6]) if len(p) > 6 else False, **kw)


@pytest.mark.parametrize("name", ["ocsort_c2", "ocsort_churn", "ocsort_byte"])
def test_ocsort_oracle_replays_reference(name):
    g = load_golden(name)
    trk = _ocsort_from_params(g["params"])
    dets, nd = g["dets"], g["ndets"]
    hw = tuple(int(v) for v in g["img_hw"])
    heavy = {int(f): k for k, f in enumerate(g["heavy_frames"])}
    p_offs = [0]
    for f in g["heavy_frames"]:
        p_offs.append(p_offs[-1] + int(g["rec_offs"][f + 1] - g["rec_offs"][f]))
    for f in range(dets.shape[0]):
        out = trk.update(dets[f, :nd[f]], hw)
        ref = g["out"][g["out_offs"][f]:g["out_offs"][f + 1]]
        assert out.reshape(-1, 8).shape == ref.shape, f"frame {f}"
        if ref.size:
            assert np.array_equal(out[:, 4:], ref[:, 4:]), f"frame {f}: id/conf/cls/det_ind"
            assert_close(out[:, :4], ref[:, :4], what=f"frame {f} boxes")
        s = trk.snapshot()
        lo, hi = g["rec_offs"][f], g["rec_offs"][f + 1]
        mine = np.stack([s["track_id"], s["age"], s["time_since_update"], s["hits"], s["hit_streak"], s["observed"]],
                        axis=1).reshape(-1, 6)
        assert np.array_equal(mine, g["rec"][lo:hi]), f"frame {f}: lifecycle records"
        assert_close(s["x"], g["x"][lo:hi], what=f"frame {f} x")
        assert_close(s["velocity"], g["vel"][lo:hi], what=f"frame {f} velocity")
        assert_close(s["last_observation"], g["last"][lo:hi], what=f"frame {f} last_observation")
        if f in heavy:
            k = heavy[f]
            assert_close(s["P"].reshape(-1, 49), g["P"][p_offs[k]:p_offs[k + 1]], abs_=1e-9, what=f"frame {f} P")
    assert trk.stats["oru"] > 0 and trk.stats["lap_frames"] > 0
    if name == "ocsort_byte":
        assert trk.stats["byte_matches"] > 20


def test_ocsort_reference_known_answer():
    from oracle.ocsort import OCSortOracle
    g = load_golden("ocsort_2box")
    base = dict(det_thresh=0, max_age=30, min_hits=1, asso_threshold=0.3, delta_t=3, asso_func="giou", inertia=0.2)
    trk = OCSortOracle(False, **base)
    for k in range(3):
        out = trk.update(g["det"], (1080, 1920))
        assert out.shape == (2, 8)
        assert_close(out, g["out"][k])
    trk = OCSortOracle(False, **dict(base, min_hits=2))
    seq = [np.empty((0, 6)), np.empty((0, 6)), g["det"], np.empty((0, 6)), g["det"], g["det"], g["det"]]
    assert [trk.update(d, (1080, 1920)).size for d in seq] == g["min_hits_sizes"].tolist()


# ----------------------------------------------------------------------------- BoT-SORT
@pytest.mark.parametrize("name", ["botsort_c3", "botsort_churn", "botsort_noreid", "botsort_fuse", "botsort_cam"])
def test_botsort_oracle_replays_reference(name):
    from _util import botsort_scenario
    from oracle.botsort import BoTSORTOracle
    sc, cfg, dets, nd, feats, g = botsort_scenario(name)
    trk = BoTSORTOracle(**cfg)
    cov_frames = {int(f): k for k, f in enumerate(g["cov_frames"])}
    cov_offs = [0]
    for f in g["cov_frames"]:
        cov_offs.append(cov_offs[-1] + int(g["counts"][f].sum()))
    for f in range(sc["n_frames"]):
        out = trk.update(dets[f, :nd[f]], feats[f, :nd[f]], warp=None if sc["warps"] is None else sc["warps"][f])
        ref = g["out"][g["out_offs"][f]:g["out_offs"][f + 1]]
        assert out.reshape(-1, 8).shape == ref.shape, f"frame {f}"
        if ref.size:
            assert np.array_equal(out[:, 4:], ref[:, 4:]), f"frame {f}: id/conf/cls/det_ind"
            assert_close(out[:, :4], ref[:, :4], what=f"frame {f} boxes")
        s = trk.snapshot()
        assert (int(s["n_tracked"]), int(s["n_lost"])) == tuple(g["counts"][f])
        lo, hi = g["rec_offs"][f], g["rec_offs"][f + 1]
        mine = np.stack([s["track_id"], s["state"], s["is_activated"], s["frame_id"], s["start_frame"],
                         s["tracklet_len"]], axis=1).reshape(-1, 6)
        assert np.array_equal(mine, g["rec"][lo:hi]), f"frame {f}: lifecycle records"
        assert_close(s["mean"], g["mean"][lo:hi], what=f"frame {f} mean")
        assert np.array_equal(np.stack([s["score"], s["cls"], s["det_ind"]], axis=1).reshape(-1, 3), g["aux"][lo:hi])
        if f in cov_frames:
            k = cov_frames[f]
            assert_close(s["cov"].reshape(-1, 64), g["cov"][cov_offs[k]:cov_offs[k + 1]], abs_=1e-10, what=f"frame {f} cov")
    if cfg.get("with_reid", True):
        assert np.array_equal(s["smooth_feat"], g["final_feat"])       # same numpy float32 operations -> same bits
    if sc.get("classes"):
        assert len(np.unique(g["aux"][:, 1])) > 1


# ----------------------------------------------------------------------------- StrongSORT
@pytest.mark.parametrize("name", ["strongsort_c4", "strongsort_churn", "strongsort_cam"])
def test_strongsort_oracle_replays_reference(name):
    from _util import strongsort_scenario
    from oracle.strongsort import StrongSORTOracle
    sc, cfg, dets, nd, feats, g = strongsort_scenario(name)
    trk = StrongSORTOracle(**cfg)
    cov_frames = {int(f): k for k, f in enumerate(g["cov_frames"])}
    cov_offs = [0]
    for f in g["cov_frames"]:
        cov_offs.append(cov_offs[-1] + int(g["rec_offs"][f + 1] - g["rec_offs"][f]))
    for f in range(sc["n_frames"]):
        out = trk.update(dets[f, :nd[f]], feats[f, :nd[f]], warp=None if sc["warps"] is None else sc["warps"][f])
        ref = g["out"][g["out_offs"][f]:g["out_offs"][f + 1]]
        assert out.reshape(-1, 8).shape == ref.shape, f"frame {f}"
        if ref.size:
            assert np.array_equal(out[:, 4:], ref[:, 4:]), f"frame {f}: id/conf/cls/det_ind"
            assert_close(out[:, :4], ref[:, :4], what=f"frame {f} boxes")
        s = trk.snapshot()
        lo, hi = g["rec_offs"][f], g["rec_offs"][f + 1]
        mine = np.stack([s["track_id"], s["state"], s["hits"], s["age"], s["time_since_update"], s["gallery"]], axis=1).reshape(-1, 6)
        assert np.array_equal(mine, g["rec"][lo:hi]), f"frame {f}: track records"
        assert_close(s["mean"], g["mean"][lo:hi], what=f"frame {f} mean")
        if f in cov_frames:
            k = cov_frames[f]
            assert_close(s["cov"].reshape(-1, 64), g["cov"][cov_offs[k]:cov_offs[k + 1]], abs_=1e-10, what=f"frame {f} cov")
    assert np.array_equal(s["feature"], g["final_feat"])
    assert len(np.unique(g["rec"][:, 1])) >= 2


def test_camera_warp_and_adaptive_weight_match_reference():
    """STrack.multi_gmc and compute_aw_max_metric restatements against outputs of the live reference."""
    from oracle import deepocsort, kalman
    g = load_golden("aux_ops")
    for k, H in enumerate(g["warps"]):
        m, c = kalman.apply_warp(g["mean"], g["cov"], H)
        assert_close(m, g[f"gmc_mean{k}"], what=f"warp {k} mean")
        assert_close(c, g[f"gmc_cov{k}"], abs_=1e-10, what=f"warp {k} cov")
    assert np.array_equal(kalman.apply_warp(g["mean"], g["cov"], g["warps"][0])[0], g["mean"])        # identity is exact
    for k in range(4):
        assert_close(deepocsort.compute_aw_max_metric(g[f"aw_in{k}"], 0.75, 0.5), g[f"aw_out{k}"], what=f"aw {k}")
    assert_close(deepocsort.compute_aw_max_metric(g["aw_in0"], 0.4, 0.3), g["aw_out0_b"], what="aw 0 b")


# ----------------------------------------------------------------------------- DeepOCSORT
@pytest.mark.parametrize("name", ["deepocsort_c4", "deepocsort_churn", "deepocsort_cam", "deepocsort_noemb", "deepocsort_ciou"])
def test_deepocsort_oracle_replays_reference(name):
    """ids, hit / age counters, observed / frozen flags exact; 8-d filter state, velocities, last observations and boxes to
    1e-9 against the live reference, through occlusions (the quirky unfreeze), OCR and a moving camera."""
    from _util import check_deepocsort_frame, deepocsort_scenario, heavy_offsets
    from oracle.deepocsort import DeepOCSortOracle
    sc, cfg, dets, nd, feats, g = deepocsort_scenario(name)
    trk = DeepOCSortOracle(**cfg)
    heavy = heavy_offsets(g)
    for f in range(sc["n_frames"]):
        out = trk.update(dets[f, :nd[f]], feats[f], warp=None if sc["warps"] is None else sc["warps"][f])
        check_deepocsort_frame(name, f, out, trk.snapshot(), g, heavy)
    assert_close(trk.snapshot()["emb"], g["final_emb"], rel=1e-6, what="embeddings")
    assert trk.stats["oru"] > 50 and trk.stats["lap_frames"] > 20 and trk.stats["ocr_frames"] >= 1


# ----------------------------------------------------------------------------- HybridSORT
@pytest.mark.parametrize("name", ["hybridsort_c4", "hybridsort_churn", "hybridsort_diou"])
def test_hybridsort_oracle_replays_reference(name):
    """ids, hit / age counters, observed flags and the reference's odd last column exact; 9-d filter state, the four corner
    velocities, last observations and boxes to 1e-9, smoothed float32 embeddings to 2e-6 against the live reference, through
    occlusions (unfreeze with the score read as aspect ratio), the long-term-ReID correction and OCR."""
    from _util import check_hybridsort_frame, heavy_offsets, hybridsort_scenario
    from oracle.hybridsort import HybridSortOracle
    sc, cfg, dets, nd, feats, g = hybridsort_scenario(name)
    trk = HybridSortOracle(**cfg)
    heavy = heavy_offsets(g)
    for f in range(sc["n_frames"]):
        out = trk.update(dets[f, :nd[f]], feats[f])
        check_hybridsort_frame(name, f, out, trk.snapshot(), g, heavy)
    assert_close(trk.snapshot()["smooth_feat"], g["final_emb"], rel=2e-6, abs_=2e-6, what="embeddings")
    assert trk.stats["oru"] > 10 and trk.stats["corrections"] > 20
    assert name != "hybridsort_c4" or trk.stats["ocr_frames"] >= 1


def test_hybridsort_oracle_two_classes_through_the_per_class_wrapper():
    """The live reference ran `hybridsort_2cls` through its PerClassDecorator (per_class = True, hybridsort.py:346): one full
    update per class and frame.  The oracle restates the undecorated update; driven by the same calls it reproduces the rows
    and the track records."""
    from _util import check_hybridsort_frame, heavy_offsets, hybridsort_scenario, per_class_calls
    from oracle.hybridsort import HybridSortOracle
    name = "hybridsort_2cls"
    sc, cfg, dets, nd, feats, g = hybridsort_scenario(name, full=True)
    trk = HybridSortOracle(**cfg)
    heavy = heavy_offsets(g)
    calls = 0
    for f in range(sc["n_frames"]):
        d = dets[f, :nd[f]]
        out = np.empty((0, 8))
        if d.size:
            for idx in per_class_calls(d, [t.cls for t in trk.trackers]):
                keep = d[idx, 4] > cfg["det_thresh"]
                o = trk.update(d[idx].reshape(-1, 6), feats[f][idx][keep])
                calls += 1
                if o.size:
                    out = np.append(out, o.reshape(-1, 8), axis=0)
        else:
            out = trk.update(d, feats[f]).reshape(-1, 8)
            calls += 1
        check_hybridsort_frame(name, f, out, trk.snapshot(), g, heavy)
    assert calls > 1.8 * sc["n_frames"] and len(set(t.cls for t in trk.trackers)) == 2


def test_camera_warp_leaves_two_independent_4x4_blocks():
    """Structure the fused frame steps can rely on when they take a camera warp (DESIGN.md performance plan): on the live
    reference's moving-camera runs the covariance couples x with y (and w with h) but the (x, y, vx, vy) and (w, h, vw, vh)
    groups stay EXACTLY uncorrelated - kron(I4, R) maps each group to itself, F, H and the diagonal Q, R never mix them.
    So a warped 8-d filter is two 4x4 symmetric blocks (20 numbers), not a dense 8x8 (36)."""
    A, B = [0, 1, 4, 5], [2, 3, 6, 7]
    for name, key in (("botsort_cam", "cov"), ("deepocsort_cam", "P"), ("deepocsort_ciou", "P")):
        cov = load_golden(name)[key].reshape(-1, 8, 8)
        assert len(cov) > 100
        assert np.abs(cov[:, A][:, :, B]).max() == 0.0 and np.abs(cov[:, B][:, :, A]).max() == 0.0, name
        assert np.abs(cov[:, 0, 1]).max() > 0.1, name                  # the warp did break the 2x2 sparsity
