"""The reference-shaped Python API (create_tracker / tracker.update) on the GPU, including the
reference's own known-answer test (tests/test_python.py:165-185 of the reference)."""
import numpy as np
import pytest
from numpy.testing import assert_allclose

pytestmark = pytest.mark.gpu


def test_bytetrack_output_like_reference_test():
    from yolo_tracking_b200 import create_tracker, get_tracker_config
    tracker = create_tracker(tracker_type="bytetrack", tracker_config=get_tracker_config("bytetrack"),
                             reid_weights=None, device="cuda:0", half=False, per_class=False)
    rgb = np.random.randint(255, size=(640, 640, 3), dtype=np.uint8)
    det = np.array([[144, 212, 578, 480, 0.82, 0], [425, 281, 576, 472, 0.86, 65]])
    for _ in range(3):
        output = tracker.update(det, rgb)
        assert output.shape == (2, 8)
    assert output[:, 4].tolist() == [1, 2] and output[:, 7].tolist() == [0, 1] and output[:, 6].tolist() == [0, 65]
    assert_allclose(det, np.delete(output, [4, 7], axis=1), atol=1, rtol=7e-3)


def test_bytetrack_contract_edges():
    from yolo_tracking_b200 import BYTETracker
    trk = BYTETracker(max_tracks=64, max_dets=64)
    out = trk.update(np.empty((0, 6)), None)
    assert out.shape == (0,) and out.size == 0                    # byte_tracker.py:280
    with pytest.raises(AssertionError):
        trk.update(np.zeros((2, 5)), None)
    with pytest.raises(AssertionError):
        trk.update([[0, 0, 1, 1, 0.9, 0]], None)
    dets = np.array([[10, 10, 50, 90, 0.9, 1.0]])
    before = dets.copy()
    trk.update(dets, None)
    assert np.array_equal(dets, before)                           # inputs are never mutated
    # conf == track_thresh and conf <= 0.1 fall in neither band (strict inequalities)
    trk2 = BYTETracker(track_thresh=0.5, max_tracks=64, max_dets=64)
    out = trk2.update(np.array([[10, 10, 50, 90, 0.5, 0], [100, 10, 150, 90, 0.1, 0]]), None)
    assert out.size == 0
    with pytest.raises(ValueError):
        trk2.update(np.zeros((65, 6)), None)


def test_kalman_and_matching_mirrors():
    from oracle import boxes, kalman
    from yolo_tracking_b200.motion.kalman_filters import KalmanFilterXYAH, KalmanFilterXYWH, chi2inv95
    from yolo_tracking_b200.utils import iou, matching
    kf = KalmanFilterXYAH()
    z = np.array([320.0, 240.0, 0.5, 100.0])
    m, c = kf.initiate(z)
    om, oc = kalman.initiate("xyah", z)
    assert np.allclose(m, om[0]) and np.allclose(c, oc[0])
    m, c = kf.predict(m, c)
    m, c = kf.update(m, c, z + 1.0)
    d = kf.gating_distance(m, c, np.array([z, z + 50.0]))
    assert d.shape == (2,) and d[1] > chi2inv95[4] > d[0]
    with pytest.raises(ValueError):
        kf.gating_distance(m, c, np.array([z]), metric="nosuch")
    mm, cc = KalmanFilterXYWH().multi_predict(np.stack([m, m]), np.stack([c, c]))
    assert mm.shape == (2, 8) and cc.shape == (2, 8, 8)
    a = np.array([[0, 0, 10, 10], [5, 5, 20, 20.0]])
    assert np.array_equal(iou.iou_batch(a, a), boxes.iou(a, a))
    assert np.array_equal(iou.run_asso_func(iou.giou_batch, a, a, 640, 480), boxes.giou(a, a))
    cost = matching.iou_distance(a, a)
    assert np.array_equal(matching.fuse_score(cost, np.array([0.9, 0.8])), 1 - (1 - cost) * np.array([0.9, 0.8])[None])
    assert matching.iou_distance(np.zeros((0, 4)), a).shape == (0, 2)


def test_per_class_tracker_is_one_stream_per_class():
    """Per-class tracking: every class is an independent stream of one device context; results equal
    one oracle tracker per class, ids are unique across classes, det_ind points into the caller's rows."""
    from oracle.bytetrack import ByteTrackOracle
    from yolo_tracking_b200.per_class import PerClassTracker
    from yolo_tracking_b200.synth import make_stream
    C_ = 3
    streams = [make_stream(6, 40 + c, 14, 30) for c in range(C_)]
    trk = PerClassTracker("bytetrack", n_classes=4, max_tracks=64, max_dets=64, track_thresh=0.5, match_thresh=0.8, track_buffer=30,
                          frame_rate=30)
    oracles = [ByteTrackOracle(0.5, 0.8, 30, 30) for _ in range(4)]
    rng = np.random.default_rng(3)
    for f in range(30):
        parts = []
        for c in range(C_):
            d = streams[c][0][f, :streams[c][1][f]].copy()
            d[:, 5] = c
            parts.append(d)
        dets = np.concatenate(parts, axis=0)
        perm = rng.permutation(len(dets))
        dets = dets[perm]
        out = trk.update(dets, None)
        ref_rows = []
        for c in range(4):
            idx = np.nonzero(dets[:, 5] == c)[0]
            r = oracles[c].update(dets[idx], None).reshape(-1, 8)
            if len(r):
                r = r.copy()
                r[:, 4] = (r[:, 4] - 1) * 4 + c + 1
                r[:, 7] = idx[r[:, 7].astype(int)]
                ref_rows.append(r)
        ref = np.concatenate(ref_rows, axis=0) if ref_rows else np.empty((0, 8))
        assert out.shape == ref.shape, f"frame {f}"
        assert np.array_equal(out[:, 4:], ref[:, 4:]), f"frame {f}"
        assert np.allclose(out[:, :4], ref[:, :4], rtol=1e-9, atol=1e-9)
        assert len(set(out[:, 4])) == len(out)
        assert np.array_equal(dets[out[:, 7].astype(int), 5], out[:, 6])
    trk.close()
