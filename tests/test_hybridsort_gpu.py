"""GPU parity of the fused HybridSORT frame step (csrc/hybridsort_step.cu) through the C-ABI:
(1) the three goldens of the live reference replayed through the reference-shaped drop-in - ids, counters, observed flags
and the reference's odd last column exact, 9-d filter state / boxes / corner velocities to 1e-9 (fp64 tolerance of
BASELINE.json's north_star), float32 smoothed embeddings to 2e-6; (2) several ragged streams in one context against one
oracle per stream; (3) the edge cases the reference handles (empty frames, nothing above det_thresh, trackers ageing out,
two classes through the PerClassDecorator semantics); (4) the factory."""
import numpy as np
import pytest

from _util import assert_close, check_hybridsort_frame, heavy_offsets, hybridsort_scenario

pytestmark = pytest.mark.gpu

IMG = (1080, 1920)


@pytest.mark.parametrize("name", ["hybridsort_c4", "hybridsort_churn", "hybridsort_diou", "hybridsort_2cls"])
def test_hybridsort_dropin_replays_reference(name):
    from yolo_tracking_b200.trackers.hybridsort import HybridSORT
    sc, cfg, dets, nd, feats, g = hybridsort_scenario(name, full=True)
    trk = HybridSORT(None, 0, False, max_tracks=128, max_dets=64, **cfg)
    heavy = heavy_offsets(g)
    for f in range(sc["n_frames"]):
        out = trk.update(dets[f, :nd[f]], IMG, feats=feats[f])
        check_hybridsort_frame(name, f, out, trk.state(), g, heavy)
    assert_close(trk.state()["smooth_feat"], g["final_emb"], rel=2e-6, abs_=2e-6, what="embeddings")
    st = trk.stats
    assert st["oru"] > 10 and st["lap_frames"] > 50 and (name == "hybridsort_2cls" or st["corrections"] > 20)
    assert name != "hybridsort_2cls" or trk.frame_count > 1.8 * sc["n_frames"]      # one update per class and frame


@pytest.mark.parametrize("E", [32, 100, 640])
def test_hybridsort_multi_stream_matches_oracle(E):
    """Five streams of different sizes and hyper-parameter-identical trackers in ONE context, every frame against one oracle
    per stream: rows exact in ids / conf / cls / last column, boxes to 1e-9.  Embedding widths: a multiple of 32 floats, one
    that leaves lanes of the register-tiled cosine pass without data (zero-padded tiles), and one above 512 (generic pass)."""
    from oracle.hybridsort import HybridSortOracle          # checker only
    from yolo_tracking_b200.batch import BatchedTracker
    from yolo_tracking_b200.synth import make_stream
    S, F, D, T = 5, 60 if E == 32 else 25, 64, 128
    cfg = dict(det_thresh=0.25, max_age=12, min_hits=2, iou_threshold=0.3, delta_t=3, asso_func="giou", inertia=0.2)
    streams = [make_stream(4, 700 + s, 8 + 9 * s, F, dmax=D, emb_dim=E, miss_prob=0.15, fp_rate=2.0, occlusion=True) for s in range(S)]
    trk = BatchedTracker("hybridsort", S, max_tracks=T, max_dets=D, feat_dim=E, **cfg)
    orc = [HybridSortOracle(**cfg) for _ in range(S)]
    dets = np.zeros((S, D, 6))
    ndv = np.zeros(S, dtype=np.int32)
    feats = np.zeros((S, D, E), dtype=np.float32)
    for f in range(F):
        for s in range(S):
            d, n, e = streams[s]
            dets[s], ndv[s] = d[f], n[f]
            raw = e[f, :n[f]]
            feats[s] = 0
            if n[f]:
                feats[s, :n[f]] = raw / np.linalg.norm(raw)
        out, nout = trk.update_batch(dets, ndv, feats=feats, img_hw=IMG)
        for s in range(S):
            n = ndv[s]
            keep = dets[s, :n, 4] > cfg["det_thresh"]
            ref = orc[s].update(dets[s, :n], feats[s, :n][keep]).reshape(-1, 8)
            got = out[s, :nout[s]]
            assert got.shape == ref.shape, (f, s, got.shape, ref.shape)
            assert np.array_equal(got[:, 4:], ref[:, 4:]), (f, s)
            assert_close(got[:, :4], ref[:, :4], what=f"frame {f} stream {s} boxes")
    trk.sync()
    assert trk.launches() == F
    assert trk.track_updates() == sum(o.track_updates for o in orc)
    for s in range(S):
        st, ref = trk.state(s), orc[s].snapshot()
        assert np.array_equal(st["track_id"], ref["track_id"]) and np.array_equal(st["hit_streak"], ref["hit_streak"])
        assert_close(st["x"], ref["x"], what=f"stream {s} x")
        assert_close(st["P"], ref["P"], abs_=1e-10, what=f"stream {s} P")
        assert_close(st["smooth_feat"], ref["smooth_feat"], rel=2e-6, abs_=2e-6, what=f"stream {s} embeddings")
    trk.close()


def test_hybridsort_edge_cases():
    """Empty frames before and between detections, a frame with nothing above det_thresh, trackers ageing out, two classes
    (every class call is a frame for all trackers, boxmot/utils/__init__.py:22-61) - against the oracle driven the same way."""
    from oracle.hybridsort import HybridSortOracle          # checker only
    from yolo_tracking_b200.trackers.hybridsort import HybridSORT
    cfg = dict(det_thresh=0.3, max_age=3, min_hits=1, iou_threshold=0.3, delta_t=3, asso_func="iou", inertia=0.2)
    rng = np.random.default_rng(3)
    trk = HybridSORT(None, 0, False, max_tracks=64, max_dets=32, **cfg)
    orc = HybridSortOracle(**cfg)
    E = 16
    protos = rng.standard_normal((6, E)).astype(np.float32)

    def frame(k, conf_scale=1.0, classes=(0,) * 6, present=range(6)):
        rows, fe = [], []
        for i in present:
            x, y = 100 + 150 * i + 3 * k, 200 + 2 * k * (i % 3)
            rows.append([x, y, x + 60, y + 120, min(0.95, (0.5 + 0.07 * i) * conf_scale), classes[i]])
            fe.append(protos[i] + 0.1 * rng.standard_normal(E).astype(np.float32))
        return np.array(rows, dtype=np.float64).reshape(-1, 6), np.array(fe, dtype=np.float32).reshape(-1, E)

    def both(dets, fe):
        got = trk.update(dets, IMG, feats=fe)
        # the oracle restates the undecorated update: drive it like the PerClassDecorator does
        if dets.size:
            classes = set(d[5] for d in dets)
            active = set(np.float64(t.cls) for t in orc.trackers)
            ref = np.empty((0, 8))
            for c in active.union(classes):
                idx = np.array([i for i, d in enumerate(dets) if d[5] == c], dtype=np.int64)
                keep = dets[idx, 4] > cfg["det_thresh"] if len(idx) else np.zeros(0, dtype=bool)
                o = orc.update(dets[idx].reshape(-1, 6), fe[idx][keep])
                if o.size:
                    ref = np.append(ref, o.reshape(-1, 8), axis=0)
        else:
            ref = orc.update(dets, np.zeros((0, E), dtype=np.float32))
        assert got.shape == ref.shape, (got.shape, ref.shape)
        if ref.size:
            assert np.array_equal(got[:, 4:], ref[:, 4:])
            assert_close(got[:, :4], ref[:, :4], what="boxes")

    empty = np.zeros((0, 6))
    for _ in range(2):
        both(empty, np.zeros((0, E), dtype=np.float32))               # before anything was seen
    for k in range(5):
        both(*frame(k))
    both(*frame(5, conf_scale=0.3))                                   # nothing above det_thresh
    both(empty, np.zeros((0, E), dtype=np.float32))
    for k in range(7, 10):
        both(*frame(k, present=range(3)))                             # three objects gone: they age out (max_age 3)
    for k in range(10, 16):
        both(*frame(k, classes=(0, 0, 0, 1, 1, 1)))                   # two classes: two calls per frame
    st, ref = trk.state(), orc.snapshot()
    assert np.array_equal(st["track_id"], ref["track_id"]) and np.array_equal(st["age"], ref["age"])
    assert_close(st["x"], ref["x"], what="x")
    assert trk.frame_count == orc.frame_count


def test_hybridsort_factory_and_rejections():
    import yolo_tracking_b200 as y
    from yolo_tracking_b200 import _lib
    from yolo_tracking_b200.batch import BatchedTracker
    t = y.create_tracker("hybridsort", y.get_tracker_config("hybridsort"), None, 0, False, False, max_tracks=64, max_dets=32)
    assert type(t).__name__ == "HybridSORT" and t.det_thresh == 0 and t.asso_func == "giou" and t.per_class is True
    out = t.update(np.array([[10., 10, 60, 110, 0.9, 0]]), IMG, feats=np.ones((1, 8), dtype=np.float32))
    assert out.shape == (1, 8) and out[0, 4] == 1 and out[0, 7] == 0.9          # last column: the score (hybridsort.py:396)
    with pytest.raises(NotImplementedError):
        y.HybridSORT(use_byte=True)
    with pytest.raises(_lib.B200TrackError):
        BatchedTracker("hybridsort", 1, max_tracks=64, max_dets=32, feat_dim=8, det_thresh=0.0, use_byte=True)
    with pytest.raises(_lib.B200TrackError):
        BatchedTracker("hybridsort", 1, max_tracks=64, max_dets=32, feat_dim=0, det_thresh=0.0)


def test_hybridsort_capacity_overflow_is_reported_not_silent():
    """More detections than max_dets / more live trackers than max_tracks: the step that overflowed reports it when its
    results are collected, b200track_sync repeats it (a context must be reset afterwards)."""
    from yolo_tracking_b200 import _lib
    from yolo_tracking_b200.batch import BatchedTracker
    cap, E = 32, 8
    rng = np.random.default_rng(0)

    def boxes(n):
        c = np.stack([rng.uniform(50, 1800, n), rng.uniform(50, 1000, n)], axis=1)
        return np.concatenate([c - 15, c + 15, np.full((n, 1), 0.9), np.zeros((n, 1))], axis=1)
    trk = BatchedTracker("hybridsort", 1, max_tracks=cap, max_dets=cap, feat_dim=E, det_thresh=0.1, max_age=30, min_hits=1)
    d = np.zeros((1, cap, 6))
    f = rng.standard_normal((1, cap, E)).astype(np.float32)
    d[0] = boxes(cap)
    with pytest.raises(_lib.B200TrackError) as e:
        trk.update_batch(d, np.array([cap + 5], dtype=np.int32), feats=f)      # claims more detections than the buffer holds
    assert e.value.code == _lib.ERR_CAPACITY and "detections" in str(e.value)
    with pytest.raises(_lib.B200TrackError):
        trk.sync()
    trk.reset()
    with pytest.raises(_lib.B200TrackError) as e:
        for _ in range(3):                                            # 3 x 32 well separated boxes with unrelated embeddings:
            d[0] = boxes(cap)                                         # every forced match is corrected away -> > 32 live trackers
            f = rng.standard_normal((1, cap, E)).astype(np.float32)
            trk.update_batch(d, np.array([cap], dtype=np.int32), feats=f)
    assert e.value.code == _lib.ERR_CAPACITY and "tracks" in str(e.value)
    trk.close()


@pytest.mark.skipif("__import__('torch').cuda.device_count() < 2")
def test_hybridsort_on_second_device():
    """HybridSORT(device=1): the context, its state probes and the per-class wrapper's class read follow the tracker's
    device and leave the caller's current device alone."""
    import torch
    from yolo_tracking_b200.trackers.hybridsort import HybridSORT
    sc, cfg, dets, nd, feats, g = hybridsort_scenario("hybridsort_2cls", full=True)
    a, b = HybridSORT(None, 0, False, max_tracks=128, max_dets=64, **cfg), HybridSORT(None, 1, False, max_tracks=128, max_dets=64, **cfg)
    for f in range(40):
        ra = a.update(dets[f, :nd[f]], IMG, feats=feats[f])
        rb = b.update(dets[f, :nd[f]], IMG, feats=feats[f])
        assert np.array_equal(ra, rb), f
    assert np.array_equal(a.state()["x"], b.state()["x"])
    assert torch.cuda.current_device() == 0


@pytest.mark.parametrize("cap", [224, 512])
def test_hybridsort_kernel_variants(cap):
    """The (224, 224) and (512, 512) instantiations of the step kernel (the other tests run the 64 / 128 / 256 ones): two
    streams against the oracle."""
    from oracle.hybridsort import HybridSortOracle          # checker only
    from yolo_tracking_b200.batch import BatchedTracker
    from yolo_tracking_b200.synth import make_stream
    S, F, E = 2, 20, 64
    cfg = dict(det_thresh=0.2, max_age=10, min_hits=1, iou_threshold=0.3, delta_t=3, asso_func="giou", inertia=0.2)
    streams = [make_stream(4, 800 + s, 40 + 30 * s, F, dmax=cap, emb_dim=E, miss_prob=0.1, fp_rate=2.0, occlusion=True) for s in range(S)]
    trk = BatchedTracker("hybridsort", S, max_tracks=cap, max_dets=cap, feat_dim=E, **cfg)
    orc = [HybridSortOracle(**cfg) for _ in range(S)]
    dets = np.zeros((S, cap, 6))
    ndv = np.zeros(S, dtype=np.int32)
    feats = np.zeros((S, cap, E), dtype=np.float32)
    for f in range(F):
        for s in range(S):
            d, n, e = streams[s]
            dets[s], ndv[s] = d[f], n[f]
            feats[s] = 0
            if n[f]:
                feats[s, :n[f]] = e[f, :n[f]] / np.linalg.norm(e[f, :n[f]])
        out, nout = trk.update_batch(dets, ndv, feats=feats, img_hw=IMG)
        for s in range(S):
            n = ndv[s]
            ref = orc[s].update(dets[s, :n], feats[s, :n][dets[s, :n, 4] > cfg["det_thresh"]]).reshape(-1, 8)
            got = out[s, :nout[s]]
            assert got.shape == ref.shape and np.array_equal(got[:, 4:], ref[:, 4:]), (f, s)
            assert_close(got[:, :4], ref[:, :4], what=f"frame {f} stream {s} boxes")
    trk.sync()
    trk.close()
