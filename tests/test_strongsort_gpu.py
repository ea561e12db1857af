"""GPU parity of StrongSORT - the batched frame step (csrc/strongsort_step.cu; the drop-in is a one-stream context of it) and
the operator-backed fallback (host logic of strongsort/sort/tracker.py in Python, every numeric step through the CUDA
operator kernels) - against the goldens of the live reference: ids, confirmation / deletion, gallery sizes exact; boxes
and Kalman state to 1e-9; smoothed embeddings (unit-norm float32, smoothed on the device) to 1e-6 absolute: the reference's
norms come from a BLAS float32 dot product whose summation order is unspecified, the kernel accumulates them in double."""
import numpy as np
import pytest

from _util import assert_close, strongsort_scenario

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("fused", [True, False], ids=["fused", "operators"])
@pytest.mark.parametrize("name", ["strongsort_c4", "strongsort_churn", "strongsort_cam"])
def test_strongsort_replays_reference_golden(name, fused):
    from yolo_tracking_b200 import StrongSORT
    sc, cfg, dets, nd, feats, g = strongsort_scenario(name)
    trk = StrongSORT(None, 0, False, fused=fused, **cfg)
    img = np.zeros((4, 4, 3), dtype=np.uint8)
    cov_frames = {int(f): k for k, f in enumerate(g["cov_frames"])}
    cov_offs = [0]
    for f in g["cov_frames"]:
        cov_offs.append(cov_offs[-1] + int(g["rec_offs"][f + 1] - g["rec_offs"][f]))
    for f in range(sc["n_frames"]):
        out = trk.update(dets[f, :nd[f]], img, feats=feats[f, :nd[f]], warp=None if sc["warps"] is None else sc["warps"][f])
        ref = g["out"][g["out_offs"][f]:g["out_offs"][f + 1]]
        assert out.reshape(-1, 8).shape == ref.shape, f"{name} frame {f}"
        if ref.size:
            assert np.array_equal(out[:, 4:], ref[:, 4:]), f"{name} frame {f}: id/conf/cls/det_ind"
            assert_close(out[:, :4], ref[:, :4], what=f"{name} frame {f} boxes")
        s = trk.state()
        lo, hi = g["rec_offs"][f], g["rec_offs"][f + 1]
        mine = np.stack([s["track_id"], s["state"], s["hits"], s["age"], s["time_since_update"], s["gallery"]], axis=1).reshape(-1, 6)
        assert np.array_equal(mine, g["rec"][lo:hi]), f"{name} frame {f}: track records"
        assert_close(s["mean"], g["mean"][lo:hi], what=f"{name} frame {f} mean")
        if f in cov_frames:
            k = cov_frames[f]
            assert_close(s["cov"].reshape(-1, 64), g["cov"][cov_offs[k]:cov_offs[k + 1]], abs_=1e-10, what=f"{name} frame {f} cov")
    assert (trk._fused is not None) == fused
    assert s["feature"].shape == g["final_feat"].shape and s["feature"].dtype == np.float32
    assert np.abs(s["feature"] - g["final_feat"]).max() < 1e-6, f"{name}: smoothed embeddings"


def test_strongsort_factory_and_seam():
    from oracle.strongsort import StrongSORTOracle
    from yolo_tracking_b200 import create_tracker, get_tracker_config
    from yolo_tracking_b200.synth import make_stream
    import sys
    from _util import GOLDEN
    if GOLDEN not in sys.path:
        sys.path.insert(0, GOLDEN)
    from scenarios import STRONGSORT_YAML
    dets, nd, embs = make_stream(4, 88, 12, 30, emb_dim=64)

    class Seam:
        def get_features(self, xyxys, img):
            f = np.asarray(Seam.next, dtype=np.float32)
            assert len(f) == len(xyxys)
            return f / np.linalg.norm(f)
    trk = create_tracker("strongsort", get_tracker_config("strongsort"), None, 0, False, False, model=Seam())
    orc = StrongSORTOracle(**STRONGSORT_YAML)
    img = np.zeros((8, 8, 3), dtype=np.uint8)
    assert trk.update(np.empty((0, 6)), img).size == 0
    orc.update(np.empty((0, 6)), np.zeros((0, 64), dtype=np.float32))
    for f in range(30):
        d = dets[f, :nd[f]]
        Seam.next = embs[f, :nd[f]]
        out = trk.update(d, img)
        ref = orc.update(d, embs[f, :nd[f]] / np.linalg.norm(embs[f, :nd[f]]))
        assert out.shape == ref.shape
        if ref.size:
            assert np.array_equal(out[:, 4:], ref[:, 4:])
            assert_close(out[:, :4], ref[:, :4])
    with pytest.raises(AssertionError):
        trk.update(np.zeros((2, 5)), img)


def test_new_strongsort_operators_match_numpy():
    """b200track_ema_unit_features / _unit_features / _camera_update_xyah against the reference's numpy arithmetic
    (strongsort/sort/track.py:129-138, :166-172)."""
    from yolo_tracking_b200 import _ops
    rng = np.random.default_rng(3)
    trk = rng.normal(0, 1, (37, 96)).astype(np.float32)
    trk /= np.linalg.norm(trk, axis=1, keepdims=True)
    det = rng.normal(0, 1, (37, 96)).astype(np.float32)
    for alpha in (0.9, 0.8):
        f = det / np.linalg.norm(det, axis=1, keepdims=True)
        ref = alpha * trk + (1 - alpha) * f
        ref /= np.linalg.norm(ref, axis=1, keepdims=True)
        assert np.abs(_ops.ema_unit_features(trk, det, alpha) - ref).max() < 2e-7
    assert np.abs(_ops.unit_features(det) - det / np.linalg.norm(det, axis=1, keepdims=True)).max() < 2e-7
    mean = np.concatenate([rng.uniform(100, 1800, (50, 2)), rng.uniform(0.3, 0.8, (50, 1)), rng.uniform(60, 220, (50, 1)),
                           rng.normal(0, 2, (50, 4))], axis=1)
    warp = np.array([[1.002, -0.004, 3.5], [0.004, 0.998, -2.25]])
    for w in (None, warp):
        ref = mean.copy()
        for m in ref:
            tl = m[:4].copy(); tl[2] *= tl[3]; tl[:2] -= tl[2:] / 2
            x1, y1, x2, y2 = tl[0], tl[1], tl[0] + tl[2], tl[1] + tl[3]
            if w is not None:
                wm = np.array([w[0], w[1], [0, 0, 1]])
                x1, y1, _ = wm @ np.array([x1, y1, 1.0])
                x2, y2, _ = wm @ np.array([x2, y2, 1.0])
            ww, hh = x2 - x1, y2 - y1
            m[:4] = [x1 + ww / 2, y1 + hh / 2, ww / hh, hh]
        assert_close(_ops.camera_update_xyah(mean, w), ref, what="camera_update")


def test_strongsort_runs_the_tensor_core_gallery_distance():
    """The appearance cost of both forms goes through b200track_gallery_cost (device-resident gallery, tcgen05 pre-filter)."""
    import yolo_tracking_b200 as pkg
    sc, cfg, dets, nd, feats, g = strongsort_scenario("strongsort_c4")
    img = np.zeros((1080, 1920, 3), dtype=np.uint8)
    for fused in (True, False):
        trk = pkg.StrongSORT(None, 0, False, fused=fused, **cfg)
        for f in range(12):
            trk.update(dets[f, :nd[f]], img, feats=feats[f, :nd[f]])
        assert (trk._fused is not None) if fused else (trk._store is not None and not trk.samples)
        st = trk.state()
        assert st["gallery"].max() >= 10


def test_batched_strongsort_streams_are_independent_and_match_the_oracle():
    """BatchedTracker("strongsort", S): S different streams in one context, every stream against its own oracle; ragged
    detection counts, an empty stream, streams that start late."""
    from oracle.strongsort import StrongSORTOracle
    from yolo_tracking_b200.batch import BatchedTracker
    from yolo_tracking_b200.synth import make_stream
    S, F, D, T, frames = 5, 64, 64, 96, 45
    cfg = dict(max_dist=0.2, max_iou_dist=0.7, max_age=6, n_init=2, nn_budget=5, mc_lambda=0.995, ema_alpha=0.8)
    streams = [make_stream(20 + s, 7 * s, 6 + 7 * s, frames, emb_dim=F, occlusion=True) for s in range(S)]
    trk = BatchedTracker("strongsort", S, max_tracks=T, max_dets=D, feat_dim=F, **cfg)
    orcs = [StrongSORTOracle(**cfg) for _ in range(S)]
    dets = np.zeros((S, D, 6))
    feats = np.zeros((S, D, F), dtype=np.float32)
    for f in range(frames):
        nd = np.zeros(S, dtype=np.int32)
        for s in range(S):
            d, n, e = streams[s]
            k = 0 if (s == 1 or f < 3 * s) else int(n[f])          # stream 1 never sees a detection; others start late
            nd[s] = k
            dets[s, :k] = d[f, :k]
            if k:
                feats[s, :k] = (e[f, :k] / np.linalg.norm(e[f, :k])).astype(np.float32)
        out, nout = trk.update_batch(np.ascontiguousarray(dets), nd, feats=np.ascontiguousarray(feats))
        for s in range(S):
            ref = orcs[s].update(dets[s, :nd[s]].copy(), feats[s, :nd[s]].copy()).reshape(-1, 8)
            got = out[s, :nout[s]]
            assert got.shape == ref.shape, (f, s)
            if ref.size:
                assert np.array_equal(got[:, 4:], ref[:, 4:]), (f, s)
                assert_close(got[:, :4], ref[:, :4], what=f"frame {f} stream {s}")
            st, snap = trk.state(s), orcs[s].snapshot()
            for key in ("track_id", "state", "hits", "age", "time_since_update", "gallery"):
                assert np.array_equal(st[key], snap[key]), (f, s, key)
            assert_close(st["mean"], snap["mean"], what=f"frame {f} stream {s} mean")
    trk.sync()
    assert trk.launches() == 8 * frames
    trk.close()


@pytest.mark.skipif("__import__('torch').cuda.device_count() < 2")
def test_operator_backed_tracker_on_second_device():
    """StrongSORT(device=1): operator kernels and buffers follow the tracker's device."""
    import torch
    import yolo_tracking_b200 as pkg
    sc, cfg, dets, nd, feats, g = strongsort_scenario("strongsort_churn")
    a, b = pkg.StrongSORT(None, 0, False, **cfg), pkg.StrongSORT(None, 1, False, **cfg)
    img = np.zeros((1080, 1920, 3), dtype=np.uint8)
    for f in range(40):
        ra = a.update(dets[f, :nd[f]], img, feats=feats[f, :nd[f]])
        rb = b.update(dets[f, :nd[f]], img, feats=feats[f, :nd[f]])
        assert np.array_equal(ra, rb), f
    assert torch.cuda.current_device() == 0
