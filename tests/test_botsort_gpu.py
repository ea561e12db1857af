"""GPU parity: the CUDA BoT-SORT step (through the C-ABI) against the goldens of the live reference and
against the oracle on multi-stream scenes.  Integer / lifecycle fields exact, boxes and Kalman state to
1e-9 relative; smoothed embeddings to 2e-6 absolute (they are unit-norm float32 vectors: the reference's
norms come from a BLAS sdot whose summation order is unspecified, ours accumulate in double)."""
import numpy as np
import pytest

from _util import assert_close, botsort_scenario

pytestmark = pytest.mark.gpu
FEAT_TOL = 2e-6


def _cfg_kwargs(cfg):
    keys = ("track_high_thresh", "track_low_thresh", "new_track_thresh", "track_buffer", "match_thresh",
            "proximity_thresh", "appearance_thresh", "frame_rate", "with_reid", "fuse_first_associate")
    return {k: cfg[k] for k in keys if k in cfg}


def _check_state(st, snap, what, with_reid):
    for a, b in (("track_id", "track_id"), ("state", "state"), ("is_activated", "is_activated"), ("frame_id_t", "frame_id"),
                 ("start_frame", "start_frame"), ("tracklet_len", "tracklet_len")):
        assert np.array_equal(st[a], snap[b]), f"{what}: {a}\n{st[a]}\n{snap[b]}"
    assert (st["n_tracked"], st["n_lost"]) == (int(snap["n_tracked"]), int(snap["n_lost"])), what
    assert np.array_equal(st["score"], snap["score"]) and np.array_equal(st["cls"], snap["cls"]), what + " score/cls"
    assert np.array_equal(st["det_ind"], snap["det_ind"]), what + " det_ind"
    assert_close(st["mean"], snap["mean"], what=what + " mean")
    assert_close(st["cov"], snap["cov"], abs_=1e-10, what=what + " cov")
    if with_reid and len(snap["track_id"]):
        assert np.abs(st["smooth_feat"] - snap["smooth_feat"]).max() < FEAT_TOL, what + " smooth_feat"


@pytest.mark.parametrize("name,cam", [("botsort_c3", False), ("botsort_churn", False), ("botsort_noreid", False), ("botsort_fuse", False),
                                      ("botsort_cam", True),          # a scripted moving camera: warps applied inside the fused step
                                      ("botsort_churn", True)])       # the camera-motion form of the filter without warps = identity
def test_botsort_replays_reference_golden(name, cam):
    from yolo_tracking_b200.batch import BatchedTracker
    sc, cfg, dets, nd, feats, g = botsort_scenario(name)
    with_reid = cfg.get("with_reid", True)
    D = dets.shape[1]
    cap = 64 if D <= 64 else 128
    F = sc["emb_dim"] if with_reid else 0
    Fpad = (F + 127) // 128 * 128
    trk = BatchedTracker("botsort", 1, max_tracks=cap, max_dets=cap, feat_dim=Fpad, camera_motion=cam, **_cfg_kwargs(cfg))
    cov_frames = {int(f): k for k, f in enumerate(g["cov_frames"])}
    cov_offs = [0]
    for f in g["cov_frames"]:
        cov_offs.append(cov_offs[-1] + int(g["counts"][f].sum()))
    d = np.zeros((1, cap, 6))
    ft = np.zeros((1, cap, Fpad), dtype=np.float32) if with_reid else None
    for f in range(sc["n_frames"]):
        d[0, :D] = dets[f]
        if with_reid:
            ft[0, :D, :F] = feats[f]
        if cam:     # packed frames carry the warp of every stream (b200track_submit_packed)
            rows = trk.update_frames([dets[f, :nd[f]]], feats=None if ft is None else [ft[0, :nd[f]]],
                                     warps=None if sc["warps"] is None else sc["warps"][f].reshape(1, 6))[0]
            out, nout = rows[None], np.array([len(rows)])
        else:
            out, nout = trk.update_batch(d, np.array([nd[f]], dtype=np.int32), feats=ft)
        ref = g["out"][g["out_offs"][f]:g["out_offs"][f + 1]]
        assert nout[0] == len(ref), f"{name} frame {f}: {nout[0]} rows vs {len(ref)}"
        o = out[0, :nout[0]]
        assert np.array_equal(o[:, 4:], ref[:, 4:]), f"{name} frame {f}: id/conf/cls/det_ind\n{o[:, 4:]}\n{ref[:, 4:]}"
        assert_close(o[:, :4], ref[:, :4], what=f"{name} frame {f} boxes")
        st = trk.state(0)
        assert (st["n_tracked"], st["n_lost"]) == tuple(g["counts"][f]), f"{name} frame {f}: list sizes"
        lo, hi = g["rec_offs"][f], g["rec_offs"][f + 1]
        mine = np.stack([st["track_id"], st["state"], st["is_activated"], st["frame_id_t"], st["start_frame"],
                         st["tracklet_len"]], axis=1).reshape(-1, 6)
        assert np.array_equal(mine, g["rec"][lo:hi]), f"{name} frame {f}: lifecycle records"
        assert_close(st["mean"], g["mean"][lo:hi], what=f"{name} frame {f} mean")
        assert np.array_equal(np.stack([st["score"], st["cls"], st["det_ind"]], axis=1).reshape(-1, 3), g["aux"][lo:hi])
        if f in cov_frames:
            k = cov_frames[f]
            assert_close(st["cov"].reshape(-1, 64), g["cov"][cov_offs[k]:cov_offs[k + 1]], abs_=1e-10, what=f"{name} frame {f} cov")
    if with_reid:
        assert np.abs(st["smooth_feat"][:, :F] - g["final_feat"]).max() < FEAT_TOL
    trk.sync()
    trk.close()


@pytest.mark.parametrize("n_streams,n_objects,n_frames,emb,kw,params", [
    (6, 40, 60, 128, {}, {}),
    (3, 100, 30, 512, {}, {}),
    (2, 100, 30, 256, dict(cap=224), {}),
    (4, 16, 120, 128, dict(miss_prob=0.3, fp_rate=3.0), dict(track_high_thresh=0.5, new_track_thresh=0.6, match_thresh=0.8,
                                                             proximity_thresh=0.5, appearance_thresh=0.25, track_buffer=20)),
    (4, 30, 50, 128, dict(miss_prob=0.1), dict(with_reid=False)),
])
def test_botsort_multistream_vs_oracle(n_streams, n_objects, n_frames, emb, kw, params):
    import sys
    from _util import GOLDEN
    if GOLDEN not in sys.path:
        sys.path.insert(0, GOLDEN)
    from scenarios import BOTSORT_YAML
    from oracle.botsort import BoTSORTOracle
    from yolo_tracking_b200.batch import BatchedTracker
    from yolo_tracking_b200.synth import make_batch
    kw = dict(kw)
    cap = kw.pop("cap", 256 if n_objects > 60 else 128)
    cfg = dict(BOTSORT_YAML)
    cfg.update(params)
    with_reid = cfg.get("with_reid", True)
    dets, nd, embs = make_batch(3, n_streams, n_objects, n_frames, dmax=cap, first_stream=70, emb_dim=emb, **kw)
    # the ReID seam: rows of first-round detections divided by the Frobenius norm of their matrix
    feats = np.zeros_like(embs)
    for f in range(n_frames):
        for s in range(n_streams):
            rows = np.nonzero(dets[f, s, :nd[f, s], 4] > cfg["track_high_thresh"])[0]
            if len(rows):
                feats[f, s, rows] = embs[f, s, rows] / np.linalg.norm(embs[f, s, rows])
    trk = BatchedTracker("botsort", n_streams, max_tracks=cap, max_dets=cap, feat_dim=emb if with_reid else 0, **_cfg_kwargs(cfg))
    oracles = [BoTSORTOracle(**cfg) for _ in range(n_streams)]
    for f in range(n_frames):
        out, nout = trk.update_batch(np.ascontiguousarray(dets[f]), np.ascontiguousarray(nd[f]),
                                     feats=np.ascontiguousarray(feats[f]) if with_reid else None)
        for s in range(n_streams):
            ref = oracles[s].update(dets[f, s, :nd[f, s]], feats[f, s, :nd[f, s]]).reshape(-1, 8)
            assert nout[s] == len(ref), f"frame {f} stream {s}: rows {nout[s]} vs {len(ref)}"
            o = out[s, :nout[s]]
            assert np.array_equal(o[:, 4:], ref[:, 4:]), f"frame {f} stream {s}: ids\n{o[:, 4:]}\n{ref[:, 4:]}"
            assert_close(o[:, :4], ref[:, :4], what=f"frame {f} stream {s} boxes")
        if f % 10 == 9 or f == n_frames - 1:
            for s in range(n_streams):
                _check_state(trk.state(s), oracles[s].snapshot(), f"frame {f} stream {s}", with_reid)
    trk.sync()
    assert trk.track_updates() == sum(o.track_updates for o in oracles)
    trk.close()


def test_botsort_reference_shaped_api():
    """create_tracker('botsort', ...) / tracker.update(dets, img) with a ReID seam object, against the oracle."""
    from oracle.botsort import BoTSORTOracle
    from yolo_tracking_b200 import create_tracker, get_tracker_config
    from yolo_tracking_b200.synth import make_stream
    import sys
    from _util import GOLDEN
    if GOLDEN not in sys.path:
        sys.path.insert(0, GOLDEN)
    from scenarios import BOTSORT_YAML
    dets, nd, embs = make_stream(3, 77, 20, 40, emb_dim=128)

    class Seam:
        queue = []

        def get_features(self, xyxys, img):
            f = np.asarray(Seam.queue.pop(0), dtype=np.float32)
            assert len(f) == len(xyxys)
            return f / np.linalg.norm(f)
    trk = create_tracker("botsort", get_tracker_config("botsort"), None, 0, False, False, model=Seam(), feat_dim=128,
                         max_tracks=64, max_dets=64)
    orc = BoTSORTOracle(**BOTSORT_YAML)
    img = np.zeros((1080, 1920, 3), dtype=np.uint8)
    assert trk.update(np.empty((0, 6)), img).shape == (0,)
    orc.update(np.empty((0, 6)))
    for f in range(40):
        d = dets[f, :nd[f]]
        rows = np.nonzero(d[:, 4] > BOTSORT_YAML["track_high_thresh"])[0]
        ft = np.zeros((len(d), 128), dtype=np.float32)
        if len(rows):
            Seam.queue.append(embs[f, rows])
            ft[rows] = embs[f, rows] / np.linalg.norm(embs[f, rows])
        out = trk.update(d, img)
        ref = orc.update(d, ft)
        assert out.shape == ref.shape
        if ref.size:
            assert np.array_equal(out[:, 4:], ref[:, 4:])
            assert_close(out[:, :4], ref[:, :4])
    with pytest.raises(AssertionError):
        trk.update(np.zeros((2, 5)), img)


def test_botsort_camera_warp_needs_the_camera_form():
    """A context created without camera_motion refuses warps (error, not silently wrong tracks); the device interface
    b200track_step_cam applies them like the packed one."""
    import torch
    from yolo_tracking_b200 import _lib
    from yolo_tracking_b200.batch import BatchedTracker
    sc, cfg, dets, nd, feats, g = botsort_scenario("botsort_cam")
    kw = _cfg_kwargs(cfg)
    plain = BatchedTracker("botsort", 1, max_tracks=64, max_dets=64, feat_dim=128, **kw)
    with pytest.raises(_lib.B200TrackError) as e:
        plain.update_frames([dets[0, :nd[0]]], feats=[np.zeros((nd[0], 128), dtype=np.float32)], warps=sc["warps"][0].reshape(1, 6))
    assert e.value.code == _lib.ERR_STATE
    plain.close()
    a = BatchedTracker("botsort", 1, max_tracks=64, max_dets=64, feat_dim=128, camera_motion=True, **kw)
    b = BatchedTracker("botsort", 1, max_tracks=64, max_dets=64, feat_dim=128, camera_motion=True, **kw)
    D, F = dets.shape[1], sc["emb_dim"]
    d_out = torch.zeros((1, 64, 8), dtype=torch.float64, device="cuda")
    d_nout = torch.zeros((1,), dtype=torch.int32, device="cuda")
    for f in range(30):
        ft = np.zeros((nd[f], 128), dtype=np.float32)
        ft[:, :F] = feats[f][:nd[f]]
        ref = a.update_frames([dets[f, :nd[f]]], feats=[ft], warps=sc["warps"][f].reshape(1, 6))[0]
        dd = np.zeros((1, 64, 6)); dd[0, :nd[f]] = dets[f, :nd[f]]
        df = np.zeros((1, 64, 128), dtype=np.float32); df[0, :nd[f]] = ft
        b.step_device(torch.from_numpy(dd).cuda(), torch.tensor([nd[f]], dtype=torch.int32, device="cuda"), d_out, d_nout,
                      d_feats=torch.from_numpy(df).cuda(), d_warps=torch.from_numpy(sc["warps"][f].reshape(1, 6).copy()).cuda())
        torch.cuda.synchronize()
        assert np.array_equal(d_out[0, :int(d_nout[0])].cpu().numpy(), ref), f
    a.close(); b.close()
