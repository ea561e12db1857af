"""GPU parity: the CUDA OC-SORT step (through the C-ABI) against the goldens of the live
reference and against the oracle on multi-stream synthetic scenes with occlusions."""
import numpy as np
import pytest

from _util import assert_close, load_golden

pytestmark = pytest.mark.gpu


def _pad(dets_list, dmax):
    d = np.zeros((len(dets_list), dmax, 6))
    n = np.zeros(len(dets_list), dtype=np.int32)
    for s, a in enumerate(dets_list):
        d[s, :len(a)] = a
        n[s] = len(a)
    return d, n


def _check_state(st, snap, what):
    for k in ("track_id", "age", "time_since_update", "hits", "hit_streak", "observed"):
        assert np.array_equal(st[k], snap[k]), f"{what}: {k}\n{st[k]}\n{snap[k]}"
    assert_close(st["x"], snap["x"], what=what + " x")
    assert_close(st["P"], snap["P"], abs_=1e-9, what=what + " P")
    assert_close(st["velocity"], snap["velocity"], what=what + " velocity")
    assert_close(st["last_observation"], snap["last_observation"], what=what + " last_observation")


@pytest.mark.parametrize("name", ["ocsort_c2", "ocsort_churn", "ocsort_byte"])
def test_ocsort_replays_reference_golden(name):
    from yolo_tracking_b200.batch import BatchedTracker
    g = load_golden(name)
    p = g["params"]
    dets, nd = g["dets"], g["ndets"]
    hw = tuple(int(v) for v in g["img_hw"])
    trk = BatchedTracker("ocsort", 1, max_tracks=128, max_dets=128, det_thresh=p[0], max_age=int(p[1]), min_hits=int(p[2]),
                         asso_threshold=p[3], delta_t=int(p[4]), asso_func="giou", inertia=p[5],
                         use_byte=bool(p[6]) if len(p) > 6 else False)
    heavy = {int(f): k for k, f in enumerate(g["heavy_frames"])}
    p_offs = [0]
    for f in g["heavy_frames"]:
        p_offs.append(p_offs[-1] + int(g["rec_offs"][f + 1] - g["rec_offs"][f]))
    for f in range(dets.shape[0]):
        d, n = _pad([dets[f, :nd[f]]], 128)
        out, nout = trk.update_batch(d, n, img_hw=hw)
        ref = g["out"][g["out_offs"][f]:g["out_offs"][f + 1]]
        assert nout[0] == len(ref), f"{name} frame {f}: {nout[0]} rows vs {len(ref)}"
        o = out[0, :nout[0]]
        assert np.array_equal(o[:, 4:], ref[:, 4:]), f"{name} frame {f}: id/conf/cls/det_ind\n{o[:, 4:]}\n{ref[:, 4:]}"
        assert_close(o[:, :4], ref[:, :4], what=f"{name} frame {f} boxes")
        st = trk.state(0)
        lo, hi = g["rec_offs"][f], g["rec_offs"][f + 1]
        mine = np.stack([st["track_id"], st["age"], st["time_since_update"], st["hits"], st["hit_streak"], st["observed"]],
                        axis=1).reshape(-1, 6)
        assert np.array_equal(mine, g["rec"][lo:hi]), f"{name} frame {f}: lifecycle records"
        assert_close(st["x"], g["x"][lo:hi], what=f"{name} frame {f} x")
        assert_close(st["velocity"], g["vel"][lo:hi], what=f"{name} frame {f} velocity")
        assert_close(st["last_observation"], g["last"][lo:hi], what=f"{name} frame {f} last_observation")
        if f in heavy:
            k = heavy[f]
            assert_close(st["P"].reshape(-1, 49), g["P"][p_offs[k]:p_offs[k + 1]], abs_=1e-9, what=f"{name} frame {f} P")
    trk.sync()
    trk.close()


def test_ocsort_reference_known_answers():
    from yolo_tracking_b200 import create_tracker, get_tracker_config, OCSORT
    g = load_golden("ocsort_2box")
    img = np.zeros((1080, 1920, 3), dtype=np.uint8)
    trk = create_tracker("ocsort", get_tracker_config("ocsort"), None, 0, False, False, max_tracks=64, max_dets=64)
    for k in range(3):
        out = trk.update(g["det"], img)
        assert out.shape == (2, 8)
        assert_close(out, g["out"][k])
    trk = OCSORT(per_class=False, det_thresh=0, max_age=30, min_hits=2, asso_threshold=0.3, delta_t=3, asso_func="giou",
                 inertia=0.2, max_tracks=64, max_dets=64)
    seq = [np.empty((0, 6)), np.empty((0, 6)), g["det"], np.empty((0, 6)), g["det"], g["det"], g["det"]]
    assert [trk.update(d, img).size for d in seq] == g["min_hits_sizes"].tolist()


@pytest.mark.parametrize("n_streams,n_objects,n_frames,kw,params", [
    (8, 40, 80, dict(occlusion=True), {}),
    (4, 16, 120, dict(miss_prob=0.3, fp_rate=3.0), dict(min_hits=3, max_age=6)),
    (2, 100, 40, dict(occlusion=True), {}),
    (2, 100, 40, dict(occlusion=True, cap=224), {}),
    (4, 30, 60, dict(occlusion=True), dict(asso_func="iou", det_thresh=0.3)),
    (2, 30, 50, dict(occlusion=True), dict(asso_func="diou")),
    (4, 40, 80, dict(occlusion=True, miss_prob=0.1), dict(use_byte=True, det_thresh=0.5, min_hits=2)),
    (3, 24, 90, dict(miss_prob=0.25, fp_rate=3.0), dict(use_byte=True, det_thresh=0.6, max_age=10, asso_func="iou")),
])
def test_ocsort_multistream_vs_oracle(n_streams, n_objects, n_frames, kw, params):
    from oracle.ocsort import OCSortOracle
    from yolo_tracking_b200.batch import BatchedTracker
    from yolo_tracking_b200.synth import make_batch
    kw = dict(kw)
    cap = kw.pop("cap", 256 if n_objects > 60 else 128)
    cfg = dict(det_thresh=0, max_age=30, min_hits=1, asso_threshold=0.3, delta_t=3, asso_func="giou", inertia=0.2)
    cfg.update(params)
    dets, nd, _ = make_batch(2, n_streams, n_objects, n_frames, dmax=cap, first_stream=50, **kw)
    trk = BatchedTracker("ocsort", n_streams, max_tracks=cap, max_dets=cap, **cfg)
    oracles = [OCSortOracle(False, **dict(dict(use_byte=False), **cfg)) for _ in range(n_streams)]
    hw = (2160, 3840) if n_objects > 64 else (1080, 1920)
    for f in range(n_frames):
        out, nout = trk.update_batch(np.ascontiguousarray(dets[f]), np.ascontiguousarray(nd[f]), img_hw=hw)
        for s in range(n_streams):
            ref = oracles[s].update(dets[f, s, :nd[f, s]], hw).reshape(-1, 8)
            assert nout[s] == len(ref), f"frame {f} stream {s}: rows {nout[s]} vs {len(ref)}"
            o = out[s, :nout[s]]
            assert np.array_equal(o[:, 4:], ref[:, 4:]), f"frame {f} stream {s}: ids\n{o[:, 4:]}\n{ref[:, 4:]}"
            assert_close(o[:, :4], ref[:, :4], what=f"frame {f} stream {s} boxes")
        if f % 10 == 9 or f == n_frames - 1:
            for s in range(n_streams):
                _check_state(trk.state(s), oracles[s].snapshot(), f"frame {f} stream {s}")
    trk.sync()
    assert trk.track_updates() == sum(o.track_updates for o in oracles)
    assert sum(o.stats["oru"] for o in oracles) > 0
    trk.close()
