"""GPU parity of the StrongSORT drop-in (host logic of strongsort/sort/tracker.py in Python, every numeric step through the
CUDA operator kernels) against the goldens of the live reference: ids, confirmation / deletion, gallery sizes exact; boxes
and Kalman state to 1e-9; smoothed embeddings bit-exact (same float32 numpy operations on the host side)."""
import numpy as np
import pytest

from _util import assert_close, strongsort_scenario

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["strongsort_c4", "strongsort_churn", "strongsort_cam"])
def test_strongsort_replays_reference_golden(name):
    from yolo_tracking_b200 import StrongSORT
    sc, cfg, dets, nd, feats, g = strongsort_scenario(name)
    trk = StrongSORT(None, 0, False, **cfg)
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
    assert np.array_equal(s["feature"], g["final_feat"])


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
