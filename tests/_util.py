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


def heavy_offsets(g):
    offs, pos = {}, 0
    for f in g["heavy_frames"]:
        n = int(g["rec_offs"][f + 1] - g["rec_offs"][f])
        offs[int(f)] = (pos, pos + n)
        pos += n
    return offs
