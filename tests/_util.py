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
