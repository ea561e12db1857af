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
