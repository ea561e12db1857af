"""Import the UNMODIFIED reference (/root/reference) in this container so golden vectors can
be generated from it.  This module is only usable where
/root/reference exists; nothing that runs on the GPU box imports it.

Shims (none of them touches the reference sources):
  * ``boxmot`` is registered as a bare namespace module so the eager imports of
    ``boxmot/__init__.py`` (ReID zoo -> gdown/ftfy/yacs, absent here) are skipped.
  * ``lap`` (pip ``lapx``, un-vendored, not installable offline): ``lap.lapjv`` is restated
    from gatagat/lap's ``_lapjv.pyx`` semantics - build the (R+C)x(R+C) extended matrix and
    solve it exactly with scipy's ``linear_sum_assignment``.  On tie-free inputs the optimum
    is unique, so any exact solver returns the same ``x, y``.  Exact ties of the no-limit call
    (OC-SORT's structurally zero costs) are broken canonically towards lower indices by a
    2**-50 * (r * C + c) perturbation - see oracle/lap.py "Ties".
  * ``filterpy.common.reshape_z`` / ``filterpy.stats.logpdf`` (only ``reshape_z`` executes).
  * ReID model -> queued fixed embeddings; GMC -> identity warp.
"""
from __future__ import annotations

import sys
import types

import numpy as np

REF_ROOT = "/root/reference"


def _lapjv(cost, extend_cost=False, cost_limit=np.inf, return_cost=True):
    from scipy.optimize import linear_sum_assignment
    cost = np.ascontiguousarray(cost, dtype=np.float64)
    R, C = cost.shape
    if not cost_limit < np.inf and R and C:
        # canonical tie-break of the no-limit call site (oracle/lap.py "Ties")
        cost = cost + (np.arange(R, dtype=np.float64)[:, None] * C + np.arange(C, dtype=np.float64)[None, :]) * 2.0 ** -50
    if extend_cost or cost_limit < np.inf:
        n = R + C
        ext = np.empty((n, n), dtype=np.float64)
        ext[:] = cost_limit / 2.0 if cost_limit < np.inf else cost.max() + 1
        ext[R:, C:] = 0
        ext[:R, :C] = cost
    else:
        n = R
        ext = cost
    r, c = linear_sum_assignment(ext)
    x = np.full(n, -1, dtype=np.int32)
    y = np.full(n, -1, dtype=np.int32)
    x[r] = c
    y[c] = r
    opt = float(ext[r, c].sum())
    if n != R:
        x[x >= C] = -1
        y[y >= R] = -1
        x = x[:R]
        y = y[:C]
    return (opt, x, y) if return_cost else (x, y)


def _reshape_z(z, dim_z, ndim):
    z = np.atleast_2d(z)
    if z.shape[1] == dim_z:
        z = z.T
    if z.shape != (dim_z, 1):
        raise ValueError("z must be convertible to shape ({}, 1)".format(dim_z))
    if ndim == 1:
        z = z[:, 0]
    if ndim == 0:
        z = z[0, 0]
    return z


class FakeReID:
    """Stands in for ReIDDetectMultiBackend: returns queued embeddings through the same
    whole-matrix Frobenius normalisation as reid_multibackend.py:304-311."""
    queue: list = []

    def __init__(self, weights=None, device=None, fp16=False):
        pass

    def warmup(self, *a, **k):
        pass

    def get_features(self, xyxys, img):
        if len(xyxys) == 0:
            return np.array([])
        f = np.asarray(FakeReID.queue.pop(0), dtype=np.float32)
        assert f.shape[0] == len(xyxys), (f.shape, len(xyxys))
        return f / np.linalg.norm(f)


class IdentityCMC:
    def apply(self, img, dets):
        return np.eye(2, 3)


class ScriptedCMC:
    """A camera-motion estimator that returns externally supplied warps: set .frame before every update()."""
    def __init__(self, warps):
        self.warps, self.frame = np.asarray(warps, dtype=np.float64), 0

    def apply(self, img, dets):
        return self.warps[self.frame].copy()


_installed = False


def install():
    global _installed
    if _installed:
        return
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    pkg = types.ModuleType("boxmot")
    pkg.__path__ = [REF_ROOT + "/boxmot"]
    sys.modules["boxmot"] = pkg
    lap = types.ModuleType("lap")
    lap.lapjv = _lapjv
    sys.modules["lap"] = lap
    fp = types.ModuleType("filterpy")
    fpc = types.ModuleType("filterpy.common")
    fpc.reshape_z = _reshape_z
    fpc.pretty_str = lambda label, x: f"{label} = {x!r}"
    fps = types.ModuleType("filterpy.stats")
    fps.logpdf = lambda *a, **k: 0.0
    fp.common, fp.stats = fpc, fps
    sys.modules.update({"filterpy": fp, "filterpy.common": fpc, "filterpy.stats": fps})
    app = types.ModuleType("boxmot.appearance")
    app.__path__ = [REF_ROOT + "/boxmot/appearance"]
    rm = types.ModuleType("boxmot.appearance.reid_multibackend")
    rm.ReIDDetectMultiBackend = FakeReID
    sys.modules["boxmot.appearance"] = app
    sys.modules["boxmot.appearance.reid_multibackend"] = rm
    _installed = True


def reset_counters():
    """Class-level ID counters are process-global in the reference; the contract here is
    per-stream ids 1,2,3..., so reset before every stream."""
    from boxmot.trackers.bytetrack.basetrack import BaseTrack as B1
    B1._count = 0
    try:
        from boxmot.trackers.botsort.basetrack import BaseTrack as B2
        B2._count = 0
    except Exception:
        pass
    try:
        from boxmot.trackers.ocsort.ocsort import KalmanBoxTracker
        KalmanBoxTracker.count = 0
    except Exception:
        pass


def make_tracker(name: str, **overrides):
    install()
    from boxmot.tracker_zoo import create_tracker, get_tracker_config
    reset_counters()
    trk = create_tracker(name, get_tracker_config(name), None, "cpu", False, False)
    for k, v in overrides.items():
        setattr(trk, k, v)
    if hasattr(trk, "cmc"):
        trk.cmc = IdentityCMC()
    return trk
