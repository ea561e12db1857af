"""ORACLE (test infrastructure): linear assignment with lap.lapjv's extended-matrix semantics.

The algorithm lives in a third-party dependency that is NOT vendored in the reference:
PyPI ``lapx>=0.5.4`` (import name ``lap``; requirements.txt:8), a C++ Jonker-Volgenant
solver.  It is not installable offline, so its published behaviour is restated here
(gatagat/lap ``_lapjv.pyx``):

    lapjv(cost[R,C], extend_cost=True, cost_limit=L)
      n = R + C;  ext = full((n,n), L/2 if L finite else cost.max()+1)
      ext[R:, C:] = 0;  ext[:R, :C] = cost;  solve the square LAP on ext;
      x[x >= C] = -1;  y[y >= R] = -1;  return x[:R], y[:C]

i.e. the optimum of  sum(c_ij over matched) + L/2 * (#unmatched rows + #unmatched cols).
It is NOT "Hungarian then threshold" (known-answer vector in tests/test_oracle_lap.py).
The extended matrix is built literally and solved exactly with scipy's
``linear_sum_assignment``; on tie-free inputs the optimum is unique so every exact solver
returns the same x, y.  Parity anchor: the reference's call sites
  boxmot/utils/matching.py:56-71     linear_assignment(cost, thresh)  (ByteTrack/BoTSORT)
  boxmot/utils/association.py:20-24  linear_assignment(cost)          (OCSORT family)

Ties.  lapjv's behaviour on EXACTLY tied optima is an implementation detail of the absent
package.  ByteTrack / BoTSORT costs are tie-free on continuous inputs, but the OC-SORT cost
-(similarity + angle) is structurally full of exact zeros (disjoint boxes, trackers without a
velocity yet), and which of several junk detections stays unassigned decides the creation
order - hence the ids - of new trackers.  To make that well defined, the no-limit call site
(`assign_no_limit`) breaks ties canonically towards lower indices: cost[r, c] += 2**-50 *
(r * C + c) before the solve (at most ~1e-8 in total, far below any real cost gap).  The same
rule is used by the lap shim that generates the goldens (tests/golden/ref_harness.py) and by
the CUDA path (csrc/ocsort_step.cu).
"""
from __future__ import annotations

import numpy as np
from scipy.optimize import linear_sum_assignment

TIE_EPS = 2.0 ** -50


def tie_break(cost):
    """cost[r, c] + TIE_EPS * (r * C + c), evaluated exactly like the CUDA path."""
    cost = np.asarray(cost, dtype=np.float64)
    R, C = cost.shape
    idx = (np.arange(R, dtype=np.float64)[:, None] * C + np.arange(C, dtype=np.float64)[None, :])
    return cost + idx * TIE_EPS


def lapjv_extended(cost, cost_limit=np.inf):
    cost = np.ascontiguousarray(cost, dtype=np.float64)
    R, C = cost.shape
    n = R + C
    fill = cost_limit / 2.0 if cost_limit < np.inf else cost.max() + 1
    ext = np.full((n, n), fill, dtype=np.float64)
    ext[R:, C:] = 0
    ext[:R, :C] = cost
    rows, cols = linear_sum_assignment(ext)
    x = np.full(n, -1, dtype=np.int32)
    y = np.full(n, -1, dtype=np.int32)
    x[rows] = cols
    y[cols] = rows
    x[x >= C] = -1
    y[y >= R] = -1
    return float(ext[rows, cols].sum()), x[:R], y[:C]


def assign_with_limit(cost, thresh):
    """matching.py:56-71 -> (matches[k,2] rows ascending, unmatched_rows, unmatched_cols)."""
    cost = np.asarray(cost)
    if cost.size == 0:
        return (np.empty((0, 2), dtype=int), np.arange(cost.shape[0]), np.arange(cost.shape[1]))
    _, x, y = lapjv_extended(cost, thresh)
    rows = np.nonzero(x >= 0)[0]
    matches = np.stack([rows, x[rows]], axis=1).astype(int).reshape(-1, 2)
    return matches, np.nonzero(x < 0)[0], np.nonzero(y < 0)[0]


def assign_no_limit(cost):
    """association.py:20-24 -> array of [row, col], rows ascending (every min(R,C) matched)."""
    cost = np.asarray(cost)
    if cost.size == 0:
        return np.empty((0, 2), dtype=int)
    _, x, _ = lapjv_extended(tie_break(cost))
    rows = np.nonzero(x >= 0)[0]
    return np.stack([rows, x[rows]], axis=1).astype(int).reshape(-1, 2)
