"""boxmot/utils/association.py pieces on the GPU."""
from __future__ import annotations

import numpy as np

from .. import _ops


def linear_assignment(cost_matrix):
    """association.py:20-24: lap.lapjv(cost, extend_cost=True) -> array of [row, col]."""
    cost_matrix = np.asarray(cost_matrix)
    if cost_matrix.size == 0:
        return np.empty((0, 2), dtype=int)
    x, _ = _ops.lapjv(cost_matrix)
    rows = np.nonzero(x >= 0)[0]
    return np.stack([rows, x[rows]], axis=1).astype(int).reshape(-1, 2)


def compute_aw_max_metric(emb_cost, w_association_emb, bottom=0.5):
    """association.py:79-108 (DeepOCSORT's adaptive appearance weight): emb_cost [R, C] -> w_emb * emb_cost."""
    return _ops.aw_max_metric(emb_cost, w_association_emb, bottom)
