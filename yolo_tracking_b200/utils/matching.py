"""boxmot/utils/matching.py on the GPU (box arrays in, numpy out)."""
from __future__ import annotations

import numpy as np

from .. import _ops


def linear_assignment(cost_matrix, thresh):
    """matching.py:56-71: lap.lapjv(cost, extend_cost=True, cost_limit=thresh)."""
    cost_matrix = np.asarray(cost_matrix)
    if cost_matrix.size == 0:
        return (np.empty((0, 2), dtype=int), tuple(range(cost_matrix.shape[0])), tuple(range(cost_matrix.shape[1])))
    x, y = _ops.lapjv(cost_matrix, thresh)
    rows = np.nonzero(x >= 0)[0]
    matches = np.stack([rows, x[rows]], axis=1).astype(int).reshape(-1, 2)
    return matches, np.where(x < 0)[0], np.where(y < 0)[0]


def iou_distance(atlbrs, btlbrs):
    """matching.py:94-119 for box arrays: 1 - iou."""
    a = np.asarray(atlbrs, dtype=np.float64).reshape(-1, 4)
    b = np.asarray(btlbrs, dtype=np.float64).reshape(-1, 4)
    if len(a) == 0 or len(b) == 0:
        return np.zeros((len(a), len(b)), dtype=np.float32)
    return _ops.iou_distance(a, b)


def fuse_score(cost_matrix, det_scores):
    """matching.py:213-221 with the detection scores passed as an array."""
    cost_matrix = np.asarray(cost_matrix)
    if cost_matrix.size == 0:
        return cost_matrix
    return 1 - (1 - cost_matrix) * np.asarray(det_scores, dtype=np.float64)[None, :]


def embedding_distance(track_features, det_features, metric="cosine"):
    """matching.py:145-167: features are cast to fp32, cosine distance in double, clamped at 0."""
    if metric != "cosine":
        raise ValueError("only the cosine metric is built")
    a = np.asarray(track_features, dtype=np.float32)
    b = np.asarray(det_features, dtype=np.float32)
    if len(a) == 0 or len(b) == 0:
        return np.zeros((len(a), len(b)), dtype=np.float32)
    return _ops.embedding_distance(a, b)
