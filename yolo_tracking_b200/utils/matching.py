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


def _track_arrays(tracks):
    if isinstance(tracks, tuple):                      # (means [T, 8], covariances [T, 8, 8])
        return np.asarray(tracks[0], dtype=np.float64), np.asarray(tracks[1], dtype=np.float64)
    return (np.asarray([t.mean for t in tracks], dtype=np.float64), np.asarray([t.covariance for t in tracks], dtype=np.float64))


def _det_xyah(detections):
    if isinstance(detections, np.ndarray):
        return np.asarray(detections, dtype=np.float64).reshape(-1, 4)
    return np.asarray([d.to_xyah() for d in detections], dtype=np.float64).reshape(-1, 4)


def gate_cost_matrix(kf, cost_matrix, tracks, detections, only_position=False):
    """matching.py:170-181.  `kf` is one of yolo_tracking_b200.motion.kalman_filters; tracks are objects with
    .mean / .covariance (or a (means, covariances) tuple), detections objects with .to_xyah() (or an [D, 4] array)."""
    cost_matrix = np.asarray(cost_matrix, dtype=np.float64)
    if cost_matrix.size == 0:
        return cost_matrix
    mean, cov = _track_arrays(tracks)
    return _ops.gate_cost(kf._kind, cost_matrix, mean, cov, _det_xyah(detections), only_position, fuse=False)


def fuse_motion(kf, cost_matrix, tracks, detections, only_position=False, lambda_=0.98):
    """matching.py:184-196."""
    cost_matrix = np.asarray(cost_matrix, dtype=np.float64)
    if cost_matrix.size == 0:
        return cost_matrix
    mean, cov = _track_arrays(tracks)
    return _ops.gate_cost(kf._kind, cost_matrix, mean, cov, _det_xyah(detections), only_position, fuse=True, lambda_=lambda_)
