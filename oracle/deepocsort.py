"""Test infrastructure (CPU oracle): DeepOCSORT's adaptive appearance weight, restated from the reference
(boxmot/utils/association.py:79-108).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may import this.
Pinned by tests/golden/aux_ops.npz (outputs of the live reference, tests/golden/make_golden.py::gen_aux)."""
import numpy as np


def _weights(m, axis, bottom):
    """per row (axis=1) / column (axis=0): 1 - max(second / first - bottom, 0) / (1 - bottom), 0 if first == 0"""
    n = m.shape[axis]
    other = m.shape[1 - axis]
    if n < 2:
        return np.ones(other)
    part = -np.sort(-m, axis=axis)
    first = np.take(part, 0, axis=axis)
    second = np.take(part, 1, axis=axis)
    with np.errstate(divide="ignore", invalid="ignore"):
        w = 1 - np.maximum(second / first - bottom, 0) / (1 - bottom)
    return np.where(first == 0, 0.0, w)


def compute_aw_max_metric(emb_cost, w_association_emb, bottom=0.5):
    """association.py:79-108: w_emb = w_assoc * row weight * column weight; returns w_emb * emb_cost."""
    m = np.asarray(emb_cost, dtype=np.float64)
    w = np.full_like(m, w_association_emb)
    w *= _weights(m, 1, bottom)[:, None]
    w *= _weights(m, 0, bottom)[None, :]
    return w * m
