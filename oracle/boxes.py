"""ORACLE (test infrastructure, not product code): box-format conversions and pairwise box
similarities, restated in numpy from the reference.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this package.  Parity pinned by tests/golden/*.npz (generated from the live
reference by tests/golden/make_golden.py).

Follows (reference file:line):
  boxmot/utils/ops.py:7-21    xyxy2xywh      boxmot/utils/ops.py:24-40   xywh2xyxy
  boxmot/utils/ops.py:43-58   xywh2tlwh      boxmot/utils/ops.py:87-97   tlwh2xyah
  boxmot/utils/iou.py:6-25    iou_batch      boxmot/utils/iou.py:28-62   giou_batch
  boxmot/utils/iou.py:65-105  diou_batch     boxmot/utils/iou.py:108-161 ciou_batch
  boxmot/utils/iou.py:164-188 centroid_batch
The order of floating-point operations is kept (it decides the last bit of every cost).
"""
from __future__ import annotations

import numpy as np


def xyxy_to_xywh(b):
    b = np.asarray(b, dtype=np.float64)
    out = np.empty_like(b)
    out[..., 0] = (b[..., 0] + b[..., 2]) / 2
    out[..., 1] = (b[..., 1] + b[..., 3]) / 2
    out[..., 2] = b[..., 2] - b[..., 0]
    out[..., 3] = b[..., 3] - b[..., 1]
    return out


def xywh_to_xyxy(b):
    b = np.asarray(b, dtype=np.float64)
    out = np.empty_like(b)
    out[..., 0] = b[..., 0] - b[..., 2] / 2
    out[..., 1] = b[..., 1] - b[..., 3] / 2
    out[..., 2] = b[..., 0] + b[..., 2] / 2
    out[..., 3] = b[..., 1] + b[..., 3] / 2
    return out


def xywh_to_tlwh(b):
    b = np.asarray(b, dtype=np.float64)
    out = b.copy()
    out[..., 0] = b[..., 0] - b[..., 2] / 2.0
    out[..., 1] = b[..., 1] - b[..., 3] / 2.0
    return out


def tlwh_to_xyah(b):
    b = np.asarray(b, dtype=np.float64)
    out = b.copy()
    out[..., 0] = b[..., 0] + (b[..., 2] / 2)
    out[..., 1] = b[..., 1] + (b[..., 3] / 2)
    out[..., 2] = b[..., 2] / b[..., 3]
    return out


def det_xyah(xyxy):
    """STrack.__init__ chain xyxy -> xywh -> tlwh -> xyah (byte_tracker.py:16-18)."""
    return tlwh_to_xyah(xywh_to_tlwh(xyxy_to_xywh(xyxy)))


def _pair(a, b):
    a = np.asarray(a, dtype=np.float64).reshape(-1, 4)[:, None, :]
    b = np.asarray(b, dtype=np.float64).reshape(-1, 4)[None, :, :]
    return a, b


def _inter_iou(a, b):
    ix1 = np.maximum(a[..., 0], b[..., 0])
    iy1 = np.maximum(a[..., 1], b[..., 1])
    ix2 = np.minimum(a[..., 2], b[..., 2])
    iy2 = np.minimum(a[..., 3], b[..., 3])
    iw = np.maximum(0.0, ix2 - ix1)
    ih = np.maximum(0.0, iy2 - iy1)
    inter = iw * ih
    with np.errstate(invalid="ignore", divide="ignore"):
        iou = inter / ((a[..., 2] - a[..., 0]) * (a[..., 3] - a[..., 1])
                       + (b[..., 2] - b[..., 0]) * (b[..., 3] - b[..., 1]) - inter)
    return inter, iou


def iou(a, b):
    a, b = _pair(a, b)
    return _inter_iou(a, b)[1]


def giou(a, b):
    a, b = _pair(a, b)
    inter, v = _inter_iou(a, b)
    ew = np.maximum(a[..., 2], b[..., 2]) - np.minimum(a[..., 0], b[..., 0])
    eh = np.maximum(a[..., 3], b[..., 3]) - np.minimum(a[..., 1], b[..., 1])
    assert (ew > 0).all() and (eh > 0).all()
    enclose = ew * eh
    g = v - (enclose - inter) / enclose
    return (g + 1.0) / 2.0


def _centre_terms(a, b):
    cxa = (a[..., 0] + a[..., 2]) / 2.0
    cya = (a[..., 1] + a[..., 3]) / 2.0
    cxb = (b[..., 0] + b[..., 2]) / 2.0
    cyb = (b[..., 1] + b[..., 3]) / 2.0
    inner = (cxa - cxb) ** 2 + (cya - cyb) ** 2
    ex1 = np.minimum(a[..., 0], b[..., 0])
    ey1 = np.minimum(a[..., 1], b[..., 1])
    ex2 = np.maximum(a[..., 2], b[..., 2])
    ey2 = np.maximum(a[..., 3], b[..., 3])
    outer = (ex2 - ex1) ** 2 + (ey2 - ey1) ** 2
    return inner, outer


def diou(a, b):
    a, b = _pair(a, b)
    _, v = _inter_iou(a, b)
    inner, outer = _centre_terms(a, b)
    return (v - inner / outer + 1) / 2.0


def ciou(a, b):
    a, b = _pair(a, b)
    _, v = _inter_iou(a, b)
    inner, outer = _centre_terms(a, b)
    wa = a[..., 2] - a[..., 0]
    ha = a[..., 3] - a[..., 1] + 1.0
    wb = b[..., 2] - b[..., 0]
    hb = b[..., 3] - b[..., 1] + 1.0
    dth = np.arctan(wb / hb) - np.arctan(wa / ha)
    vv = (4 / (np.pi ** 2)) * (dth ** 2)
    alpha = vv / ((1 - v) + vv)
    return (v - inner / outer - alpha * vv + 1) / 2.0


def centroid(a, b, w, h):
    a = np.asarray(a, dtype=np.float64).reshape(-1, 4)
    b = np.asarray(b, dtype=np.float64).reshape(-1, 4)
    ca = np.stack(((a[:, 0] + a[:, 2]) / 2, (a[:, 1] + a[:, 3]) / 2), axis=-1)[:, None, :]
    cb = np.stack(((b[:, 0] + b[:, 2]) / 2, (b[:, 1] + b[:, 3]) / 2), axis=-1)[None, :, :]
    dist = np.sqrt(np.sum((ca - cb) ** 2, axis=-1))
    return 1 - dist / np.sqrt(w ** 2 + h ** 2)


ASSO = {"iou": iou, "giou": giou, "diou": diou, "ciou": ciou, "centroid": centroid}


def similarity(name, a, b, w=None, h=None):
    """run_asso_func (iou.py:191-212): only centroid uses the frame size."""
    if name == "centroid":
        return centroid(a, b, w, h)
    return ASSO[name](a, b)
