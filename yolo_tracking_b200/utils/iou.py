"""boxmot/utils/iou.py on the GPU: same function names and argument meaning."""
from __future__ import annotations

from .. import _ops


def iou_batch(bboxes1, bboxes2):
    return _ops.box_similarity("iou", bboxes1, bboxes2)


def giou_batch(bboxes1, bboxes2):
    return _ops.box_similarity("giou", bboxes1, bboxes2)


def diou_batch(bboxes1, bboxes2):
    return _ops.box_similarity("diou", bboxes1, bboxes2)


def ciou_batch(bboxes1, bboxes2):
    return _ops.box_similarity("ciou", bboxes1, bboxes2)


def centroid_batch(bboxes1, bboxes2, w, h):
    return _ops.box_similarity("centroid", bboxes1, bboxes2, w, h)


_ASSO = {"iou": iou_batch, "giou": giou_batch, "ciou": ciou_batch, "diou": diou_batch, "centroid": centroid_batch}


def get_asso_func(asso_mode):
    return _ASSO[asso_mode]


def run_asso_func(func, *args):
    """iou.py:191-212: the box functions take two box arrays; centroid also takes (w, h)."""
    if func not in _ASSO.values():
        raise ValueError("Invalid function specified. Must be either '(g,d,c, )iou_batch' or 'centroid_batch'.")
    if len(args) != 4:
        raise ValueError("Invalid arguments. Expected two bounding boxes and two size parameters.")
    if func is centroid_batch:
        return func(*args)
    return func(*args[0:2])
