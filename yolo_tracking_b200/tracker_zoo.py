"""Factory with the reference's signature (boxmot/tracker_zoo.py:10-118): YAML -> tracker.

The same YAML keys are read and the same subset is forwarded to each constructor
(tracker_zoo.py:43-81).  `reid_weights`, `half` and `per_class` are accepted for signature
compatibility; appearance features are passed to update() directly (the ReID networks are
out of scope), `device` selects the CUDA device.  Unknown type: ValueError (the reference
prints 'No such tracker' and exits the interpreter).
"""
from __future__ import annotations

from pathlib import Path
from types import SimpleNamespace

import yaml

_CONFIGS = Path(__file__).resolve().parent / "configs"


def get_tracker_config(tracker_type):
    return _CONFIGS / (tracker_type + ".yaml")


def create_tracker(tracker_type, tracker_config, reid_weights=None, device=0, half=False, per_class=False, **capacity):
    with open(tracker_config, "r") as f:
        cfg = SimpleNamespace(**yaml.load(f.read(), Loader=yaml.FullLoader))
    if tracker_type == "bytetrack":
        from .trackers.bytetrack import BYTETracker
        return BYTETracker(track_thresh=cfg.track_thresh, match_thresh=cfg.match_thresh,
                           track_buffer=cfg.track_buffer, frame_rate=cfg.frame_rate, device=device, **capacity)
    if tracker_type == "ocsort":
        from .trackers.ocsort import OCSort
        return OCSort(per_class, det_thresh=cfg.det_thresh, max_age=cfg.max_age, min_hits=cfg.min_hits,
                      asso_threshold=cfg.iou_thresh, delta_t=cfg.delta_t, asso_func=cfg.asso_func,
                      inertia=cfg.inertia, use_byte=cfg.use_byte, device=device, **capacity)
    if tracker_type == "botsort":
        from .trackers.botsort import BoTSORT
        return BoTSORT(reid_weights, device, half, track_high_thresh=cfg.track_high_thresh,
                       track_low_thresh=cfg.track_low_thresh, new_track_thresh=cfg.new_track_thresh,
                       track_buffer=cfg.track_buffer, match_thresh=cfg.match_thresh,
                       proximity_thresh=cfg.proximity_thresh, appearance_thresh=cfg.appearance_thresh,
                       cmc_method=cfg.cmc_method, frame_rate=cfg.frame_rate, **capacity)
    if tracker_type == "strongsort":
        from .trackers.strongsort import StrongSORT
        return StrongSORT(reid_weights, device, half, max_dist=cfg.max_dist, max_iou_dist=cfg.max_iou_dist, max_age=cfg.max_age,
                          n_init=cfg.n_init, nn_budget=cfg.nn_budget, mc_lambda=cfg.mc_lambda, ema_alpha=cfg.ema_alpha, **capacity)
    if tracker_type == "deepocsort":
        from .trackers.deepocsort import DeepOCSort
        return DeepOCSort(reid_weights, device, half, per_class, det_thresh=cfg.det_thresh, max_age=cfg.max_age,
                          min_hits=cfg.min_hits, iou_threshold=cfg.iou_thresh, delta_t=cfg.delta_t, asso_func=cfg.asso_func,
                          inertia=cfg.inertia, **capacity)
    if tracker_type == "hybridsort":
        from .trackers.hybridsort import HybridSORT
        return HybridSORT(reid_weights, device, half, det_thresh=cfg.det_thresh, max_age=cfg.max_age, min_hits=cfg.min_hits,
                          iou_threshold=cfg.iou_thresh, delta_t=cfg.delta_t, asso_func=cfg.asso_func, inertia=cfg.inertia, **capacity)
    raise ValueError(f"No such tracker: {tracker_type!r} (built: bytetrack, ocsort, botsort, strongsort, deepocsort, hybridsort)")
