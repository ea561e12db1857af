"""Deterministic inputs of the BoT-SORT golden scenarios.  Shared by make_golden.py (which runs the live
reference on them) and by the tests (which re-generate them instead of storing ~10 MB of embeddings;
the fixtures keep a checksum).  No reference access here."""
from __future__ import annotations

import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from yolo_tracking_b200.synth import make_stream  # noqa: E402

BOTSORT_YAML = dict(track_high_thresh=0.33824964456239337, track_low_thresh=0.1, new_track_thresh=0.21144301345190655,
                    track_buffer=60, match_thresh=0.22734550911325851, proximity_thresh=0.5945380911899254,
                    appearance_thresh=0.4818211117541298, frame_rate=30)       # boxmot/configs/botsort.yaml


def _through_camera(sc, dets, nd):
    """The scene as the moving camera sees it: box corners pushed through the accumulated warps (in place)."""
    acc = np.eye(3)
    for f, w in enumerate(camera_warps(sc)):
        acc = np.vstack([w, [0.0, 0.0, 1.0]]) @ acc
        for j in range(nd[f]):
            x1, y1, x2, y2 = dets[f, j, :4]
            p = acc[:2, :2] @ np.array([[x1, x2], [y1, y2]]) + acc[:2, 2:3]
            dets[f, j, :4] = [p[0].min(), p[1].min(), p[0].max(), p[1].max()]


def botsort_inputs(sc):
    """Deterministic inputs of a BoT-SORT scenario (shared with the tests, which re-generate them instead
    of storing ~10 MB of embeddings): dets[F, D, 6], ndets[F], seam features[F, D, emb] (what
    ReIDDetectMultiBackend.get_features returns for the first-round rows: raw / Frobenius norm, scattered
    back to detection rows; zero elsewhere)."""
    dets, nd, embs = make_stream(3, sc["stream"], sc["n_objects"], sc["n_frames"], emb_dim=sc["emb_dim"], **sc["kw"])
    if sc.get("camera"):
        _through_camera(sc, dets, nd)
    if sc.get("classes"):
        rng = np.random.default_rng(sc["stream"] + 7)
        dets[..., 5] = rng.integers(0, sc["classes"], dets.shape[:2]).astype(np.float64) * (dets[..., 4] > 0)
    high = sc["params"].get("track_high_thresh", BOTSORT_YAML["track_high_thresh"])
    feats = np.zeros_like(embs)
    for f in range(sc["n_frames"]):
        rows = np.nonzero(dets[f, :nd[f], 4] > high)[0]
        if len(rows):
            raw = embs[f, rows]
            feats[f, rows] = raw / np.linalg.norm(raw)
    return dets, nd, embs, feats


BOTSORT_SCENARIOS = {
    # BASELINE config 3 shape, scaled down
    "botsort_c3": dict(stream=0, n_objects=40, n_frames=100, emb_dim=512, kw={}, params={}),
    # misses, false positives, three classes (class voting), ByteTrack-like thresholds
    "botsort_churn": dict(stream=903, n_objects=18, n_frames=220, emb_dim=128, kw=dict(miss_prob=0.3, fp_rate=3.0),
                          classes=3, params=dict(track_high_thresh=0.5, new_track_thresh=0.6, match_thresh=0.8,
                                                 proximity_thresh=0.5, appearance_thresh=0.25, track_buffer=30)),
    # appearance disabled: IoU-only association with the XYWH filter
    "botsort_noreid": dict(stream=904, n_objects=25, n_frames=120, emb_dim=32, kw=dict(miss_prob=0.15, fp_rate=2.0),
                           params=dict(with_reid=False)),
    # fuse_first_associate=True (bot_sort.py:300-301): detection scores fused into the first association's IoU cost
    "botsort_fuse": dict(stream=905, n_objects=20, n_frames=150, emb_dim=128, kw=dict(miss_prob=0.2, fp_rate=2.5),
                         params=dict(fuse_first_associate=True, track_high_thresh=0.5, new_track_thresh=0.6, match_thresh=0.8,
                                     proximity_thresh=0.5, appearance_thresh=0.25, track_buffer=30)),
    # a moving camera: STrack.multi_gmc (bot_sort.py:94-111) with a scripted warp per frame; covariances become dense
    "botsort_cam": dict(stream=908, n_objects=16, n_frames=120, emb_dim=128, kw=dict(miss_prob=0.15, fp_rate=1.5), camera=True,
                        params=dict(track_high_thresh=0.5, new_track_thresh=0.6, match_thresh=0.8, proximity_thresh=0.5,
                                    appearance_thresh=0.25, track_buffer=30)),
}


# ----------------------------------------------------------------------------- StrongSORT
STRONGSORT_YAML = dict(max_dist=0.2, max_iou_dist=0.7, max_age=30, n_init=1, nn_budget=100, mc_lambda=0.995,
                       ema_alpha=0.8)                               # boxmot/configs/strongsort.yaml
STRONGSORT_SCENARIOS = {
    "strongsort_c4": dict(stream=0, n_objects=25, n_frames=120, emb_dim=128, kw={}, params={}),
    # misses and false positives, short memory, confirmation after 3 hits, small gallery
    "strongsort_churn": dict(stream=906, n_objects=16, n_frames=200, emb_dim=64, kw=dict(miss_prob=0.3, fp_rate=3.0),
                             params=dict(max_age=8, n_init=3, nn_budget=5, ema_alpha=0.9)),
    # a moving camera: every frame comes with a 2x3 warp (small rotation, zoom, shift) that the detections follow and that
    # Track.camera_update (track.py:129-138) applies to the tracks before the prediction
    "strongsort_cam": dict(stream=907, n_objects=14, n_frames=120, emb_dim=64, kw=dict(miss_prob=0.1, fp_rate=1.0), camera=True,
                           params=dict(n_init=2, nn_budget=20)),
}


def camera_warps(sc):
    """warps[F, 2, 3]: frame f's camera motion relative to frame f - 1 (identity for frame 0), seeded by the scenario."""
    rng = np.random.default_rng(70000 + sc["stream"])
    F = sc["n_frames"]
    ang = rng.normal(0.0, 0.004, F)
    zoom = 1.0 + rng.normal(0.0, 0.003, F)
    shift = rng.normal(0.0, 4.0, (F, 2))
    w = np.zeros((F, 2, 3))
    for f in range(F):
        c, s = np.cos(ang[f]) * zoom[f], np.sin(ang[f]) * zoom[f]
        w[f] = [[c, -s, shift[f, 0]], [s, c, shift[f, 1]]]
    w[0] = np.eye(2, 3)
    return w


def strongsort_inputs(sc):
    """dets[F, D, 6], ndets[F], raw embeddings, seam features (every detection row / Frobenius norm of the frame's matrix)."""
    dets, nd, embs = make_stream(4, sc["stream"], sc["n_objects"], sc["n_frames"], emb_dim=sc["emb_dim"], **sc["kw"])
    if sc.get("camera"):
        _through_camera(sc, dets, nd)
    feats = np.zeros_like(embs)
    for f in range(sc["n_frames"]):
        if nd[f]:
            raw = embs[f, :nd[f]]
            feats[f, :nd[f]] = raw / np.linalg.norm(raw)
    return dets, nd, embs, feats


# ----------------------------------------------------------------------------- DeepOCSORT
# boxmot/configs/deepocsort.yaml as forwarded by tracker_zoo.py:86-98 (the rest are the constructor defaults,
# deep_ocsort.py:308-330)
DEEPOCSORT_YAML = dict(det_thresh=0, max_age=30, min_hits=1, iou_threshold=0.3, delta_t=3, asso_func="giou", inertia=0.2)
DEEPOCSORT_SCENARIOS = {
    # occlusion runs exercise freeze / the quirky unfreeze (deepocsort_kf.py:433-478) / OCR
    "deepocsort_c4": dict(stream=0, n_objects=25, n_frames=120, emb_dim=128, kw=dict(occlusion=True), params={}),
    # misses and false positives, a detection threshold, short memory, output after 3 hits, plain IoU, no adaptive weight
    "deepocsort_churn": dict(stream=911, n_objects=16, n_frames=160, emb_dim=64, kw=dict(miss_prob=0.3, fp_rate=3.0),
                             params=dict(det_thresh=0.3, max_age=8, min_hits=3, asso_func="iou", aw_off=True)),
    # a moving camera: apply_affine_correction (deep_ocsort.py:226-244, deepocsort_kf.py:387-405) on live and frozen state
    "deepocsort_cam": dict(stream=912, n_objects=14, n_frames=120, emb_dim=64, kw=dict(miss_prob=0.15, fp_rate=1.0, occlusion=True),
                           camera=True, params={}),
    # appearance switched off (dets_embs = ones, deep_ocsort.py:386-387), DIoU, output after two hits
    "deepocsort_noemb": dict(stream=913, n_objects=18, n_frames=100, emb_dim=16, kw=dict(miss_prob=0.1, fp_rate=2.0, occlusion=True),
                             params=dict(embedding_off=True, asso_func="diou", min_hits=2)),
    # CIoU, the yaml's appearance weight (never forwarded by the reference's factory), another adaptive-weight floor, a
    # moving camera.  (The centroid similarity cannot be pinned this way: the reference's OCR round calls it without the
    # image size, deep_ocsort.py:463, and raises TypeError as soon as a detection and a track are left over.)
    "deepocsort_ciou": dict(stream=914, n_objects=12, n_frames=100, emb_dim=32, kw=dict(miss_prob=0.1, fp_rate=1.0, occlusion=True),
                            camera=True, params=dict(asso_func="ciou", w_association_emb=0.75, aw_param=0.4)),
}


def deepocsort_inputs(sc, det_thresh):
    """dets[F, D, 6], ndets[F], raw embeddings, and per frame the seam features of the detections that pass
    `conf > det_thresh` (deep_ocsort.py:382-390: the filter runs before get_features, so the Frobenius norm is theirs)."""
    dets, nd, embs = make_stream(4, sc["stream"], sc["n_objects"], sc["n_frames"], emb_dim=sc["emb_dim"], **sc["kw"])
    if sc.get("camera"):
        _through_camera(sc, dets, nd)
    feats = []
    for f in range(sc["n_frames"]):
        keep = dets[f, :nd[f], 4] > det_thresh
        raw = embs[f, :nd[f]][keep].astype(np.float32)
        feats.append(raw / np.linalg.norm(raw) if len(raw) else np.zeros((0, sc["emb_dim"]), dtype=np.float32))
    return dets, nd, embs, feats


# ----------------------------------------------------------------------------- HybridSORT
# boxmot/configs/hybridsort.yaml as forwarded by tracker_zoo.py:100-115 (use_byte is not forwarded; everything else is
# fixed in HybridSORT.__init__, hybridsort.py:337-364)
HYBRIDSORT_YAML = dict(det_thresh=0, max_age=30, min_hits=1, iou_threshold=0.3, delta_t=3, asso_func="giou", inertia=0.2)
HYBRIDSORT_SCENARIOS = {
    # occlusion runs exercise freeze / unfreeze (hybridsort_kf.py:390-436, score read as aspect ratio) / OCR
    "hybridsort_c4": dict(stream=0, n_objects=25, n_frames=120, emb_dim=128, kw=dict(occlusion=True), params={}),
    # misses and false positives, a detection threshold (the class / last column then come from the UNFILTERED row of the
    # same index, hybridsort.py:396-404), short memory, output after 3 hits, plain IoU
    "hybridsort_churn": dict(stream=921, n_objects=16, n_frames=160, emb_dim=64, kw=dict(miss_prob=0.3, fp_rate=3.0),
                             params=dict(det_thresh=0.3, max_age=8, min_hits=3, asso_func="iou")),
    # crowded, long occlusions, DIoU, a two-frame velocity window
    "hybridsort_diou": dict(stream=922, n_objects=40, n_frames=100, emb_dim=32, kw=dict(miss_prob=0.1, fp_rate=2.0, occlusion=True),
                            params=dict(asso_func="diou", delta_t=2, min_hits=2)),
    # two classes (wide boxes are class 1): HybridSORT is the one tracker whose update runs under the PerClassDecorator with
    # per_class = True (hybridsort.py:346, :372; boxmot/utils/__init__.py:22-61) - one full update per class and frame
    "hybridsort_2cls": dict(stream=923, n_objects=14, n_frames=90, emb_dim=32, kw=dict(miss_prob=0.1, fp_rate=1.0, occlusion=True),
                            two_classes=True, params=dict(max_age=10)),
}


def hybridsort_inputs(sc, det_thresh, full=False):
    """dets[F, D, 6], ndets[F], raw embeddings, and per frame the seam features of the detections that pass
    `conf > det_thresh`.  The reference extracts features for EVERY detection (hybridsort.py:394) and the seam divides by
    the Frobenius norm of that whole matrix; only the rows above det_thresh are used (:403).  `full`: the rows of all
    detections instead (what the device path is handed)."""
    dets, nd, embs = make_stream(4, sc["stream"], sc["n_objects"], sc["n_frames"], emb_dim=sc["emb_dim"], **sc["kw"])
    if sc.get("two_classes"):
        dets[:, :, 5] = (dets[:, :, 2] - dets[:, :, 0] > 60.0).astype(np.float64)
    feats = []
    for f in range(sc["n_frames"]):
        keep = dets[f, :nd[f], 4] > det_thresh
        raw = embs[f, :nd[f]].astype(np.float32)
        rows = np.zeros((nd[f], sc["emb_dim"]), dtype=np.float32)
        for c in np.unique(dets[f, :nd[f], 5]):               # the seam normalises the matrix of ONE get_features call: one class
            m = dets[f, :nd[f], 5] == c
            rows[m] = raw[m] / np.linalg.norm(raw[m])
        feats.append(rows if full else rows[keep])
    return dets, nd, embs, feats


def mot_feats(seq_index, frame, n, dim=32):
    """Seeded stand-in embeddings for the detections of frame `frame` of MOT17-mini sequence `seq_index` (raw, before the
    seam's whole-matrix normalisation)."""
    return np.random.default_rng(50000 + 1000 * seq_index + frame).normal(0.0, 1.0, (n, dim)).astype(np.float32)


# ----------------------------------------------------------------------------- BASELINE config sizes (rows-only fixtures)
# One stream per BASELINE.json config at its stated size, run through the live reference; the fixtures (full_*.npz) keep
# the result rows only - ids and det_ind of every frame, boxes of every `box_every`-th frame - so they stay small.
FULLSIZE = {
    # config 1: ByteTrack, 1 stream, 1000 frames x ~50 detections (53 objects)
    "full_bytetrack_c1": dict(kind="bytetrack", config=1, stream=0, n_objects=53, n_frames=1000, emb_dim=0, kw={}, box_every=10),
    # config 2: OC-SORT, 100 objects with occlusion runs (one of the 64 streams)
    "full_ocsort_c2": dict(kind="ocsort", config=2, stream=7, n_objects=100, n_frames=200, emb_dim=0, kw=dict(occlusion=True), box_every=5),
    # config 3: BoT-SORT, 100 objects, 512-d embeddings (one of the 256 streams)
    "full_botsort_c3": dict(kind="botsort", config=3, stream=11, n_objects=100, n_frames=200, emb_dim=512, kw={}, box_every=5),
    # config 4: 200 tracks x 200 detections with 512-d embeddings through DeepOCSORT and StrongSORT
    "full_deepocsort_c4": dict(kind="deepocsort", config=4, stream=3, n_objects=190, n_frames=100, emb_dim=512, kw={}, box_every=5),
    "full_strongsort_c4": dict(kind="strongsort", config=4, stream=3, n_objects=200, n_frames=60, emb_dim=512, kw={}, box_every=5),
    "full_hybridsort_c4": dict(kind="hybridsort", config=4, stream=3, n_objects=190, n_frames=100, emb_dim=512, kw={}, box_every=5),
}


def fullsize_inputs(sc):
    """dets[F, D, 6], ndets[F], raw embeddings (or None) of a full-size scenario."""
    return make_stream(sc["config"], sc["stream"], sc["n_objects"], sc["n_frames"], emb_dim=sc["emb_dim"], **sc["kw"])
