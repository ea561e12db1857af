"""Small end-to-end exercise of every kernel, meant to run under `compute-sanitizer --tool memcheck` (or racecheck /
initcheck, one tool per gpurun call): a few frames of each tracker on a handful of streams plus every operator."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolo_tracking_b200 import _lib, _ops  # noqa: E402
from yolo_tracking_b200.batch import BatchedTracker  # noqa: E402
from yolo_tracking_b200.synth import make_batch  # noqa: E402

F = int(sys.argv[1]) if len(sys.argv) > 1 else 12
ONLY = sys.argv[2].split(",") if len(sys.argv) > 2 else None          # tracker kinds to run (then the operators are skipped)
for kind, cap, kw, params in (
        ("bytetrack", 64, dict(miss_prob=0.2, fp_rate=3.0), dict(track_thresh=0.5, match_thresh=0.8, track_buffer=5, frame_rate=30)),
        ("bytetrack", 224, {}, dict(track_thresh=0.5, match_thresh=0.8, track_buffer=30, frame_rate=30)),
        ("ocsort", 128, dict(occlusion=True, miss_prob=0.2), dict(det_thresh=0.4, max_age=5, min_hits=1, asso_threshold=0.3, delta_t=3,
                                                                  asso_func="giou", inertia=0.2, use_byte=True)),
        ("botsort", 128, dict(miss_prob=0.2, fp_rate=2.0, emb_dim=128), dict(track_high_thresh=0.5, track_low_thresh=0.1, new_track_thresh=0.6,
                                                                            track_buffer=5, match_thresh=0.8, proximity_thresh=0.5,
                                                                            appearance_thresh=0.25, frame_rate=30)),
        ("hybridsort", 64, dict(occlusion=True, miss_prob=0.2, fp_rate=2.0, emb_dim=36), dict(det_thresh=0.3, max_age=5, min_hits=1, iou_threshold=0.3,
                                                                                             delta_t=3, asso_func="giou", inertia=0.2)),
        ("hybridsort", 128, dict(occlusion=True, emb_dim=512), dict(det_thresh=0.0, max_age=30, min_hits=1, iou_threshold=0.3, delta_t=3,
                                                                    asso_func="diou", inertia=0.2))):
    if ONLY and kind not in ONLY:
        continue
    n_obj = 150 if cap == 224 else 30
    dets, nd, embs = make_batch(7, 3, n_obj, F, dmax=cap, **kw)
    trk = BatchedTracker(kind, 3, max_tracks=cap, max_dets=cap, feat_dim=embs.shape[-1] if embs is not None else 0, **params)
    rows = 0
    for f in range(F):
        feats = None
        if embs is not None:
            feats = np.ascontiguousarray(embs[f] / 10.0)
        out, nout = trk.update_batch(np.ascontiguousarray(dets[f]), np.ascontiguousarray(nd[f]), feats=feats, img_hw=(1080, 1920))
        rows += int(nout.sum())
    trk.state(1)
    trk.sync()
    trk.close()
    print(kind, cap, "rows", rows, flush=True)
if ONLY:
    sys.exit(0)
rng = np.random.default_rng(0)
z = np.stack([rng.uniform(100, 900, 50), rng.uniform(100, 900, 50), rng.uniform(0.3, 0.8, 50), rng.uniform(60, 220, 50)], axis=1)
for kind in (_lib.KF_XYAH, _lib.KF_XYWH, _lib.KF_XYAH_CONF):
    m, c = _ops.kf_initiate(kind, z)
    m, c = _ops.kf_predict(kind, m, c)
    _ops.kf_project(kind, m, c, 0.5 if kind == _lib.KF_XYAH_CONF else None)
    m, c = _ops.kf_update(kind, m, c, z + 1.0, 0.5 if kind == _lib.KF_XYAH_CONF else None)
    _ops.kf_gating_distance(kind, m, c, z[:17] + 2.0, False, "maha", 0.5 if kind == _lib.KF_XYAH_CONF else None)
    _ops.gate_cost(kind, rng.random((50, 17)), m, c, z[:17] + 2.0, fuse=True)
a = np.concatenate([z[:, :2] - 20, z[:, :2] + 30], axis=1)
for name in ("iou", "giou", "diou", "ciou", "centroid"):
    _ops.box_similarity(name, a, a[:23] + 3.0, 1920, 1080)
_ops.iou_distance(a, a[:23] + 3.0)
_ops.embedding_distance(rng.standard_normal((37, 96)), rng.standard_normal((29, 96)))
_ops.appearance_cost(rng.standard_normal((2, 150, 128)), rng.standard_normal((2, 70, 128)), 0.5, 0.45, 1.0, gate=rng.random((2, 150, 70)) < 0.3)
_ops.lapjv(rng.random((3, 40, 55)), 0.6)
_ops.lapjv(-rng.random((3, 40, 55)))
_ops.linear_sum_assignment(np.minimum(rng.random((2, 30, 45)), 0.4))
_ops.kf_apply_warp(m, c, np.array([[1.01, 0.02, 3.0], [-0.02, 0.99, -1.0]]))
_ops.aw_max_metric(rng.random((2, 33, 47)), 0.75)
gal = rng.standard_normal((2, 6, 100, 128)).astype(np.float32)
_ops.gallery_cost(gal, rng.integers(0, 101, (2, 6)), gal[:, :, 3] + 0.1 * rng.standard_normal((2, 6, 128)).astype(np.float32), 0.2)
_ops.nn_cosine_distance([gal[0, 0, :5], gal[0, 1, :9]], gal[0, :, 3])
# DeepOCSORT operators (csrc/kf8.cu): a track count that is not a multiple of the 32 / 64 tracks a CTA stages
x8 = np.concatenate([rng.uniform(100, 900, (77, 2)), rng.uniform(30, 200, (77, 2)), rng.normal(0, 2, (77, 4))], axis=1)
P8 = np.stack([np.diag(rng.uniform(1, 30, 8)) for _ in range(77)])
x8, P8 = _ops.kf8_predict(x8, P8)
x8, P8 = _ops.kf8_predict(x8, P8, unit_q=True)
x8, P8 = _ops.kf8_update(x8, P8, x8[:, :4] + 1.0, x8[:, 2:4])
x8, P8 = _ops.kf8_update(x8, P8, x8[:, :4] + 1.0)
_ops.kf8_oru(x8, P8, x8[:, :4], x8[:, :4] + 5.0, rng.integers(1, 31, 77))
d5 = np.concatenate([a[:23] + 3.0, rng.uniform(0.1, 1, (23, 1))], axis=1)
p5 = np.concatenate([a, rng.uniform(-1, 1, (50, 1))], axis=1)
_ops.ocm_cost(rng.random((23, 50)), d5, rng.normal(0, 1, (50, 2)), p5, 0.2, rng.random((23, 50)))
_ops.ocm_cost(rng.random((23, 50)))
_ops.dot_matrix(rng.standard_normal((23, 96)), rng.standard_normal((50, 96)))
# crowded scene: the pair list and the edge cache of the ByteTrack step overflow (bitmask graph, union-find solver)
dets, nd, _ = make_batch(7, 2, 40, 8, dmax=64, fp_rate=0.5)
ctr, half = 0.5 * (dets[..., :2] + dets[..., 2:4]), 0.5 * (dets[..., 2:4] - dets[..., :2])
dets[..., :2], dets[..., 2:4] = ctr * 0.1 - half, ctr * 0.1 + half
dets[np.arange(64)[None, None, :] >= nd[:, :, None]] = 0.0
trk = BatchedTracker("bytetrack", 2, max_tracks=64, max_dets=64, track_thresh=0.5, match_thresh=0.8, track_buffer=30, frame_rate=30)
for f in range(8):
    trk.update_batch(np.ascontiguousarray(dets[f]), np.ascontiguousarray(nd[f]))
trk.sync()
trk.close()
print("ops ok", flush=True)
