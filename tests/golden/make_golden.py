"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) in this
container through tests/golden/ref_harness.py.  Run:  python tests/golden/make_golden.py [names]

The .npz files are committed; this script cannot run on the GPU box (no /root/reference).
Every fixture stores its inputs next to the reference's outputs so it is self-contained.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import ref_harness as rh  # noqa: E402
from yolo_tracking_b200.synth import make_stream  # noqa: E402


def _save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"{name}.npz  {os.path.getsize(path) / 1024:.0f} KiB")


def _ragged(list_of_arrays, width):
    """-> (flat[n,width], offsets[len+1])"""
    offs = np.zeros(len(list_of_arrays) + 1, dtype=np.int64)
    rows = []
    for i, a in enumerate(list_of_arrays):
        a = np.asarray(a, dtype=np.float64).reshape(-1, width)
        rows.append(a)
        offs[i + 1] = offs[i] + len(a)
    return (np.concatenate(rows, axis=0) if rows else np.zeros((0, width))), offs


# ----------------------------------------------------------------------------- Kalman filters
def gen_kf():
    rh.install()
    from boxmot.motion.kalman_filters.bytetrack_kf import KalmanFilter as KFxyah
    from boxmot.motion.kalman_filters.botsort_kf import KalmanFilter as KFxywh
    from boxmot.motion.kalman_filters.strongsort_kf import KalmanFilter as KFconf
    rng = np.random.default_rng(11)
    n, steps = 48, 6
    for name, cls in (("xyah", KFxyah), ("xywh", KFxywh), ("xyah_conf", KFconf)):
        kf = cls()
        z0 = np.stack([rng.uniform(100, 1800, n), rng.uniform(100, 1000, n),
                       rng.uniform(0.3, 0.8, n) if name != "xywh" else rng.uniform(30, 90, n),
                       rng.uniform(60, 220, n)], axis=1)
        out = dict(z0=z0)
        mean = np.zeros((n, 8))
        cov = np.zeros((n, 8, 8))
        for i in range(n):
            mean[i], cov[i] = kf.initiate(z0[i])
        out["init_mean"], out["init_cov"] = mean.copy(), cov.copy()
        zs = np.zeros((steps, n, 4))
        confs = rng.uniform(0.3, 0.95, (steps, n))
        pm_all, pc_all = np.zeros((steps, n, 8)), np.zeros((steps, n, 8, 8))
        um_all, uc_all = np.zeros((steps, n, 8)), np.zeros((steps, n, 8, 8))
        prj_m, prj_c = np.zeros((steps, n, 4)), np.zeros((steps, n, 4, 4))
        for s in range(steps):
            if name == "xyah_conf":
                for i in range(n):
                    mean[i], cov[i] = kf.predict(mean[i], cov[i])
            else:
                mean, cov = kf.multi_predict(mean, cov)
            pm_all[s], pc_all[s] = mean, cov
            zs[s] = mean[:, :4] + rng.normal(0, 1, (n, 4)) * np.array([2, 2, 0.01 if name != "xywh" else 2, 2])
            for i in range(n):
                if name == "xyah_conf":
                    prj_m[s, i], prj_c[s, i] = kf.project(mean[i], cov[i], confs[s, i])
                    mean[i], cov[i] = kf.update(mean[i], cov[i], zs[s, i], confs[s, i])
                else:
                    prj_m[s, i], prj_c[s, i] = kf.project(mean[i], cov[i])
                    mean[i], cov[i] = kf.update(mean[i], cov[i], zs[s, i])
            um_all[s], uc_all[s] = mean, cov
        out.update(z=zs, conf=confs, pred_mean=pm_all, pred_cov=pc_all, proj_mean=prj_m,
                   proj_cov=prj_c, upd_mean=um_all, upd_cov=uc_all)
        # gating distance: every track of the last predicted state against 40 measurements
        D = 40
        pick = rng.integers(0, n, D)
        meas = pm_all[-1, pick, :4] + rng.normal(0, 1, (D, 4)) * np.array([6, 6, 0.02 if name != "xywh" else 6, 6])
        g4 = np.zeros((n, D))
        g2 = np.zeros((n, D))
        gg = np.zeros((n, D))
        for i in range(n):
            g4[i] = kf.gating_distance(pm_all[-1, i], pc_all[-1, i], meas.copy(), False)
            g2[i] = kf.gating_distance(pm_all[-1, i], pc_all[-1, i], meas.copy(), True)
            if name != "xyah_conf":
                gg[i] = kf.gating_distance(pm_all[-1, i], pc_all[-1, i], meas.copy(), False, "gaussian")
        out.update(gate_meas=meas, gate_maha4=g4, gate_maha2=g2, gate_gauss=gg)
        _save("kf_" + name, **out)


# ----------------------------------------------------------------------------- pairwise costs
def gen_costs():
    rh.install()
    from boxmot.utils import iou as riou
    from boxmot.utils import matching as rm
    rng = np.random.default_rng(12)

    def rand_boxes(n):
        c = rng.uniform(0, 400, (n, 2))
        wh = rng.uniform(20, 160, (n, 2))
        return np.concatenate([c - wh / 2, c + wh / 2], axis=1)
    a, b = rand_boxes(37), rand_boxes(29)
    out = dict(a=a, b=b, iou=riou.iou_batch(a, b), giou=riou.giou_batch(a, b),
               diou=riou.diou_batch(a, b), ciou=riou.ciou_batch(a, b),
               centroid=riou.centroid_batch(a, b, 640, 480))
    score = rng.uniform(0.1, 1.0, 29)

    class _D:
        def __init__(self, s): self.score = s
    cost = rm.iou_distance(list(a), list(b))
    out["iou_distance"] = cost
    out["score"] = score
    out["fuse_score"] = rm.fuse_score(cost.copy(), [_D(s) for s in score])
    # embedding_distance (matching.py:145-167): fp32 cast, cdist cosine, clamp at 0
    ta = rng.standard_normal((37, 512))
    tb = rng.standard_normal((29, 512)) + 0.5 * ta[rng.integers(0, 37, 29)]

    class _T:
        def __init__(self, f): self.smooth_feat = f; self.curr_feat = f
    out["feat_a"], out["feat_b"] = ta, tb
    out["embedding_distance"] = rm.embedding_distance([_T(f) for f in ta], [_T(f) for f in tb])
    _save("costs", **out)


# ----------------------------------------------------------------------------- camera-motion warp, adaptive appearance weight
def gen_aux():
    """STrack.multi_gmc (bot_sort.py:95-111) on filter states that went through predict / update rounds, and
    compute_aw_max_metric (association.py:79-108) on similarity matrices with the shapes / degeneracies it branches on."""
    rh.install()
    from boxmot.motion.kalman_filters.botsort_kf import KalmanFilter as KFxywh
    from boxmot.trackers.botsort.bot_sort import STrack
    from boxmot.utils.association import compute_aw_max_metric
    rng = np.random.default_rng(21)
    kf = KFxywh()
    n = 40
    mean = np.zeros((n, 8))
    cov = np.zeros((n, 8, 8))
    for i in range(n):
        z = np.array([rng.uniform(100, 1800), rng.uniform(100, 1000), rng.uniform(30, 90), rng.uniform(60, 220)])
        m, c = kf.initiate(z)
        for _ in range(3):
            m, c = kf.predict(m, c)
            m, c = kf.update(m, c, m[:4] + rng.normal(0, 1.5, 4))
        mean[i], cov[i] = m, c
    ang = 0.03
    warps = np.stack([np.eye(2, 3),
                      np.array([[np.cos(ang), -np.sin(ang), 3.5], [np.sin(ang), np.cos(ang), -2.25]]),
                      np.array([[1.02, 0.01, -7.0], [-0.015, 0.98, 4.0]])])

    class _S:                                             # the two attributes multi_gmc reads and writes
        def __init__(self, m, c): self.mean, self.covariance = m, c
    out = dict(mean=mean, cov=cov, warps=warps)
    for k, H in enumerate(warps):
        ss = [_S(mean[i].copy(), cov[i].copy()) for i in range(n)]
        STrack.multi_gmc(ss, H)
        out[f"gmc_mean{k}"] = np.stack([x.mean for x in ss])
        out[f"gmc_cov{k}"] = np.stack([x.covariance for x in ss])
    mats = []
    a = rng.uniform(0.0, 1.0, (17, 23))
    a[3] = 0.0                                            # a row whose largest entry is 0
    a[:, 5] = 0.0                                         # a zero column
    a[7, 2] = a[7, 9] = a[7].max() + 0.1                  # tied top-2 in a row
    mats.append(a)
    mats.append(rng.uniform(-0.2, 1.0, (1, 9)))           # a single row: column weights untouched
    mats.append(rng.uniform(-0.2, 1.0, (6, 1)))           # a single column
    mats.append(rng.uniform(0.0, 1.0, (96, 120)))
    for k, m in enumerate(mats):
        out[f"aw_in{k}"] = m
        out[f"aw_out{k}"] = compute_aw_max_metric(m.copy(), 0.75, 0.5)
    out["aw_out0_b"] = compute_aw_max_metric(mats[0].copy(), 0.4, 0.3)
    _save("aux_ops", **out)


# ----------------------------------------------------------------------------- ByteTrack
def _bt_snapshot(trk):
    ts = trk.tracked_stracks + trk.lost_stracks
    ints = np.array([[t.track_id, t.state, int(t.is_activated), t.frame_id, t.start_frame,
                      t.tracklet_len] for t in ts], dtype=np.int32).reshape(-1, 6)
    mean = np.stack([t.mean for t in ts]) if ts else np.zeros((0, 8))
    cov = np.stack([t.covariance for t in ts]) if ts else np.zeros((0, 8, 8))
    return len(trk.tracked_stracks), len(trk.lost_stracks), ints, mean, cov


def gen_bytetrack():
    scenarios = {
        # BASELINE config 1 shape (53 objects ~ 50 dets/frame), shortened to keep the file small
        "bytetrack_c1": dict(config=1, stream=0, n_objects=53, n_frames=150, kw={}),
        # heavy misses + false positives: exercises lost / re-found / removed-lag / sticky-removed
        "bytetrack_churn": dict(config=1, stream=901, n_objects=20, n_frames=300,
                                kw=dict(miss_prob=0.3, fp_rate=3.0)),
    }
    for name, sc in scenarios.items():
        dets, nd, _ = make_stream(sc["config"], sc["stream"], sc["n_objects"], sc["n_frames"], **sc["kw"])
        trk = rh.make_tracker("bytetrack")
        outs, ints, counts, means, covs, pool, cov_frames = [], [], [], [], [], [], []
        for f in range(sc["n_frames"]):
            pool.append(len(trk.tracked_stracks) + len(trk.lost_stracks))
            o = trk.update(dets[f, :nd[f]], None)
            outs.append(o)
            nt, nl, ii, m, c = _bt_snapshot(trk)
            counts.append((nt, nl))
            ints.append(ii)
            means.append(m)
            if f % 10 == 9 or f == sc["n_frames"] - 1:      # dense covariances every 10th frame
                covs.append(c.reshape(-1, 64))
                cov_frames.append(f)
        out_flat, out_offs = _ragged(outs, 8)
        int_flat, int_offs = _ragged(ints, 6)
        mean_flat, _ = _ragged(means, 8)
        cov_flat, _ = _ragged(covs, 64)
        _save(name, dets=dets, ndets=nd, out=out_flat, out_offs=out_offs,
              rec=int_flat.astype(np.int32), rec_offs=int_offs, counts=np.array(counts, dtype=np.int32),
              mean=mean_flat, cov=cov_flat.astype(np.float64), cov_frames=np.array(cov_frames, dtype=np.int32), pool=np.array(pool, dtype=np.int32),
              params=np.array([0.5, 0.8, 30, 30], dtype=np.float64))
    # the reference's own known-answer test input (tests/test_python.py:165-185)
    trk = rh.make_tracker("bytetrack")
    det = np.array([[144, 212, 578, 480, 0.82, 0], [425, 281, 576, 472, 0.86, 65]], dtype=np.float64)
    outs = [trk.update(det, None) for _ in range(3)]
    _save("bytetrack_2box", det=det, out=np.stack(outs))


# ----------------------------------------------------------------------------- OC-SORT
def _oc_snapshot(trk):
    ts = trk.trackers
    ints = np.array([[t.id, t.age, t.time_since_update, t.hits, t.hit_streak, int(t.kf.observed)] for t in ts],
                    dtype=np.int32).reshape(-1, 6)
    x = np.stack([t.kf.x[:, 0] for t in ts]) if ts else np.zeros((0, 7))
    P = np.stack([t.kf.P.reshape(49) for t in ts]) if ts else np.zeros((0, 49))
    vel = np.array([t.velocity if t.velocity is not None else np.zeros(2) for t in ts]).reshape(-1, 2)
    last = np.array([t.last_observation for t in ts], dtype=np.float64).reshape(-1, 5)
    return ints, x, P, vel, last


def gen_ocsort():
    rh.install()
    from boxmot.trackers.ocsort.ocsort import OCSort
    base = dict(det_thresh=0, max_age=30, min_hits=1, asso_threshold=0.3, delta_t=3, asso_func="giou", inertia=0.2,
                use_byte=False)                                      # boxmot/configs/ocsort.yaml
    scenarios = {
        # BASELINE config 2 shape, scaled down: occlusion runs exercise freeze / ORU / OCR
        "ocsort_c2": dict(stream=0, n_objects=30, n_frames=150, kw=dict(occlusion=True), params={}),
        "ocsort_churn": dict(stream=902, n_objects=16, n_frames=160, kw=dict(miss_prob=0.3, fp_rate=3.0),
                             params=dict(min_hits=3, max_age=8, det_thresh=0.3)),
        # BYTE stage (ocsort.py:293-317): low-confidence detections rescue unmatched trackers
        "ocsort_byte": dict(stream=905, n_objects=24, n_frames=140, kw=dict(miss_prob=0.1, fp_rate=2.0, occlusion=True),
                            params=dict(use_byte=True, det_thresh=0.5, min_hits=2)),
    }
    only = os.environ.get("GOLDEN_ONLY")
    if only:
        scenarios = {k: v for k, v in scenarios.items() if k in only.split(",")}
    img = np.zeros((1080, 1920, 3), dtype=np.uint8)
    for name, sc in scenarios.items():
        dets, nd, _ = make_stream(2, sc["stream"], sc["n_objects"], sc["n_frames"], **sc["kw"])
        cfg = dict(base)
        cfg.update(sc["params"])
        rh.install()
        rh.reset_counters()
        trk = OCSort(False, **cfg)
        outs, ints, xs, Ps, vels, lasts, pool, heavy_frames = [], [], [], [], [], [], [], []
        for f in range(sc["n_frames"]):
            o = trk.update(dets[f, :nd[f]], img)
            outs.append(o)
            ii, x, P, vel, last = _oc_snapshot(trk)
            ints.append(ii)
            vels.append(vel)
            lasts.append(last)
            xs.append(x)
            if f % 10 == 9 or f == sc["n_frames"] - 1:
                Ps.append(P)
                heavy_frames.append(f)
        out_flat, out_offs = _ragged(outs, 8)
        int_flat, int_offs = _ragged(ints, 6)
        _save(name, dets=dets, ndets=nd, out=out_flat, out_offs=out_offs, rec=int_flat.astype(np.int32),
              rec_offs=int_offs, x=_ragged(xs, 7)[0], vel=_ragged(vels, 2)[0], last=_ragged(lasts, 5)[0],
              P=_ragged(Ps, 49)[0], heavy_frames=np.array(heavy_frames, dtype=np.int32),
              params=np.array([cfg["det_thresh"], cfg["max_age"], cfg["min_hits"], cfg["asso_threshold"], cfg["delta_t"],
                               cfg["inertia"], float(cfg["use_byte"])], dtype=np.float64), img_hw=np.array([1080, 1920]))
    if only:
        return
    # the reference's own known-answer inputs (tests/test_python.py:97-139)
    det = np.array([[144, 212, 578, 480, 0.82, 0], [425, 281, 576, 472, 0.56, 65]], dtype=np.float64)
    rh.reset_counters()
    trk = OCSort(False, **base)
    outs = [trk.update(det, img) for _ in range(3)]
    rh.reset_counters()
    trk = OCSort(False, **dict(base, min_hits=2))
    seq = [np.empty((0, 6)), np.empty((0, 6)), det, np.empty((0, 6)), det, det, det]
    sizes = [trk.update(d, img).size for d in seq]
    _save("ocsort_2box", det=det, out=np.stack(outs), min_hits_sizes=np.array(sizes))


# ----------------------------------------------------------------------------- BoT-SORT
from scenarios import (BOTSORT_SCENARIOS, BOTSORT_YAML, STRONGSORT_SCENARIOS, STRONGSORT_YAML, botsort_inputs,  # noqa: E402
                       camera_warps, strongsort_inputs)


def _bs_snapshot(trk):
    ts = trk.tracked_stracks + trk.lost_stracks
    ints = np.array([[t.id, t.state, int(t.is_activated), t.frame_id, t.start_frame, t.tracklet_len] for t in ts],
                    dtype=np.int32).reshape(-1, 6)
    mean = np.stack([t.mean for t in ts]) if ts else np.zeros((0, 8))
    cov = np.stack([t.covariance for t in ts]) if ts else np.zeros((0, 8, 8))
    aux = np.array([[t.score, t.cls, t.det_ind] for t in ts], dtype=np.float64).reshape(-1, 3)
    feat = [t.smooth_feat for t in ts]
    return len(trk.tracked_stracks), len(trk.lost_stracks), ints, mean, cov, aux, feat


def gen_botsort():
    rh.install()
    from boxmot.trackers.botsort.bot_sort import BoTSORT
    img = np.zeros((1080, 1920, 3), dtype=np.uint8)
    only = os.environ.get("GOLDEN_ONLY")
    for name, sc in BOTSORT_SCENARIOS.items():
        if only and name not in only.split(","):
            continue
        dets, nd, embs, feats = botsort_inputs(sc)
        cfg = dict(BOTSORT_YAML)
        cfg.update(sc["params"])
        rh.reset_counters()
        trk = BoTSORT(None, "cpu", False, **cfg)
        trk.cmc = rh.ScriptedCMC(camera_warps(sc)) if sc.get("camera") else rh.IdentityCMC()
        outs, ints, counts, means, auxs, covs, cov_frames = [], [], [], [], [], [], []
        for f in range(sc["n_frames"]):
            trk.cmc.frame = f
            rows = np.nonzero(dets[f, :nd[f], 4] > cfg["track_high_thresh"])[0]
            if cfg.get("with_reid", True) and len(rows):
                rh.FakeReID.queue.append(embs[f, rows])
            o = trk.update(dets[f, :nd[f]], img)
            outs.append(o)
            nt, nl, ii, m, c, aux, feat = _bs_snapshot(trk)
            counts.append((nt, nl))
            ints.append(ii)
            means.append(m)
            auxs.append(aux)
            if f % 10 == 9 or f == sc["n_frames"] - 1:
                covs.append(c.reshape(-1, 64))
                cov_frames.append(f)
        assert not rh.FakeReID.queue
        final_feat = (np.stack([x for x in feat]).astype(np.float32) if cfg.get("with_reid", True) and feat
                      else np.zeros((0, sc["emb_dim"]), dtype=np.float32))
        out_flat, out_offs = _ragged(outs, 8)
        int_flat, int_offs = _ragged(ints, 6)
        _save(name, ndets=nd, dets_sum=np.array([dets.sum(), float(np.abs(feats).sum())]),
              out=out_flat, out_offs=out_offs, rec=int_flat.astype(np.int32), rec_offs=int_offs,
              counts=np.array(counts, dtype=np.int32), mean=_ragged(means, 8)[0], aux=_ragged(auxs, 3)[0],
              cov=_ragged(covs, 64)[0], cov_frames=np.array(cov_frames, dtype=np.int32), final_feat=final_feat)


# ----------------------------------------------------------------------------- MOT17-mini replay
MOT_SEQS = ["MOT17-02-FRCNN", "MOT17-05-FRCNN", "MOT17-09-FRCNN"]


def gen_mot():
    """Public detections of three MOT17-mini sequences (assets/MOT17-mini/train/*/det/det.txt, dataset data) replayed
    through the reference's ByteTrack and OC-SORT; the fixture holds the detection rows and the integer MOT rows the
    reference's writer (examples/utils.py:8-28) would put in the result files."""
    rh.install()
    from yolo_tracking_b200 import mot_io
    from yolo_tracking_b200.replay import dense_frames
    from boxmot.trackers.ocsort.ocsort import OCSort
    out = {}
    img = np.zeros((1080, 1920, 3), dtype=np.uint8)
    for name in MOT_SEQS:
        base = os.path.join(rh.REF_ROOT, "assets", "MOT17-mini", "train", name)
        raw = np.loadtxt(os.path.join(base, "det", "det.txt"), delimiter=",", ndmin=2)[:, :7]
        info = mot_io.read_seqinfo(base)
        length = min(info["length"], 300)
        raw = raw[raw[:, 0] <= length]
        out[name + "_det"] = raw
        out[name + "_len"] = np.int64(length)
        frames, dets = mot_io.split_det_rows(raw)
        seq = dense_frames(frames, dets, length)
        for kind in ("bytetrack", "ocsort"):
            if kind == "bytetrack":
                trk = rh.make_tracker("bytetrack")
            else:
                rh.reset_counters()
                trk = OCSort(False, det_thresh=0, max_age=30, min_hits=1, asso_threshold=0.3, delta_t=3, asso_func="giou",
                             inertia=0.2, use_byte=False)
            rows = []
            for f, d in enumerate(seq):
                o = trk.update(d, img)
                if o.size:
                    rows.append(mot_io.mot_rows(o, f))
            rows = np.concatenate(rows, axis=0)
            out[f"{name}_{kind}"] = mot_io.as_int_rows(rows)
            print(name, kind, len(rows), "rows")
    _save("mot17_mini", **out)


def gen_mot_deepocsort():
    """The same three MOT17-mini detection streams (rows taken from tests/golden/mot17_mini.npz) through the reference's
    DeepOCSORT with seeded stand-in embeddings (scenarios.mot_feats): integer MOT rows per sequence."""
    rh.install()
    from scenarios import DEEPOCSORT_YAML, mot_feats
    from yolo_tracking_b200 import mot_io
    from yolo_tracking_b200.replay import dense_frames
    from boxmot.trackers.deepocsort.deep_ocsort import DeepOCSort
    g = np.load(os.path.join(HERE, "mot17_mini.npz"))
    img = np.zeros((1080, 1920, 3), dtype=np.uint8)
    out = {}
    for si, name in enumerate(MOT_SEQS):
        frames, dets = mot_io.split_det_rows(g[name + "_det"])
        seq = dense_frames(frames, dets, int(g[name + "_len"]))
        trk = DeepOCSort(None, "cpu", False, False, **DEEPOCSORT_YAML)
        trk.cmc = rh.IdentityCMC()
        rows = []
        for f, d in enumerate(seq):
            keep = d[:, 4] > DEEPOCSORT_YAML["det_thresh"]
            if keep.any():
                rh.FakeReID.queue.append(mot_feats(si, f, int(keep.sum())))
            o = trk.update(d, img)
            if o.size:
                rows.append(mot_io.mot_rows(o, f))
        assert not rh.FakeReID.queue
        out[name] = mot_io.as_int_rows(np.concatenate(rows, axis=0))
        print(name, "deepocsort", len(out[name]), "rows")
    _save("mot17_mini_deepocsort", **out)


def gen_mot_botsort():
    """The three MOT17-mini detection streams through the reference's BoT-SORT (botsort.yaml, identity camera, seeded
    stand-in embeddings for the detections above track_high_thresh): integer MOT rows per sequence."""
    rh.install()
    from scenarios import mot_feats
    from yolo_tracking_b200 import mot_io
    from yolo_tracking_b200.replay import dense_frames
    from boxmot.trackers.botsort.bot_sort import BoTSORT
    g = np.load(os.path.join(HERE, "mot17_mini.npz"))
    img = np.zeros((1080, 1920, 3), dtype=np.uint8)
    out = {}
    for si, name in enumerate(MOT_SEQS):
        frames, dets = mot_io.split_det_rows(g[name + "_det"])
        seq = dense_frames(frames, dets, int(g[name + "_len"]))
        rh.reset_counters()
        trk = BoTSORT(None, "cpu", False, **BOTSORT_YAML)
        trk.cmc = rh.IdentityCMC()
        rows = []
        for f, d in enumerate(seq):
            hi = np.nonzero(d[:, 4] > BOTSORT_YAML["track_high_thresh"])[0]
            if len(hi):
                rh.FakeReID.queue.append(mot_feats(si, f, len(d))[hi])
            o = trk.update(d, img)
            if o.size:
                rows.append(mot_io.mot_rows(o, f))
        assert not rh.FakeReID.queue
        out[name] = mot_io.as_int_rows(np.concatenate(rows, axis=0))
        print(name, "botsort", len(out[name]), "rows")
    _save("mot17_mini_botsort", **out)


def gen_mot_strongsort():
    """The three MOT17-mini detection streams through the reference's StrongSORT (identity camera, seeded stand-in
    embeddings for every detection row): integer MOT rows per sequence."""
    rh.install()
    from scenarios import mot_feats
    from yolo_tracking_b200 import mot_io
    from yolo_tracking_b200.replay import dense_frames
    from boxmot.trackers.strongsort.strong_sort import StrongSORT
    g = np.load(os.path.join(HERE, "mot17_mini.npz"))
    img = np.zeros((1080, 1920, 3), dtype=np.uint8)
    out = {}
    for si, name in enumerate(MOT_SEQS):
        frames, dets = mot_io.split_det_rows(g[name + "_det"])
        seq = dense_frames(frames, dets, int(g[name + "_len"]))
        trk = StrongSORT(None, "cpu", False, **STRONGSORT_YAML)
        trk.cmc = rh.IdentityCMC()
        rows = []
        for f, d in enumerate(seq):
            if len(d):
                rh.FakeReID.queue.append(mot_feats(si, f, len(d)))
            o = trk.update(d, img)
            if o.size:
                rows.append(mot_io.mot_rows(o, f))
        assert not rh.FakeReID.queue
        out[name] = mot_io.as_int_rows(np.concatenate(rows, axis=0))
        print(name, "strongsort", len(out[name]), "rows")
    _save("mot17_mini_strongsort", **out)


# ----------------------------------------------------------------------------- StrongSORT
def gen_strongsort():
    rh.install()
    from boxmot.trackers.strongsort.strong_sort import StrongSORT
    img = np.zeros((1080, 1920, 3), dtype=np.uint8)
    only = os.environ.get("GOLDEN_ONLY")
    for name, sc in STRONGSORT_SCENARIOS.items():
        if only and name not in only.split(","):
            continue
        dets, nd, embs, feats = strongsort_inputs(sc)
        cfg = dict(STRONGSORT_YAML)
        cfg.update(sc["params"])
        trk = StrongSORT(None, "cpu", False, **cfg)
        trk.cmc = rh.ScriptedCMC(camera_warps(sc)) if sc.get("camera") else rh.IdentityCMC()
        outs, recs, means, covs, cov_frames = [], [], [], [], []
        for f in range(sc["n_frames"]):
            trk.cmc.frame = f
            if nd[f]:
                rh.FakeReID.queue.append(embs[f, :nd[f]])
            o = trk.update(dets[f, :nd[f]], img)
            outs.append(o)
            ts = trk.tracker.tracks
            recs.append(np.array([[t.id, t.state, t.hits, t.age, t.time_since_update, len(trk.tracker.metric.samples.get(t.id, []))]
                                  for t in ts], dtype=np.int32).reshape(-1, 6))
            means.append(np.stack([t.mean for t in ts]) if ts else np.zeros((0, 8)))
            if f % 10 == 9 or f == sc["n_frames"] - 1:
                covs.append(np.stack([t.covariance.reshape(64) for t in ts]) if ts else np.zeros((0, 64)))
                cov_frames.append(f)
        assert not rh.FakeReID.queue
        ts = trk.tracker.tracks
        final_feat = np.stack([t.features[-1] for t in ts]).astype(np.float32) if ts else np.zeros((0, sc["emb_dim"]), dtype=np.float32)
        out_flat, out_offs = _ragged(outs, 8)
        rec_flat, rec_offs = _ragged(recs, 6)
        _save(name, ndets=nd, dets_sum=np.array([dets.sum(), float(np.abs(feats).sum())]), out=out_flat, out_offs=out_offs,
              rec=rec_flat.astype(np.int32), rec_offs=rec_offs, mean=_ragged(means, 8)[0], cov=_ragged(covs, 64)[0],
              cov_frames=np.array(cov_frames, dtype=np.int32), final_feat=final_feat)


def _doc_snapshot(trk):
    ts = trk.trackers
    ints = np.array([[t.id, t.age, t.time_since_update, t.hits, t.hit_streak, int(t.kf.observed), int(t.frozen)] for t in ts],
                    dtype=np.int32).reshape(-1, 7)
    x = np.stack([t.kf.x[:, 0] for t in ts]) if ts else np.zeros((0, 8))
    P = np.stack([t.kf.P.reshape(64) for t in ts]) if ts else np.zeros((0, 64))
    vel = np.array([t.velocity if t.velocity is not None else np.zeros(2) for t in ts]).reshape(-1, 2)
    last = np.array([t.last_observation for t in ts], dtype=np.float64).reshape(-1, 5)
    return ints, x, P, vel, last


def gen_deepocsort():
    rh.install()
    from scenarios import DEEPOCSORT_SCENARIOS, DEEPOCSORT_YAML, deepocsort_inputs
    from boxmot.trackers.deepocsort.deep_ocsort import DeepOCSort
    img = np.zeros((1080, 1920, 3), dtype=np.uint8)
    only = os.environ.get("GOLDEN_ONLY")
    for name, sc in DEEPOCSORT_SCENARIOS.items():
        if only and name not in only.split(","):
            continue
        cfg = dict(DEEPOCSORT_YAML)
        cfg.update(sc["params"])
        dets, nd, embs, feats = deepocsort_inputs(sc, cfg["det_thresh"])
        trk = DeepOCSort(None, "cpu", False, False, **cfg)          # resets KalmanBoxTracker.count to 1 (deep_ocsort.py:347)
        trk.cmc = rh.ScriptedCMC(camera_warps(sc)) if sc.get("camera") else rh.IdentityCMC()
        outs, ints, xs, Ps, vels, lasts, heavy = [], [], [], [], [], [], []
        for f in range(sc["n_frames"]):
            trk.cmc.frame = f
            keep = dets[f, :nd[f], 4] > cfg["det_thresh"]
            if keep.any() and not cfg.get("embedding_off"):
                rh.FakeReID.queue.append(embs[f, :nd[f]][keep])
            o = trk.update(dets[f, :nd[f]], img)
            outs.append(o)
            ii, x, P, vel, last = _doc_snapshot(trk)
            ints.append(ii)
            xs.append(x)
            vels.append(vel)
            lasts.append(last)
            if f % 10 == 9 or f == sc["n_frames"] - 1:
                Ps.append(P)
                heavy.append(f)
        assert not rh.FakeReID.queue
        ts = trk.trackers
        final_emb = np.stack([np.asarray(t.emb, dtype=np.float64).reshape(-1) for t in ts]) if ts else np.zeros((0, sc["emb_dim"]))
        out_flat, out_offs = _ragged(outs, 8)
        int_flat, int_offs = _ragged(ints, 7)
        _save(name, ndets=nd, dets_sum=np.array([dets.sum(), float(sum(np.abs(f).sum() for f in feats))]), out=out_flat,
              out_offs=out_offs, rec=int_flat.astype(np.int32), rec_offs=int_offs, x=_ragged(xs, 8)[0], vel=_ragged(vels, 2)[0],
              last=_ragged(lasts, 5)[0], P=_ragged(Ps, 64)[0], heavy_frames=np.array(heavy, dtype=np.int32), final_emb=final_emb)


def _hyb_snapshot(trk):
    ts = trk.trackers
    ints = np.array([[t.id, t.age, t.time_since_update, t.hits, t.hit_streak, int(t.kf.observed)] for t in ts],
                    dtype=np.int32).reshape(-1, 6)
    x = np.stack([t.kf.x[:, 0] for t in ts]) if ts else np.zeros((0, 9))
    P = np.stack([t.kf.P.reshape(81) for t in ts]) if ts else np.zeros((0, 81))
    vel = np.array([[v if v is not None else np.zeros(2) for v in (t.velocity_lt, t.velocity_rt, t.velocity_lb, t.velocity_rb)]
                    for t in ts], dtype=np.float64).reshape(-1, 8)
    last = np.array([t.last_observation for t in ts], dtype=np.float64).reshape(-1, 5)
    return ints, x, P, vel, last


def gen_hybridsort():
    rh.install()
    from scenarios import HYBRIDSORT_SCENARIOS, HYBRIDSORT_YAML, hybridsort_inputs
    from boxmot.trackers.hybridsort.hybridsort import HybridSORT
    img = np.zeros((1080, 1920, 3), dtype=np.uint8)
    only = os.environ.get("GOLDEN_ONLY")
    for name, sc in HYBRIDSORT_SCENARIOS.items():
        if only and name not in only.split(","):
            continue
        cfg = dict(HYBRIDSORT_YAML)
        cfg.update(sc["params"])
        dets, nd, embs, feats = hybridsort_inputs(sc, cfg["det_thresh"])
        trk = HybridSORT(None, "cpu", False, **cfg)                 # resets KalmanBoxTracker.count (hybridsort.py:364)
        outs, ints, xs, Ps, vels, lasts, heavy = [], [], [], [], [], [], []
        for f in range(sc["n_frames"]):
            d = dets[f, :nd[f]]
            if nd[f]:
                # features of every detection of a call (hybridsort.py:394); the PerClassDecorator makes one call per class, in
                # the iteration order of these very expressions (boxmot/utils/__init__.py:33-45)
                by_cls = {class_id: np.array([i for i, det in enumerate(d) if det[5] == class_id]) for class_id in set(det[5] for det in d)}
                relevant = set([t.cls for t in trk.trackers]).union(set(by_cls.keys()))
                for class_id in relevant:
                    idx = by_cls.get(int(class_id))
                    if idx is not None and len(idx):
                        rh.FakeReID.queue.append(embs[f, :nd[f]][idx])
            o = trk.update(d, img)                                   # through PerClassDecorator
            outs.append(o)
            ii, x, P, vel, last = _hyb_snapshot(trk)
            ints.append(ii)
            xs.append(x)
            vels.append(vel)
            lasts.append(last)
            if f % 10 == 9 or f == sc["n_frames"] - 1:
                Ps.append(P)
                heavy.append(f)
        assert not rh.FakeReID.queue
        ts = trk.trackers
        final_emb = np.stack([np.asarray(t.smooth_feat, dtype=np.float32).reshape(-1) for t in ts]) if ts else np.zeros((0, sc["emb_dim"]), dtype=np.float32)
        out_flat, out_offs = _ragged(outs, 8)
        int_flat, int_offs = _ragged(ints, 6)
        _save(name, ndets=nd, dets_sum=np.array([dets.sum(), float(sum(np.abs(f).sum() for f in feats))]), out=out_flat,
              out_offs=out_offs, rec=int_flat.astype(np.int32), rec_offs=int_offs, x=_ragged(xs, 9)[0], vel=_ragged(vels, 8)[0],
              last=_ragged(lasts, 5)[0], P=_ragged(Ps, 81)[0], heavy_frames=np.array(heavy, dtype=np.int32), final_emb=final_emb)


def gen_fullsize():
    """Rows-only goldens at the BASELINE config sizes (scenarios.FULLSIZE)."""
    rh.install()
    from scenarios import (BOTSORT_YAML, DEEPOCSORT_YAML, FULLSIZE, HYBRIDSORT_YAML, STRONGSORT_YAML, fullsize_inputs)
    img = np.zeros((2160, 3840, 3), dtype=np.uint8)
    only = os.environ.get("GOLDEN_ONLY")
    for name, sc in FULLSIZE.items():
        if only and name not in only.split(","):
            continue
        dets, nd, embs = fullsize_inputs(sc)
        kind = sc["kind"]
        rh.reset_counters()
        if kind == "bytetrack":
            trk = rh.make_tracker("bytetrack")
        elif kind == "ocsort":
            from boxmot.trackers.ocsort.ocsort import OCSort
            trk = OCSort(det_thresh=0, max_age=30, min_hits=1, asso_threshold=0.3, delta_t=3, asso_func="giou", inertia=0.2, use_byte=False)
        elif kind == "botsort":
            from boxmot.trackers.botsort.bot_sort import BoTSORT
            trk = BoTSORT(None, "cpu", False, **BOTSORT_YAML)
            trk.cmc = rh.IdentityCMC()
        elif kind == "deepocsort":
            from boxmot.trackers.deepocsort.deep_ocsort import DeepOCSort
            trk = DeepOCSort(None, "cpu", False, False, **DEEPOCSORT_YAML)
            trk.cmc = rh.IdentityCMC()
        elif kind == "hybridsort":
            from boxmot.trackers.hybridsort.hybridsort import HybridSORT
            trk = HybridSORT(None, "cpu", False, **HYBRIDSORT_YAML)
        else:
            from boxmot.trackers.strongsort.strong_sort import StrongSORT
            trk = StrongSORT(None, "cpu", False, **STRONGSORT_YAML)
            trk.cmc = rh.IdentityCMC()
        ids, dis, offs, boxes, box_frames, lasts = [], [], [0], [], [], []
        for f in range(sc["n_frames"]):
            d = dets[f, :nd[f]]
            if kind == "botsort":
                rows = np.nonzero(d[:, 4] > BOTSORT_YAML["track_high_thresh"])[0]
                if len(rows):
                    rh.FakeReID.queue.append(embs[f, rows])
            elif kind == "deepocsort":
                keep = d[:, 4] > DEEPOCSORT_YAML["det_thresh"]
                if keep.any():
                    rh.FakeReID.queue.append(embs[f, :nd[f]][keep])
            elif kind in ("strongsort", "hybridsort") and nd[f]:
                rh.FakeReID.queue.append(embs[f, :nd[f]])              # one class: one get_features call on every box
            o = np.asarray(trk.update(d, None if kind == "bytetrack" else img), dtype=np.float64).reshape(-1, 8)
            ids.append(o[:, 4].astype(np.int32))
            lasts.append(o[:, 7].copy())
            dis.append(o[:, 7].astype(np.int32))
            offs.append(offs[-1] + len(o))
            if f % sc["box_every"] == sc["box_every"] - 1 or f == sc["n_frames"] - 1:
                boxes.append(o[:, :4])
                box_frames.append(f)
        assert not rh.FakeReID.queue
        _save(name, ndets=nd, dets_sum=np.array([dets.sum(), 0.0 if embs is None else float(np.abs(embs.astype(np.float64)).sum())]),
              ids=np.concatenate(ids), det_ind=np.concatenate(dis).astype(np.int16), offs=np.array(offs, dtype=np.int64),
              boxes=np.concatenate(boxes), box_frames=np.array(box_frames, dtype=np.int32),
              **({"last_col": np.concatenate(lasts)} if kind == "hybridsort" else {}))      # HybridSORT's last column is a score


def gen_mot_hybridsort():
    """The three MOT17-mini detection streams through the reference's HybridSORT (hybridsort.yaml as the factory forwards it)
    with seeded stand-in embeddings for every detection: integer MOT rows per sequence."""
    rh.install()
    from scenarios import HYBRIDSORT_YAML, mot_feats
    from yolo_tracking_b200 import mot_io
    from yolo_tracking_b200.replay import dense_frames
    from boxmot.trackers.hybridsort.hybridsort import HybridSORT
    g = np.load(os.path.join(HERE, "mot17_mini.npz"))
    img = np.zeros((1080, 1920, 3), dtype=np.uint8)
    out = {}
    for si, name in enumerate(MOT_SEQS):
        frames, dets = mot_io.split_det_rows(g[name + "_det"])
        seq = dense_frames(frames, dets, int(g[name + "_len"]))
        trk = HybridSORT(None, "cpu", False, **HYBRIDSORT_YAML)
        rows = []
        for f, d in enumerate(seq):
            assert len(set(d[:, 5])) <= 1                           # one class: the per-class wrapper makes one call
            if len(d):
                rh.FakeReID.queue.append(mot_feats(si, f, len(d)))
            o = trk.update(d, img)
            if o.size:
                rows.append(mot_io.mot_rows(o, f))
        assert not rh.FakeReID.queue
        out[name] = mot_io.as_int_rows(np.concatenate(rows, axis=0))
        print(name, "hybridsort", len(out[name]), "rows")
    _save("mot17_mini_hybridsort", **out)


GENERATORS = {"fullsize": gen_fullsize, "hybridsort": gen_hybridsort, "mot_hybridsort": gen_mot_hybridsort, "deepocsort": gen_deepocsort, "mot_deepocsort": gen_mot_deepocsort, "mot_strongsort": gen_mot_strongsort, "mot_botsort": gen_mot_botsort, "kf": gen_kf, "costs": gen_costs, "bytetrack": gen_bytetrack, "ocsort": gen_ocsort, "botsort": gen_botsort,
              "mot": gen_mot, "strongsort": gen_strongsort, "aux": gen_aux}

if __name__ == "__main__":
    names = sys.argv[1:] or list(GENERATORS)
    for n in names:
        GENERATORS[n]()
