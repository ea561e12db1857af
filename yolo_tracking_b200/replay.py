"""Replay MOTChallenge public detections through the batched CUDA trackers and write MOT result files
(the reference does this one sequence per subprocess, examples/val.py:189-259).

  python -m yolo_tracking_b200.replay --tracker bytetrack --source <dir with <seq>/det/det.txt> --out runs/mot

Every sequence is one stream of a single `BatchedTracker`; shorter sequences simply stop receiving detections.
"""
from __future__ import annotations

import argparse
import os

import numpy as np
import yaml

from . import mot_io
from .tracker_zoo import get_tracker_config


def tracker_params(kind, config_path=None):
    cfg = yaml.safe_load(open(config_path or get_tracker_config(kind)))
    if kind == "bytetrack":
        return dict(track_thresh=cfg["track_thresh"], match_thresh=cfg["match_thresh"], track_buffer=cfg["track_buffer"],
                    frame_rate=cfg["frame_rate"])
    if kind == "ocsort":
        return dict(det_thresh=cfg["det_thresh"], max_age=cfg["max_age"], min_hits=cfg["min_hits"], asso_threshold=cfg["iou_thresh"],
                    delta_t=cfg["delta_t"], asso_func=cfg["asso_func"], inertia=cfg["inertia"], use_byte=cfg["use_byte"])
    if kind == "botsort":
        return dict(track_high_thresh=cfg["track_high_thresh"], track_low_thresh=cfg["track_low_thresh"],
                    new_track_thresh=cfg["new_track_thresh"], track_buffer=cfg["track_buffer"], match_thresh=cfg["match_thresh"],
                    proximity_thresh=cfg["proximity_thresh"], appearance_thresh=cfg["appearance_thresh"],
                    frame_rate=cfg["frame_rate"], with_reid=False)
    raise ValueError(f"No such tracker: {kind}")


def replay(kind, sequences, params=None, img_hw=(1080, 1920), device=0):
    """sequences: list of per-frame detection lists (frame f of sequence s = sequences[s][f], dets[N, 6]).
    Returns per sequence the stacked MOT rows [n, 9] (float; `mot_io.as_int_rows` gives what the file holds)."""
    from .batch import BatchedTracker
    S = len(sequences)
    n_frames = max(len(seq) for seq in sequences)
    max_nd = max((len(d) for seq in sequences for d in seq), default=0)
    cap = max(64, (max_nd + 31) // 32 * 32)
    tcap = min(512, max(64, 2 * cap))
    trk = BatchedTracker(kind, S, max_tracks=tcap, max_dets=cap, device=device, **(params or tracker_params(kind)))
    dets = np.zeros((S, cap, 6))
    nd = np.zeros((S,), dtype=np.int32)
    out_rows = [[] for _ in range(S)]
    for f in range(n_frames):
        for s, seq in enumerate(sequences):
            d = seq[f] if f < len(seq) else np.empty((0, 6))
            dets[s, :len(d)] = d
            nd[s] = len(d)
        out, nout = trk.update_batch(dets, nd, img_hw=img_hw)
        for s, seq in enumerate(sequences):
            if f < len(seq) and nout[s]:
                out_rows[s].append(mot_io.mot_rows(out[s, :nout[s]], f))
    trk.sync()
    trk.close()
    return [np.concatenate(r, axis=0) if r else np.zeros((0, 9)) for r in out_rows]


def dense_frames(frames, dets, length):
    """Detections per frame 1..length (frames without a row get an empty array)."""
    by = {int(f): d for f, d in zip(frames, dets)}
    return [by.get(f, np.empty((0, 6))) for f in range(1, length + 1)]


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__)
    ap.add_argument("--tracker", default="bytetrack", choices=["bytetrack", "ocsort", "botsort"])
    ap.add_argument("--source", required=True, help="directory holding <sequence>/det/det.txt (+ seqinfo.ini)")
    ap.add_argument("--out", default="runs/mot")
    ap.add_argument("--config", default=None, help="tracker YAML (default: the packaged boxmot config)")
    ap.add_argument("--device", type=int, default=0)
    args = ap.parse_args(argv)
    names, seqs, hw = [], [], (1080, 1920)
    for name in sorted(os.listdir(args.source)):
        det = os.path.join(args.source, name, "det", "det.txt")
        if not os.path.exists(det):
            continue
        frames, dets = mot_io.read_det_txt(det)
        length = int(frames.max()) if len(frames) else 0
        if os.path.exists(os.path.join(args.source, name, "seqinfo.ini")):
            info = mot_io.read_seqinfo(os.path.join(args.source, name))
            length, hw = info["length"], (info["height"], info["width"])
        names.append(name)
        seqs.append(dense_frames(frames, dets, length))
    if not seqs:
        raise SystemExit(f"no <sequence>/det/det.txt under {args.source}")
    results = replay(args.tracker, seqs, tracker_params(args.tracker, args.config), img_hw=hw, device=args.device)
    os.makedirs(args.out, exist_ok=True)
    for name, rows in zip(names, results):
        np.savetxt(os.path.join(args.out, name + ".txt"), rows, fmt="%d")
        print(f"{name}: {len(rows)} rows -> {os.path.join(args.out, name + '.txt')}")


if __name__ == "__main__":
    main()
