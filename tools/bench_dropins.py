"""Frame rate of the operator-backed single-stream drop-ins (StrongSORT, DeepOCSORT) on one B200 next to the oracle port
of the reference on one host core, same synthetic stream (seam features passed in, no ReID network, identity camera).
These two trackers are NOT fused frame steps: every frame is ~10 operator launches with host list logic in between, so
this is a latency figure, not the multi-stream throughput of bench.py.
usage: python tools/bench_dropins.py [--objects 100] [--frames 120]"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolo_tracking_b200.synth import make_stream  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--objects", type=int, default=100)
ap.add_argument("--frames", type=int, default=120)
ap.add_argument("--dim", type=int, default=512)
ap.add_argument("--no-cpu", action="store_true")
args = ap.parse_args()
dets, nd, embs = make_stream(4, 1, args.objects, args.frames, emb_dim=args.dim, occlusion=True)
feats = []
for f in range(args.frames):
    raw = embs[f, :nd[f]].astype(np.float32)
    feats.append(raw / np.linalg.norm(raw) if len(raw) else raw)
img = (2160, 3840) if args.objects > 64 else (1080, 1920)


def run(step, warm=10):
    t0 = None
    for f in range(args.frames):
        if f == warm:
            t0 = time.perf_counter()
        step(f)
    return (time.perf_counter() - t0) / (args.frames - warm) * 1e3


def main():
    from yolo_tracking_b200 import DeepOCSORT, StrongSORT
    do_cfg = dict(det_thresh=0, max_age=30, min_hits=1, iou_threshold=0.3, delta_t=3, asso_func="giou", inertia=0.2)
    ss_cfg = dict(max_dist=0.2, max_iou_dist=0.7, max_age=30, n_init=1, nn_budget=100, mc_lambda=0.995, ema_alpha=0.8)
    rgb = np.zeros((img[0], img[1], 3), dtype=np.uint8)
    for name, mk, mk_orc in (
            ("deepocsort", lambda: DeepOCSORT(None, 0, False, False, **do_cfg), "oracle.deepocsort:DeepOCSortOracle"),
            ("strongsort", lambda: StrongSORT(None, 0, False, **ss_cfg), "oracle.strongsort:StrongSORTOracle")):
        trk = mk()
        gpu_ms = run(lambda f: trk.update(dets[f, :nd[f]], rgb, feats=feats[f]))
        line = {"tracker": name, "objects": args.objects, "frames": args.frames, "emb_dim": args.dim, "gpu_ms_per_frame": gpu_ms,
                "gpu_frames_per_s": 1e3 / gpu_ms, "note": "operator-backed drop-in, one stream, host list logic + ~10 operator launches per frame"}
        if not args.no_cpu:
            mod, cls = mk_orc.split(":")
            orc = getattr(__import__(mod, fromlist=[cls]), cls)(**(do_cfg if name == "deepocsort" else ss_cfg))
            if name == "deepocsort":
                cpu_ms = run(lambda f: orc.update(dets[f, :nd[f]], feats[f], img))
            else:
                cpu_ms = run(lambda f: orc.update(dets[f, :nd[f]], feats[f]))
            line.update(cpu_port_ms_per_frame=cpu_ms, cpu_cores=1)
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
