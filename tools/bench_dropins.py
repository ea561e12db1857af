"""Single-stream latency of the reference-shaped drop-ins on one B200: `tracker.update(dets, img)` p50 / p99 per frame for
BYTETracker, OCSort, BoTSORT at BASELINE config 1 (1 stream, ~50 detections per frame) and for DeepOCSORT / StrongSORT / HybridSORT at
100 objects with 512-d embeddings, next to the oracle port of the reference on one host core (same synthetic stream, seam
features passed in, identity camera).  One call = pack -> one H2D copy -> one fused step -> one D2H copy -> rebuild the
reference's [M, 8] rows (StrongSORT: the eight-launch batched step on one stream).  This is a latency figure; the multi-stream
throughput is bench.py.  Every line carries the SM clocks sampled while it ran.
usage: python tools/bench_dropins.py [--frames 300] > profiles/rNN_dropins.jsonl"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import ClockSampler  # noqa: E402
from yolo_tracking_b200.synth import make_stream  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=300)
ap.add_argument("--dim", type=int, default=512)
ap.add_argument("--no-cpu", action="store_true")
args = ap.parse_args()


def run(step, frames, warm=20):
    ms = []
    for f in range(frames):
        t0 = time.perf_counter()
        step(f)
        if f >= warm:
            ms.append((time.perf_counter() - t0) * 1e3)
    return float(np.percentile(ms, 50)), float(np.percentile(ms, 99)), float(np.mean(ms))


def main():
    import yolo_tracking_b200 as pkg
    from oracle.botsort import BoTSORTOracle
    from oracle.bytetrack import ByteTrackOracle
    from oracle.deepocsort import DeepOCSortOracle
    from oracle.hybridsort import HybridSortOracle
    from oracle.ocsort import OCSortOracle
    from oracle.strongsort import StrongSORTOracle
    F = args.frames
    bt = dict(track_thresh=0.5, match_thresh=0.8, track_buffer=30, frame_rate=30)
    oc = dict(det_thresh=0, max_age=30, min_hits=1, asso_threshold=0.3, delta_t=3, asso_func="giou", inertia=0.2)
    bs = dict(track_high_thresh=0.33824964456239337, track_low_thresh=0.1, new_track_thresh=0.21144301345190655, track_buffer=60,
              match_thresh=0.22734550911325851, proximity_thresh=0.5945380911899254, appearance_thresh=0.4818211117541298, frame_rate=30)
    do = dict(det_thresh=0, max_age=30, min_hits=1, iou_threshold=0.3, delta_t=3, asso_func="giou", inertia=0.2)
    ss = dict(max_dist=0.2, max_iou_dist=0.7, max_age=30, n_init=1, nn_budget=100, mc_lambda=0.995, ema_alpha=0.8)
    rgb = np.zeros((1080, 1920, 3), dtype=np.uint8)
    cases = [
        ("bytetrack", 53, 0, lambda: pkg.BYTETracker(**bt), lambda: ByteTrackOracle(**bt)),
        ("ocsort", 53, 0, lambda: pkg.OCSORT(False, **oc), lambda: OCSortOracle(False, use_byte=False, **oc)),
        ("botsort", 53, args.dim, lambda: pkg.BoTSORT(None, 0, False, feat_dim=args.dim, **bs), lambda: BoTSORTOracle(**bs)),
        ("deepocsort", 100, args.dim, lambda: pkg.DeepOCSORT(None, 0, False, False, **do), lambda: DeepOCSortOracle(**do)),
        ("strongsort", 100, args.dim, lambda: pkg.StrongSORT(None, 0, False, **ss), lambda: StrongSORTOracle(**ss)),
        ("hybridsort", 100, args.dim, lambda: pkg.HybridSORT(None, 0, False, **do), lambda: HybridSortOracle(**do)),
    ]
    for name, objects, dim, mk, mk_orc in cases:
        frames = F if name != "strongsort" else min(F, 120)
        dets, nd, embs = make_stream(1 if objects <= 64 else 4, 1, objects, frames, emb_dim=dim, occlusion=(name in ("ocsort", "deepocsort", "hybridsort")))
        dets = dets.astype(np.float32).astype(np.float64)

        def seam(f, kind):
            d, e = dets[f, :nd[f]], embs[f, :nd[f]]
            if kind == "botsort":
                rows = np.nonzero(d[:, 4] > bs["track_high_thresh"])[0]
                out = np.zeros_like(e)
                if len(rows):
                    out[rows] = e[rows] / np.linalg.norm(e[rows])
                return out
            raw = e[d[:, 4] > 0] if kind == "deepocsort" else e
            return raw / np.linalg.norm(raw) if len(raw) else raw
        feats = [seam(f, name) for f in range(frames)] if dim else None

        def gpu_step(trk):
            if name == "bytetrack":
                return lambda f: trk.update(dets[f, :nd[f]], None)
            if name == "ocsort":
                return lambda f: trk.update(dets[f, :nd[f]], rgb)
            return lambda f: trk.update(dets[f, :nd[f]], rgb, feats=feats[f])

        def cpu_step(orc):
            if name == "bytetrack":
                return lambda f: orc.update(dets[f, :nd[f]], None)
            if name == "ocsort":
                return lambda f: orc.update(dets[f, :nd[f]], (1080, 1920))
            if name == "deepocsort":
                return lambda f: orc.update(dets[f, :nd[f]], feats[f], (1080, 1920))
            if name == "hybridsort":
                return lambda f: orc.update(dets[f, :nd[f]], feats[f][dets[f, :nd[f], 4] > 0], (1080, 1920))
            return lambda f: orc.update(dets[f, :nd[f]], feats[f])
        trk = mk()
        sampler = ClockSampler(0)
        sampler.start()
        p50, p99, mean = run(gpu_step(trk), frames)
        sampler.stop_flag = True
        sampler.join()
        line = {"tracker": name, "objects": objects, "dets_per_frame": float(nd.mean()), "frames": frames, "emb_dim": dim,
                "update_ms_p50": p50, "update_ms_p99": p99, "update_ms_mean": mean, "frames_per_s": 1e3 / mean,
                "path": "one-stream context of the batched StrongSORT step (eight launches per frame, device buffers via torch)" if name == "strongsort" else
                        ("one-stream context of the fused frame step through the padded host interface (plus one state read per frame: "
                         "the per-class wrapper needs the classes of the live trackers)" if name == "hybridsort" else
                         "one-stream context of the fused frame step through the packed host interface"),
                "clocks": sampler.summary()}
        if not args.no_cpu:
            c50, c99, cmean = run(cpu_step(mk_orc()), frames)
            line.update(cpu_port_ms_p50=c50, cpu_port_ms_mean=cmean, cpu_cores=1)
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
