"""Operator-kernel benchmark on one B200: the Kalman / cost kernels against the HBM roofline and BASELINE
config 4 (1024 streams x 200 tracks x 200 detections: Mahalanobis gating_distance + cosine cost on 512-d
embeddings, separately and back to back).  One JSON line per kernel; CUDA events on the launch stream, inputs
larger than L2 or an L2 flush between iterations (stated per line).

usage: python tools/bench_ops.py [--streams 1024] [--iters 20] > profiles/rNN_ops.jsonl
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from yolo_tracking_b200 import _lib  # noqa: E402
from bench import ClockSampler  # noqa: E402  (NVML SM clock / throttle-reason sampler of the headline bench)

ap = argparse.ArgumentParser()
ap.add_argument("--streams", type=int, default=1024)
ap.add_argument("--tracks", type=int, default=200)
ap.add_argument("--dets", type=int, default=200)
ap.add_argument("--dim", type=int, default=512)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--kf-tracks", type=int, default=819200, help="tracks of the Kalman-operator lines (state 4x the L2)")
ap.add_argument("--only", default="")
ap.add_argument("--gallery-streams", type=int, default=256, help="streams of the StrongSORT gallery line (21 MB of fp32 gallery per stream at 200 x 100 x 512)")
ap.add_argument("--budget", type=int, default=100)
args = ap.parse_args()
S, T, D, F = args.streams, args.tracks, args.dets, args.dim
lib = _lib.load()
dev = torch.device("cuda", 0)
peaks = {}
try:
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
except Exception:
    pass
HBM = float(peaks.get("hbm_gbs", 6650.0))
TF = float(peaks.get("bf16_tflops", 1590.0))
src = "MEASURED_PEAKS.json" if peaks else "fallback"
flush = torch.zeros(64 << 20, dtype=torch.int32, device=dev)          # 256 MB
flush_sink = torch.zeros((), dtype=torch.int64, device=dev)


def p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


LAST_CLOCKS = {}


def timeit(fn, flush_l2):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    sampler = ClockSampler(0)                     # every line carries the clocks sampled while IT was timed
    sampler.start()
    ms = []
    for _ in range(args.iters):
        if flush_l2:
            flush_sink.copy_(flush[:1 << 20].sum())      # READ 256 MB: fills L2 with clean lines (a write would leave
            flush.sum()                                  # dirty lines whose write-back lands in the timed kernel)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    sampler.stop_flag = True
    sampler.join()
    LAST_CLOCKS.clear()
    LAST_CLOCKS.update(sampler.summary())
    return float(np.median(ms)), float(np.min(ms))


def report(name, ms, best, alg_bytes=None, flops=None, extra=None, l2=""):
    line = {"kernel": name, "ms_median": ms, "ms_best": best, "l2": l2}
    if alg_bytes is not None:
        ach = alg_bytes / (ms * 1e-3) / 1e9
        line["roofline"] = {"bound": "hbm", "achieved": ach, "peak": HBM, "unit": "GB/s", "frac": ach / HBM,
                            "alg_bytes_per_launch": alg_bytes, "peak_source": src}
    if flops is not None:
        ach = flops / (ms * 1e-3) / 1e12
        line["tensor"] = {"achieved": ach, "peak": TF, "unit": "TFLOP/s", "frac": ach / TF, "flops_per_launch": flops, "peak_source": src}
    if extra:
        line.update(extra)
    line["clocks"] = dict(LAST_CLOCKS)
    print(json.dumps(line), flush=True)


rng = np.random.default_rng(4)
N = max(args.kf_tracks, S * T)                        # tracks of the Kalman-operator lines
# Kalman state of N tracks (dense reference layout: mean[N, 8], cov[N, 8, 8]): initiate -> a few predict / update rounds
z0 = torch.from_numpy(np.stack([rng.uniform(100, 3700, N), rng.uniform(100, 2000, N), rng.uniform(0.3, 0.8, N),
                                rng.uniform(60, 220, N)], axis=1)).to(dev)
mean = torch.empty((N, 8), dtype=torch.float64, device=dev)
cov = torch.empty((N, 8, 8), dtype=torch.float64, device=dev)
KIND = _lib.KF_XYAH
_lib.check(lib.b200track_kf_initiate(KIND, N, p(z0), p(mean), p(cov), None))
zs = z0 + torch.randn_like(z0) * torch.tensor([2.0, 2.0, 0.01, 2.0], device=dev, dtype=torch.float64)
state_bytes = N * (8 + 64) * 8
big = state_bytes * 2 > (256 << 20)
l2note = "state %d MB > L2" % (state_bytes >> 20) if big else "256 MB L2 flush between iterations"
want = set(args.only.split(",")) if args.only else None


def on(name):
    return want is None or name in want


if on("kf_predict"):
    ms, best = timeit(lambda: _lib.check(lib.b200track_kf_predict(KIND, N, p(mean), p(cov), None)), not big)
    report("kf_predict_kernel (multi_predict, dense 8x8 layout)", ms, best, alg_bytes=2 * state_bytes, l2=l2note,
           extra={"tracks": N, "bytes_per_track": 2 * 576})
if on("kf_update"):
    for _ in range(2):
        _lib.check(lib.b200track_kf_predict(KIND, N, p(mean), p(cov), None))
    ms, best = timeit(lambda: _lib.check(lib.b200track_kf_update(KIND, N, p(mean), p(cov), p(zs), None, None)), not big)
    report("kf_update_kernel (project + Cholesky update, dense layout)", ms, best, alg_bytes=2 * state_bytes + N * 32, l2=l2note,
           extra={"tracks": N, "bytes_per_track": 2 * 576 + 32})
if on("kf_project"):
    pm = torch.empty((N, 4), dtype=torch.float64, device=dev)
    pc = torch.empty((N, 4, 4), dtype=torch.float64, device=dev)
    ms, best = timeit(lambda: _lib.check(lib.b200track_kf_project(KIND, N, p(mean), p(cov), None, p(pm), p(pc), None)), not big)
    report("kf_project_kernel (reads mean[:4] + the 4x4 block: 160 B, writes 160 B per track)", ms, best, alg_bytes=N * 320, l2=l2note,
           extra={"tracks": N, "bytes_per_track": 320})

if on("kf8"):
    # DeepOCSORT's 8-d filter (csrc/kf8.cu): same dense layout and bytes as the lines above, one thread per track
    m8 = mean.clone()
    m8[:, 2] = torch.from_numpy(rng.uniform(30, 90, N)).to(dev)
    c8 = cov.clone()
    z8 = (m8[:, :4] + torch.randn((N, 4), device=dev, dtype=torch.float64) * 2).contiguous()
    wh8 = m8[:, 2:4].contiguous()
    ms, best = timeit(lambda: _lib.check(lib.b200track_kf8_predict(N, p(m8), p(c8), 0, None)), not big)
    report("kf8_predict_kernel (DeepOCSORT filter, Q from w, h)", ms, best, alg_bytes=2 * state_bytes, l2=l2note,
           extra={"tracks": N, "bytes_per_track": 2 * 576})
    ms, best = timeit(lambda: _lib.check(lib.b200track_kf8_update(N, p(m8), p(c8), p(z8), p(wh8), None)), not big)
    report("kf8_update_kernel (Joseph form, Cholesky solves per lane, R from w, h)", ms, best, alg_bytes=2 * state_bytes + N * 48, l2=l2note,
           extra={"tracks": N, "bytes_per_track": 2 * 576 + 48})
    del m8, c8, z8, wh8

# ---- config 4 ------------------------------------------------------------------------------------------
mean4 = mean[:S * T].view(S, T, 8)
meas = (mean4[:, torch.randint(0, T, (D,), device=dev), :4] + torch.randn((S, D, 4), device=dev, dtype=torch.float64) * 4).contiguous()
gd = torch.empty((S, T, D), dtype=torch.float64, device=dev)
if on("gating"):
    ms, best = timeit(lambda: _lib.check(lib.b200track_kf_gating_distance_batched(KIND, S, T, D, p(mean), p(cov), p(meas), 0, 0, None,
                                                                                  p(gd), None)), False)
    report("kf_gating_kernel (config 4: squared Mahalanobis, 4 dof)", ms, best, alg_bytes=S * T * D * 8 + S * T * 576 + S * D * 32,
           l2="output %d MB > L2" % ((S * T * D * 8) >> 20), extra={"pairs": S * T * D, "pairs_per_s": S * T * D / (ms * 1e-3),
                                                                     "gated_in_frac": float((gd <= 9.4877).double().mean().item())})
proto = torch.randn((S, max(T, D), F), device=dev)
trk = (proto[:, :T] + 0.3 * torch.randn((S, T, F), device=dev)).contiguous()
det = (proto[:, :D] + 0.3 * torch.randn((S, D, F), device=dev)).contiguous()
out = torch.empty((S, T, D), dtype=torch.float64, device=dev)
nbytes = C.c_uint64()
_lib.check(lib.b200track_appearance_cost_workspace(S, T, D, F, C.byref(nbytes)))
ws = torch.empty((int(nbytes.value),), dtype=torch.uint8, device=dev)
stats = torch.zeros((2,), dtype=torch.int64, device=dev)
gate = (gd > 9.4877).to(torch.uint8).contiguous()                   # chi2inv95[4], matching.py:15-25
flops = 2.0 * S * T * D * F
emb_in = S * (T + D) * F * 4
for name, thr, gt in (("cosine cost, appearance_thresh 0.25 (BoT-SORT default), no gate", 0.25, None),
                      ("cosine cost, appearance_thresh 0.4818 (botsort.yaml), no gate", 0.4818211117541298, None),
                      ("cosine cost, appearance_thresh 0.4818, Mahalanobis gate fused as mask", 0.4818211117541298, gate)):
    if not on("appearance"):
        break
    stats.zero_()
    fn = lambda: _lib.check(lib.b200track_appearance_cost(S, T, D, F, p(trk), p(det), p(gt), 0.5, thr, 1.0, p(out), p(ws),
                                                          int(nbytes.value), p(stats), None))
    ms, best = timeit(fn, False)
    n_calls = 3 + args.iters
    exact = int(stats[0].item()) / n_calls
    report("unit_bf16_kernel x2 + appearance_cost_kernel (tcgen05): " + name, ms, best, flops=flops,
           alg_bytes=emb_in + S * T * D * 8 + (S * T * D if gt is not None else 0),
           l2="inputs + output %d MB > L2" % ((emb_in + S * T * D * 8) >> 20),
           extra={"pairs": S * T * D, "pairs_per_s": S * T * D / (ms * 1e-3), "exact_rechecks_per_launch": exact,
                  "exact_frac": exact / (S * T * D), "errors": int(stats[1].item())})
if on("embedding_fp64"):
    S2 = min(S, 128)
    fn = lambda: [_lib.check(lib.b200track_embedding_distance(T, D, F, p(trk[b]), p(det[b]), p(out[b]), None)) for b in range(S2)]
    ms, best = timeit(fn, True)
    report("embedding_distance_kernel (dense fp64 on fp32 inputs, %d streams = %d launches)" % (S2, S2), ms, best,
           extra={"pairs": S2 * T * D, "pairs_per_s": S2 * T * D / (ms * 1e-3), "fp64_gflops": 2.0 * S2 * T * D * F / (ms * 1e-3) / 1e9})
if on("config4_pair"):
    def both():
        _lib.check(lib.b200track_kf_gating_distance_batched(KIND, S, T, D, p(mean), p(cov), p(meas), 0, 0, None, p(gd), None))
        _lib.check(lib.b200track_appearance_cost(S, T, D, F, p(trk), p(det), p(gate), 0.5, 0.4818211117541298, 1.0, p(out), p(ws),
                                                 int(nbytes.value), p(stats), None))
    ms, best = timeit(both, False)
    report("config 4 back to back: gating_distance + gated cosine cost", ms, best,
           extra={"streams": S, "tracks": T, "dets": D, "dim": F, "track_updates_per_s": S * T / (ms * 1e-3)})

if on("gallery"):
    # StrongSORT's gallery distance: T tracks x G stored features each against D detections per stream.
    # The fp32 gallery and its unit-norm bf16 copy are tracker state (one row per confirmed track changes per frame), so
    # both are resident; the detections are converted inside the timed region.
    S3, G = args.gallery_streams, args.budget
    del trk, out, ws, gd, gate
    torch.cuda.empty_cache()
    proto3 = torch.randn((S3, max(T, D), F), device=dev)
    gal = torch.empty((S3, T, G, F), dtype=torch.float32, device=dev)
    for b0 in range(0, S3, 16):                              # chunked: 16 streams of noise at a time
        b1 = min(S3, b0 + 16)
        gal[b0:b1] = proto3[b0:b1, :T, None, :] + 0.25 * torch.randn((b1 - b0, T, G, F), device=dev)
    det3 = (proto3[:, :D] + 0.45 * torch.randn((S3, D, F), device=dev)).contiguous()
    cnt3 = torch.full((S3, T), G, dtype=torch.int32, device=dev)
    gal16 = torch.empty((S3, T, G, F), dtype=torch.bfloat16, device=dev)
    _lib.check(lib.b200track_unit_bf16(S3 * T * G, F, p(gal), p(gal16), None))
    out3 = torch.empty((S3, T, D), dtype=torch.float64, device=dev)
    nb3 = C.c_uint64()
    _lib.check(lib.b200track_gallery_cost_workspace(S3, T, G, D, F, 0, C.byref(nb3)))
    ws3 = torch.empty((int(nb3.value),), dtype=torch.uint8, device=dev)
    st3 = torch.zeros((3,), dtype=torch.int64, device=dev)
    fn = lambda: _lib.check(lib.b200track_gallery_cost(S3, T, G, D, F, p(gal), p(gal16), p(cnt3), p(det3), 0.2, 0.20001, p(out3), p(ws3),
                                                       int(nb3.value), p(st3), None))
    ms, best = timeit(fn, False)
    n_calls = 3 + args.iters
    flops3 = 2.0 * S3 * T * G * D * F
    rows3 = int(st3[0].item()) / n_calls
    line_bytes = S3 * T * G * F * 2 + S3 * D * F * 4 + S3 * T * D * 8 + rows3 * F * 4
    report("unit_bf16_kernel (detections) + gallery_cost_kernel (tcgen05): StrongSORT gallery distance, max_dist 0.2", ms, best,
           flops=flops3, alg_bytes=line_bytes, l2="bf16 gallery %d MB > L2" % ((S3 * T * G * F * 2) >> 20),
           extra={"streams": S3, "tracks": T, "budget": G, "dets": D, "dim": F, "pairs": S3 * T * D, "pairs_per_s": S3 * T * D / (ms * 1e-3),
                  "track_updates_per_s": S3 * T / (ms * 1e-3), "surviving_pairs_per_launch": int(st3[2].item()) / n_calls,
                  "exact_rows_per_launch": rows3, "errors": int(st3[1].item()),
                  "kept_frac": float((out3 <= 0.2).double().mean().item())})
