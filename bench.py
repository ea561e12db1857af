#!/usr/bin/env python
"""Headline benchmark: track-updates/s of the batched ByteTrack frame step (BASELINE.json
config 5: 4096 streams x 200 objects per GPU, sharded by stream, weak scaling).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

A "step" is one frame over every stream of the rank.  `value` is measured with the
detections of all frames already resident in HBM; `e2e` drives the same frames through the
packed host interface of the C-ABI (b200track_submit_packed: one pinned input block -> ONE H2D
copy -> step -> ONE D2H copy of the compact result block) with a 3-deep pipeline.  Both legs
start from the same steady state: PREROLL untimed frames (track_buffer ages out lost tracks
from frame 31 on) run before the --warmup frames.  Detections are fp32-representable (what a
detector emits), so the fp32 packed path, the fp64 padded path and the CPU arm see identical values.
`--impl reference` times the CPU oracle port of the reference on the host cores.
One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import mmap
import multiprocessing as mp
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
# stdout carries exactly one JSON line.  Libraries print there too (NCCL's version banner is a plain printf), so the
# process's fd 1 is pointed at stderr for its whole life and the result line is written to the saved descriptor.
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
_REAL_STDOUT = None


def _claim_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)

# Workloads (BASELINE.json configs).  The headline (default) is config 5; the others are reported in DESIGN.md.
# b_slot: algorithmic HBM bytes of one track slot, in + out; b_feat: embedding bytes per
# track-update (smoothed row read + detection row read + smoothed row written).
WORKLOADS = {
    "bytetrack": dict(kind="bytetrack", config=5, objects=200, streams=4096, max_dets=224, max_tracks=224, emb=0,
                      params=dict(track_thresh=0.5, match_thresh=0.8, track_buffer=30, frame_rate=30),      # bytetrack.yaml
                      b_slot=2 * 200, b_feat=0, kernel="bytetrack_step_kernel", img_hw=(0, 0),
                      label="config5: ByteTrack multi-stream (bytetrack.yaml: track_thresh 0.5, match_thresh 0.8, "
                            "track_buffer 30), 200 objects/stream, sharded by stream"),
    "ocsort": dict(kind="ocsort", config=2, objects=100, streams=64, max_dets=128, max_tracks=256, emb=0,
                   params=dict(det_thresh=0, max_age=30, min_hits=1, asso_threshold=0.3, delta_t=3, asso_func="giou",
                               inertia=0.2), occlusion=True,                                                  # ocsort.yaml
                   b_slot=2 * (37 * 8 + 10 * 4), b_feat=0, kernel="ocsort_step_kernel", img_hw=(2160, 3840),
                   label="config2: OC-SORT (ocsort.yaml: giou, det_thresh 0, min_hits 1, max_age 30, delta_t 3, inertia 0.2), "
                         "100 objects/stream with occlusion runs, sharded by stream"),
    "botsort": dict(kind="botsort", config=3, objects=100, streams=256, max_dets=128, max_tracks=224, emb=512,
                    params=dict(track_high_thresh=0.33824964456239337, track_low_thresh=0.1,
                                new_track_thresh=0.21144301345190655, track_buffer=60, match_thresh=0.22734550911325851,
                                proximity_thresh=0.5945380911899254, appearance_thresh=0.4818211117541298, frame_rate=30),
                    b_slot=2 * (22 * 8 + 7 * 4 + 72), b_feat=3 * 512 * 4, kernel="bytetrack_step_kernel<BOT>", img_hw=(0, 0),
                    label="config3: BoT-SORT (botsort.yaml, identity camera motion, 512-d appearance embeddings), "
                          "100 objects/stream, sharded by stream"),
    # config 4 as a TRACKER run: DeepOCSORT (Mahalanobis-free, appearance-weighted association) at 1024 streams.  det_thresh 0
    # (deepocsort.yaml) keeps every false positive alive for max_age frames, so 190 objects give ~220 live trackers per stream
    # in 256 slots and ~182 detections per frame.
    "deepocsort": dict(kind="deepocsort", config=4, objects=190, streams=1024, max_dets=224, max_tracks=256, emb=512, distinct=128,
                       steps=20, warmup=5,
                       params=dict(det_thresh=0, max_age=30, min_hits=1, iou_threshold=0.3, delta_t=3, asso_func="giou", inertia=0.2,
                                   w_association_emb=0.5, alpha_fixed_emb=0.95, aw_param=0.5),      # deepocsort.yaml through the factory
                       b_slot=2 * (48 * 8 + 11 * 4), b_feat=512 * (8 + 8 + 4), kernel="deepocsort_step_kernel", img_hw=(2160, 3840),
                       label="config4 as a tracker: DeepOCSORT (deepocsort.yaml: giou, det_thresh 0, min_hits 1, max_age 30, adaptive "
                             "appearance weight, 512-d embeddings), 190 objects + ~30 false-positive trackers per stream, sharded by stream"),
    # config 4 in its StrongSORT form: Mahalanobis gate + gallery cosine cost (100 stored features per track) at 200 x 200;
    # the batched step is eight launches per frame, dominated by the tensor-core gallery distance.  256 streams: the gallery
    # ring alone is 256 x 256 x 100 x 512 x 6 B = 20 GB of HBM.  Padded host interface (the context takes no packed frames).
    "strongsort": dict(kind="strongsort", config=4, objects=190, streams=256, max_dets=224, max_tracks=256, emb=512, distinct=64,
                       steps=10, warmup=3, padded_e2e=True, cpu_sample=(6, 2),
                       params=dict(max_dist=0.2, max_iou_dist=0.7, max_age=30, n_init=1, nn_budget=100, mc_lambda=0.995, ema_alpha=0.8),  # strongsort.yaml
                       b_slot=2 * (72 * 8 + 2 * 8 + 7 * 4) + 2 * 576, b_feat=512 * (4 + 4 + 4 + 4 + 2), kernel="gallery_cost_kernel (+ 7 smaller launches)",
                       img_hw=(2160, 3840),
                       label="config4 as a tracker: StrongSORT (strongsort.yaml: max_dist 0.2, nn_budget 100, mc_lambda 0.995, ema_alpha 0.8; "
                             "Mahalanobis gate + gallery cosine cost, 512-d embeddings), 190 objects per stream, sharded by stream"),
    # config 4 through HybridSORT: every (detection, tracker) pair carries an appearance term (a dense 512-d cosine matrix per
    # stream and frame) next to the four-corner direction cost; padded host interface (the context takes no packed frames).
    "hybridsort": dict(kind="hybridsort", config=4, objects=190, streams=512, max_dets=224, max_tracks=256, emb=512, distinct=128,
                       steps=20, warmup=5, padded_e2e=True, cpu_sample=(10, 3),
                       params=dict(det_thresh=0, max_age=30, min_hits=1, iou_threshold=0.3, delta_t=3, asso_func="giou", inertia=0.2),  # hybridsort.yaml
                       b_slot=2 * (48 * 8 + 11 * 4), b_feat=512 * (4 + 4 + 4), kernel="hybridsort_step_kernel", img_hw=(2160, 3840),
                       label="config4 as a tracker: HybridSORT (hybridsort.yaml: giou, det_thresh 0, min_hits 1, max_age 30; four-corner "
                             "direction cost + 1.3 x cosine distance on 512-d embeddings for every pair), 190 objects + ~30 false-positive "
                             "trackers per stream, sharded by stream"),
}
W = dict(WORKLOADS["bytetrack"])          # the active workload (set in main)
CONFIG_ID, N_OBJECTS, STREAMS_PER_GPU = W["config"], W["objects"], W["streams"]
MAX_DETS, MAX_TRACKS, PARAMS = W["max_dets"], W["max_tracks"], W["params"]
B_DET, B_ROW = 48, 64


def select_workload(name):
    global CONFIG_ID, N_OBJECTS, STREAMS_PER_GPU, MAX_DETS, MAX_TRACKS, PARAMS
    W.clear()
    W.update(WORKLOADS[name])
    CONFIG_ID, N_OBJECTS, STREAMS_PER_GPU = W["config"], W["objects"], W["streams"]
    MAX_DETS, MAX_TRACKS, PARAMS = W["max_dets"], W["max_tracks"], W["params"]


PREROLL = 40      # untimed frames before the warm-up: the timed window then holds lost, re-found and aged-out tracks


def _stream_inputs(stream, n_frames):
    """dets[F, MAX_DETS, 6], ndets[F], seam features[F, MAX_DETS, emb] or None for one stream of the workload."""
    from yolo_tracking_b200.synth import make_stream
    kw = dict(occlusion=True) if W.get("occlusion") else {}
    d, n, e = make_stream(CONFIG_ID, stream, N_OBJECTS, n_frames, dmax=MAX_DETS, emb_dim=W["emb"], **kw)
    d = d.astype(np.float32).astype(np.float64)          # detector output precision; every consumer sees these values
    if e is not None:
        # the ReID seam (reid_multibackend.py:304-311): first-round rows / Frobenius norm of their matrix
        high = PARAMS["track_high_thresh"] if W["kind"] == "botsort" else PARAMS.get("det_thresh", -1.0)    # StrongSORT: every row
        feats = np.zeros_like(e)
        for f in range(n_frames):
            rows = np.nonzero(d[f, :n[f], 4] > high)[0]
            if len(rows):
                feats[f, rows] = e[f, rows] / np.linalg.norm(e[f, rows])
        e = feats
    return d, n, e


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ----------------------------------------------------------------------------- data
_shared = {}


def _gen_worker(args):
    first, count, stream0, n_frames = args
    dets, nd, feats = _shared["dets"], _shared["nd"], _shared.get("feats")
    for i in range(first, first + count):
        d, n, e = _stream_inputs(stream0 + i, n_frames)
        dets[:, i] = d
        nd[:, i] = n
        if feats is not None:
            feats[:, i] = e
    return count


def generate(n_streams, stream0, n_frames, workers):
    """dets[F, S, MAX_DETS, 6] f32, ndets[F, S] i32 (and feats[F, S, MAX_DETS, emb] f32) in fork-shared anonymous memory."""
    nbytes = n_frames * n_streams * MAX_DETS * 6 * 4
    buf = mmap.mmap(-1, nbytes)
    buf2 = mmap.mmap(-1, n_frames * n_streams * 4)
    dets = np.frombuffer(buf, dtype=np.float32).reshape(n_frames, n_streams, MAX_DETS, 6)
    nd = np.frombuffer(buf2, dtype=np.int32).reshape(n_frames, n_streams)
    _shared["dets"], _shared["nd"] = dets, nd
    _shared.pop("feats", None)
    if W["emb"]:
        buf3 = mmap.mmap(-1, n_frames * n_streams * MAX_DETS * W["emb"] * 4)
        _shared["feats"] = np.frombuffer(buf3, dtype=np.float32).reshape(n_frames, n_streams, MAX_DETS, W["emb"])
    workers = max(1, min(workers, n_streams))
    chunk = max(1, (n_streams + workers * 4 - 1) // (workers * 4))
    tasks = [(i, min(chunk, n_streams - i), stream0, n_frames) for i in range(0, n_streams, chunk)]
    if workers == 1:
        for t in tasks:
            _gen_worker(t)
    else:
        with mp.get_context("fork").Pool(workers) as pool:
            pool.map(_gen_worker, tasks)
    return dets, nd, _shared.get("feats")


# ----------------------------------------------------------------------------- CPU legs
def _make_oracle():
    if W["kind"] == "bytetrack":
        from oracle.bytetrack import ByteTrackOracle
        return ByteTrackOracle(**PARAMS)
    if W["kind"] == "ocsort":
        from oracle.ocsort import OCSortOracle
        return OCSortOracle(False, use_byte=False, **PARAMS)
    if W["kind"] == "deepocsort":
        from oracle.deepocsort import DeepOCSortOracle
        return DeepOCSortOracle(**PARAMS)
    if W["kind"] == "strongsort":
        from oracle.strongsort import StrongSORTOracle
        return StrongSORTOracle(**PARAMS)
    if W["kind"] == "hybridsort":
        from oracle.hybridsort import HybridSortOracle
        return HybridSortOracle(**PARAMS)
    from oracle.botsort import BoTSORTOracle
    return BoTSORTOracle(**PARAMS)


def _oracle_worker(args):
    streams, n_frames, warm = args
    try:                                          # one BLAS thread per worker process: the workers already fill the cores
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)
    except Exception:
        pass
    data = [_stream_inputs(s, n_frames) for s in streams]
    trks = [_make_oracle() for _ in streams]

    def step(f):
        for k, t in enumerate(trks):
            d, n, e = data[k]
            if W["kind"] in ("deepocsort", "hybridsort"):
                rows = d[f, :n[f], 4] > PARAMS["det_thresh"]
                t.update(d[f, :n[f]], e[f, :n[f]][rows], W["img_hw"])
                continue
            second = W["img_hw"] if W["kind"] == "ocsort" else (e[f, :n[f]] if e is not None else None)
            t.update(d[f, :n[f]], second)
    for f in range(warm):                         # pre-roll + warm-up frames, untimed
        step(f)
    base = sum(t.track_updates for t in trks)
    t0 = time.perf_counter()
    for f in range(warm, n_frames):
        step(f)
    dt = time.perf_counter() - t0
    return sum(t.track_updates for t in trks) - base, dt


def cpu_oracle_run(n_workers, streams_per_worker, steps, warmup):
    """The oracle port (oracle/<tracker>.py) on `n_workers` processes; returns
    (track_updates, wall_seconds = slowest worker)."""
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    tasks = [([100000 + w * streams_per_worker + k for k in range(streams_per_worker)], steps + warmup + PREROLL, warmup + PREROLL)
             for w in range(n_workers)]
    if n_workers == 1:
        res = [_oracle_worker(tasks[0])]
    else:
        with mp.get_context("fork").Pool(n_workers) as pool:
            res = pool.map(_oracle_worker, tasks)
    return sum(r[0] for r in res), max(r[1] for r in res)


# ----------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake_slowdown",
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                break
            time.sleep(0.005)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ----------------------------------------------------------------------------- main arms
def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = host_cores()
    spw = 2
    tu, dt = cpu_oracle_run(cores, spw, args.steps, args.warmup)
    val = tu / dt
    line = {
        "impl": "reference", "metric": "track-updates/s", "value": val, "unit": "track-updates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": W["label"], "streams_per_gpu": args.streams, "sample_streams": cores * spw, "objects": N_OBJECTS},
        "cpu_baseline": {"value": val, "unit": "track-updates/s", "cores": cores, "kind": "port",
                         "sample": f"{cores * spw} streams x {args.steps} frames after {PREROLL} pre-roll + {args.warmup} warm-up frames, "
                                   f"oracle/{W['kind']}.py (numpy port of the reference; /root/reference is Python and "
                                   f"cannot travel), one process per core"},
        "e2e": {"value": val, "unit": "track-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def run_b200(args, rank, world, local_rank):
    S = args.streams
    W0 = PREROLL + args.warmup                   # first timed frame
    F = W0 + args.steps
    cores = host_cores()
    gen_workers = max(1, cores // max(1, world))
    t_gen = time.time()
    from yolo_tracking_b200.shard import shard_bounds
    stream0, stream1 = shard_bounds(S * world, rank, world)          # weak scaling: S streams per GPU, block-sharded
    assert stream1 - stream0 == S
    # workloads with large embeddings generate `distinct` streams and replicate them (physically: separate copies in HBM
    # and in the pinned blocks, so nothing is shared between replicas but the values)
    Sd = min(S, W.get("distinct", S))
    assert S % Sd == 0
    reps = S // Sd
    dets_h, nd_h, feats_h = generate(Sd, rank * Sd if reps > 1 else stream0, F, gen_workers)     # fp32 [F, Sd, D, 6]
    t_gen = time.time() - t_gen

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cs, cw = W.get("cpu_sample", (40, 10))
        tu, dt = cpu_oracle_run(cores, 1, cs, cw)
        cpu_base = {"value": tu / dt, "unit": "track-updates/s", "cores": cores, "kind": "port",
                    "sample": f"{cores} streams x {cs} frames after {PREROLL} pre-roll + {cw} warm-up frames of the same workload, "
                              f"oracle/{W['kind']}.py, one process per core"}

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from yolo_tracking_b200.batch import BatchedTracker

    dev = torch.device("cuda", local_rank)
    # device leg: all frames resident in HBM as the padded fp64 [S, max_dets, 6] blocks of b200track_step
    d_dets = torch.from_numpy(dets_h).to(dev).to(torch.float64).repeat(1, reps, 1, 1)
    d_nd = torch.from_numpy(nd_h).to(dev).repeat(1, reps)
    d_feats = torch.from_numpy(feats_h).to(dev).repeat(1, reps, 1, 1) if feats_h is not None else None
    if reps > 1:
        nd_h = np.tile(nd_h, (1, reps))
    hw = W["img_hw"]

    def feats_dev(f):
        return d_feats[f] if d_feats is not None else None

    d_out = torch.empty((S, MAX_TRACKS, 8), dtype=torch.float64, device=dev)
    d_nout = torch.empty((S,), dtype=torch.int32, device=dev)
    trk = BatchedTracker(W["kind"], S, max_tracks=MAX_TRACKS, max_dets=MAX_DETS, device=local_rank, feat_dim=W["emb"], **PARAMS)
    stream = torch.cuda.Stream(device=dev)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def preroll(upto):
        with torch.cuda.stream(stream):
            for f in range(upto):
                trk.step_device(d_dets[f], d_nd[f], d_out, d_nout, d_feats=feats_dev(f), img_hw=hw, stream=stream.cuda_stream)

    # ---------------- device-resident leg -------------------------------------------------
    preroll(W0)                                  # PREROLL + warm-up frames, untimed
    barrier()
    tu0, l0 = trk.track_updates(), trk.launches()
    gal0 = trk.counters()["gallery_rows"] if W["kind"] == "strongsort" else 0
    sampler = ClockSampler(local_rank)
    sampler.start()
    # L2 hygiene: config 5's per-step working set is several times the 126 MB L2 and every step reads new
    # detections; small workloads (config 2 / 3 at their own stream counts) get a 256 MB L2 flush between steps,
    # outside the per-step event pairs.
    working_set = S * N_OBJECTS * (W["b_slot"] + W["b_feat"] + 48 + 64)
    flush = torch.zeros(64 << 20, dtype=torch.int32, device=dev) if working_set < (256 << 20) else None
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    barrier()
    with torch.cuda.stream(stream):
        for k in range(args.steps):
            f = W0 + k
            if flush is not None:
                flush.sum()                       # read 256 MB: L2 ends up holding clean lines of another buffer
            ev0[k].record(stream)
            trk.step_device(d_dets[f], d_nd[f], d_out, d_nout, d_feats=feats_dev(f), img_hw=hw, stream=stream.cuda_stream)
            ev1[k].record(stream)
    barrier()
    sampler.stop_flag = True
    sampler.join()
    step_ms = np.array([ev0[k].elapsed_time(ev1[k]) for k in range(args.steps)])
    total_ms = ev0[0].elapsed_time(ev1[-1]) if flush is None else float(step_ms.sum())
    tu_dev = trk.track_updates() - tu0
    launches = trk.launches() - l0
    dets_timed = int(nd_h[W0:].sum())
    gal_rows = (trk.counters()["gallery_rows"] - gal0) if W["kind"] == "strongsort" else 0

    # ---------------- end-to-end leg: pinned host blocks through the packed C-ABI ----------------
    # Every timed frame is one pinned input block (offsets + fp32 detection rows [+ fp32 embeddings]) prepared before the
    # clock starts - the form a detector hands frames over in; per step: ONE H2D copy, the step, ONE D2H copy of the result
    # block (header, per-stream row counts, compact rows), and the host reads the row counts of the finished frame.
    trk.reset()
    preroll(PREROLL)
    torch.cuda.synchronize(dev)
    nslot = trk.host_slots
    n_e2e = args.warmup + args.steps
    rows_of, flags_of, in_blocks = [], [], []
    max_rows = int(nd_h[PREROLL:].sum(axis=1).max())
    padded = bool(W.get("padded_e2e"))
    if padded:
        # contexts without packed frames (StrongSORT): pinned padded blocks dets [S, D, 6] f64 + feats [S, D, F] f32 in,
        # out [S, T, 8] + nout [S] back through b200track_submit_host / wait_host (pitched copies of the used rows)
        def pinned(shape, dtype):
            t = torch.empty(shape, dtype=dtype, pin_memory=True)
            return t, t.numpy()
        keep = []
        for k in range(n_e2e):
            f = PREROLL + k
            td, a = pinned((S, MAX_DETS, 6), torch.float64)
            a[:] = np.tile(dets_h[f], (reps, 1, 1))
            tf, b = pinned((S, MAX_DETS, W["emb"]), torch.float32)
            b[:] = np.tile(feats_h[f], (reps, 1, 1))
            keep.append((td, tf))
            in_blocks.append((a, b, np.ascontiguousarray(nd_h[f])))
        outs = [pinned((S, MAX_TRACKS, 8), torch.float64) + pinned((S,), torch.int32) for _ in range(nslot)]
    for k in range(0 if padded else n_e2e):
        f = PREROLL + k
        blk, _ = trk.frame_buffers(max_rows=int(nd_h[f].sum()))
        r, fl = trk.pack(blk, np.tile(dets_h[f], (reps, 1, 1)) if reps > 1 else dets_h[f], ndets=nd_h[f],
                         feats=None if feats_h is None else (np.tile(feats_h[f], (reps, 1, 1)) if reps > 1 else feats_h[f]), dtype=np.float32)
        in_blocks.append(blk); rows_of.append(r); flags_of.append(fl)
    out_blocks = [] if padded else [trk.frame_buffers(max_rows=max_rows)[1] for _ in range(nslot)]
    del dets_h

    def submit(k, slot):
        if padded:
            a, b, n = in_blocks[k]
            trk.submit(slot, a, n, outs[slot][1], outs[slot][3], feats=b, img_hw=hw)
            return
        trk.submit_packed(slot, in_blocks[k], out_blocks[slot], np.float32, flags_of[k], img_hw=hw)

    def collect(k, slot):
        if padded:
            trk.wait(slot)
            return int(outs[slot][3].sum())
        trk.wait_packed(slot)
        v = trk.frame_views(None, out_blocks[slot], rows_of[k], np.float32)
        return int(v["nout"].sum())               # device->host read of the step's result

    for k in range(args.warmup):
        submit(k, k % nslot)
        collect(k, k % nslot)
    barrier()
    tu1 = trk.track_updates()
    h2d = d2h = rows = 0
    sampler_e2e = ClockSampler(local_rank)
    sampler_e2e.start()
    t0 = time.perf_counter()
    for k in range(args.steps):
        slot = k % nslot
        if k >= nslot:
            rows += collect(args.warmup + k - nslot, slot)
        submit(args.warmup + k, slot)
    for k in range(max(0, args.steps - nslot), args.steps):
        rows += collect(args.warmup + k, k % nslot)
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    sampler_e2e.stop_flag = True
    sampler_e2e.join()
    for k in range(args.warmup, n_e2e):
        if padded:
            mx = int(in_blocks[k][2].max())
            h2d += S * 4 + S * mx * (48 + 4 * W["emb"]); d2h += S * 4 + 4 + S * min(mx, MAX_TRACKS) * 64
            continue
        L = trk.frame_layout(rows_of[k], np.float32)
        h2d += int(L.in_bytes); d2h += int(L.out_bytes)
    row_bytes = 64 if padded else int(trk.frame_layout(0, np.float32).row_bytes)
    tu_e2e = trk.track_updates() - tu1
    trk.sync()
    assert tu_e2e == tu_dev, "the two legs ran different work"

    # ---------------- optional final gather of the padded outputs (the only collective; not part of the step) ----
    gather_ms = None
    if world > 1:
        from yolo_tracking_b200.shard import gather_outputs
        for _ in range(2):
            gather_outputs(d_out, d_nout)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        all_out, all_nout = gather_outputs(d_out, d_nout)
        g1.record()
        torch.cuda.synchronize(dev)
        assert all_out.shape[0] == S * world and all_nout.shape[0] == S * world
        gather_ms = g0.elapsed_time(g1)

    # ---------------- reduce over ranks ----------------------------------------------------
    from yolo_tracking_b200.shard import reduce_timing
    total_ms_max, (tu_all, tu_e2e_all, launches_all, rows_all, dets_all) = reduce_timing(
        total_ms, [tu_dev, tu_e2e, launches, rows, dets_timed], device=dev)
    e2e_ms_max, _ = reduce_timing(e2e_s * 1e3, [], device=dev)
    if rank == 0:
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
                peaks = json.load(fh)
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        # DRAM bytes per launch cannot be measured outside a profiler: the number comes from the ncu --set full capture of
        # this same command committed under profiles/ (tools/summarize_profiles.py writes the file), null if none matches
        traffic, traffic_src = None, None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
                tr = json.load(fh)[W["kernel"]]
            if tr["streams"] == S:
                traffic = tr["dram_bytes_per_launch"]
                traffic_src = tr.get("source", "profiles/traffic.json")
        except Exception:
            pass
        # rank-0 kernel: algorithmic bytes per launch / mean launch duration (events on the launch stream); the device leg
        # reads padded fp64 detection rows (48 B) and writes the reference's 64-byte result rows
        alg_bytes = (tu_dev * (W["b_slot"] + W["b_feat"]) + dets_timed * B_DET + rows * B_ROW) / args.steps     # rank-0 shard
        if W["kind"] == "strongsort":
            # + the stored gallery rows every confirmed track is compared against (bf16 operand of the tensor-core distance)
            # and the detection embeddings (fp32 read + bf16 copy written and read)
            alg_bytes += (gal_rows * W["emb"] * 2 + dets_timed * W["emb"] * (4 + 2 + 2)) / args.steps
        mean_ms = float(step_ms.mean())
        achieved = alg_bytes / (mean_ms * 1e-3) / 1e9
        line = {
            "metric": "track-updates/s", "value": tu_all / (total_ms_max * 1e-3), "unit": "track-updates/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": W["label"],
                       "streams_per_gpu": S, "streams_total": S * world, "objects_per_stream": N_OBJECTS,
                       "max_tracks": MAX_TRACKS, "max_dets": MAX_DETS,
                       "l2": ("per-step working set (state + detections + outputs) is ~%.0f MB > 126 MB L2; every step reads "
                              "new detections" % (working_set / 1e6)) if flush is None else
                             ("per-step working set ~%.0f MB: a 256 MB buffer is read between steps (L2 flush), outside "
                              "the per-step event pairs; value = units / sum of step times" % (working_set / 1e6)),
                       "preroll_frames": PREROLL, "detections": "fp32-representable values (detector precision)",
                       "distinct_streams_per_gpu": Sd,
                       "data_gen_s": round(t_gen, 1)},
            "p50_step_ms": float(np.percentile(step_ms, 50)), "p99_step_ms": float(np.percentile(step_ms, 99)),
            "e2e": {"value": tu_e2e_all / (e2e_ms_max * 1e-3), "unit": "track-updates/s",
                    "h2d_bytes_per_step": h2d // args.steps, "d2h_bytes_per_step": d2h // args.steps,
                    "ms_per_step": e2e_ms_max / args.steps, "pipeline_depth": nslot,
                    "output_rows_per_step": rows_all / args.steps / world,
                    "interface": ("b200track_submit_host / wait_host: pinned padded blocks (fp64 detection rows + fp32 embeddings), pitched "
                                  "copies of the used rows, 64-byte result rows; the host reads the per-stream row counts of every "
                                  "finished frame") if padded else
                                 "b200track_submit_packed / wait_packed: one pinned input block per frame (int32 offsets + fp32 "
                                 "detection rows%s), one linear cudaMemcpyAsync per direction, compact %d-byte result rows; "
                                 "the host reads the per-stream row counts of every finished frame"
                                 % (" + fp32 embeddings" if W["emb"] else "", row_bytes),
                    "clocks": sampler_e2e.summary()},
            "gpu_launches": int(launches_all),
            "gather": None if gather_ms is None else {"ms": gather_ms, "bytes_per_rank": S * MAX_TRACKS * 64 + 4 * S,
                                                       "what": "NCCL all_gather of out[S, max_tracks, 8] + nout[S] after the run (optional, outside the step)"},
            "roofline": {"bound": "hbm", "kernel": W["kernel"], "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650",
                         "alg_bytes_per_launch": alg_bytes, "mean_launch_ms": mean_ms,
                         "bytes_per_track_update": alg_bytes * args.steps / max(1, tu_dev)},
            "cpu_baseline": cpu_base,
            "clocks": sampler.summary(),
        }
        if W["kind"] == "strongsort":
            # the GEMM-shaped part: every stored gallery row against every detection of its stream (2 F flop per pair)
            flops = 2.0 * gal_rows * W["emb"] * (dets_timed / max(1, S * args.steps)) / args.steps
            line["tensor"] = {"what": "gallery rows x detections x 2 F per step, over the whole step time (eight launches)",
                              "gallery_rows_per_step": gal_rows / args.steps, "tflops_over_step": flops / (mean_ms * 1e-3) / 1e12,
                              "peak_bf16_tflops": peaks.get("bf16_tflops")}
        if W["kind"] == "hybridsort":
            # the step's arithmetic: one exact (fp64-accumulated) 512-d dot product per (detection, live tracker) pair; pairs
            # estimated from the mean detections and live trackers per stream and step
            pairs = (dets_timed / max(1, S * args.steps)) * (tu_dev / max(1, S * args.steps)) * S
            flops = 2.0 * W["emb"] * pairs
            line["fp64"] = {"what": "dense cosine matrix: 2 F flop per (detection, tracker) pair, over the whole step time",
                            "pairs_per_step_approx": pairs, "tflops_over_step": flops / (mean_ms * 1e-3) / 1e12,
                            "peak_fp64_tflops": 34.0, "peak_source": "profiles/r01_fp64_rate.txt (tools/micro/fp64_rate.cu on this B200 pool)"}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="default 100 (20 for the embedding-heavy deepocsort workload)")
    ap.add_argument("--warmup", type=int, default=None, help="default 20 (5 for deepocsort)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="bytetrack", choices=sorted(WORKLOADS),
                    help="bytetrack = BASELINE config 5 (the headline); ocsort = config 2; botsort = config 3; deepocsort = config 4 as a tracker")
    ap.add_argument("--streams", type=int, default=0, help="streams per GPU (default: the workload's own count)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    select_workload(args.workload)
    if args.streams <= 0:
        args.streams = STREAMS_PER_GPU
    if args.steps is None:
        args.steps = W.get("steps", 100)
    if args.warmup is None:
        args.warmup = W.get("warmup", 20)
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    _claim_stdout()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
