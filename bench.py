#!/usr/bin/env python
"""Headline benchmark: track-updates/s of the batched ByteTrack frame step (BASELINE.json
config 5: 4096 streams x 200 objects per GPU, sharded by stream, weak scaling).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

A "step" is one frame over every stream of the rank.  `value` is measured with the
detections of all frames already resident in HBM; `e2e` drives the same frames through the
host-buffer C-ABI (pinned host memory -> H2D -> step -> D2H) with a 3-deep pipeline.
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

CONFIG_ID = 5
N_OBJECTS = 200
STREAMS_PER_GPU = 4096
MAX_DETS = 224
MAX_TRACKS = 256
PARAMS = dict(track_thresh=0.5, match_thresh=0.8, track_buffer=30, frame_rate=30)   # bytetrack.yaml
# algorithmic HBM bytes: track slot in+out, detection row in, output row out
B_SLOT, B_DET, B_ROW = 2 * 200, 48, 64


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ----------------------------------------------------------------------------- data
_shared = {}


def _gen_worker(args):
    first, count, stream0, n_frames = args
    from yolo_tracking_b200.synth import make_stream
    dets, nd = _shared["dets"], _shared["nd"]
    for i in range(first, first + count):
        d, n, _ = make_stream(CONFIG_ID, stream0 + i, N_OBJECTS, n_frames, dmax=MAX_DETS)
        dets[:, i] = d
        nd[:, i] = n
    return count


def generate(n_streams, stream0, n_frames, workers):
    """dets[F, S, MAX_DETS, 6] f64, ndets[F, S] i32 in fork-shared anonymous memory."""
    nbytes = n_frames * n_streams * MAX_DETS * 6 * 8
    buf = mmap.mmap(-1, nbytes)
    buf2 = mmap.mmap(-1, n_frames * n_streams * 4)
    dets = np.frombuffer(buf, dtype=np.float64).reshape(n_frames, n_streams, MAX_DETS, 6)
    nd = np.frombuffer(buf2, dtype=np.int32).reshape(n_frames, n_streams)
    _shared["dets"], _shared["nd"] = dets, nd
    workers = max(1, min(workers, n_streams))
    chunk = max(1, (n_streams + workers * 4 - 1) // (workers * 4))
    tasks = [(i, min(chunk, n_streams - i), stream0, n_frames) for i in range(0, n_streams, chunk)]
    if workers == 1:
        for t in tasks:
            _gen_worker(t)
    else:
        with mp.get_context("fork").Pool(workers) as pool:
            pool.map(_gen_worker, tasks)
    return dets, nd


# ----------------------------------------------------------------------------- CPU legs
def _oracle_worker(args):
    streams, n_frames, warm = args
    from oracle.bytetrack import ByteTrackOracle
    from yolo_tracking_b200.synth import make_stream
    data = [make_stream(CONFIG_ID, s, N_OBJECTS, n_frames, dmax=MAX_DETS) for s in streams]
    trks = [ByteTrackOracle(**PARAMS) for _ in streams]
    for f in range(warm):
        for k, t in enumerate(trks):
            t.update(data[k][0][f, :data[k][1][f]], None)
    base = sum(t.track_updates for t in trks)
    t0 = time.perf_counter()
    for f in range(warm, n_frames):
        for k, t in enumerate(trks):
            t.update(data[k][0][f, :data[k][1][f]], None)
    dt = time.perf_counter() - t0
    return sum(t.track_updates for t in trks) - base, dt


def cpu_oracle_run(n_workers, streams_per_worker, steps, warmup):
    """The oracle port (oracle/bytetrack.py) on `n_workers` processes; returns
    (track_updates, wall_seconds = slowest worker)."""
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    tasks = [([100000 + w * streams_per_worker + k for k in range(streams_per_worker)], steps + warmup, warmup)
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
        "config": {"workload": "config5: ByteTrack 4096 streams/GPU x 200 objects (bytetrack.yaml)",
                   "sample_streams": cores * spw, "objects": N_OBJECTS},
        "cpu_baseline": {"value": val, "unit": "track-updates/s", "cores": cores, "kind": "port",
                         "sample": f"{cores * spw} streams x {args.steps} frames after {args.warmup} warm-up frames, "
                                   f"oracle/bytetrack.py (numpy port of the reference; /root/reference is Python and "
                                   f"cannot travel), one process per core"},
        "e2e": {"value": val, "unit": "track-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def run_b200(args, rank, world, local_rank):
    S = args.streams
    F = args.steps + args.warmup
    cores = host_cores()
    gen_workers = max(1, cores // max(1, world))
    t_gen = time.time()
    dets_h, nd_h = generate(S, rank * S, F, gen_workers)
    t_gen = time.time() - t_gen

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        tu, dt = cpu_oracle_run(cores, 1, 40, 10)
        cpu_base = {"value": tu / dt, "unit": "track-updates/s", "cores": cores, "kind": "port",
                    "sample": f"{cores} streams x 40 frames after 10 warm-up frames of the same workload, "
                              f"oracle/bytetrack.py, one process per core"}

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from yolo_tracking_b200.batch import BatchedTracker

    dev = torch.device("cuda", local_rank)
    # page-locked copies of the generated frames: source of the per-step H2D copies of the e2e leg
    pin_all = torch.empty(dets_h.shape, dtype=torch.float64, pin_memory=True)
    pin_all.numpy()[...] = dets_h
    pin_nd_all = torch.empty(nd_h.shape, dtype=torch.int32, pin_memory=True)
    pin_nd_all.numpy()[...] = nd_h
    dets_h, nd_h = pin_all.numpy(), pin_nd_all.numpy()
    d_dets = pin_all.to(dev)                            # all frames resident in HBM for the device leg
    d_nd = pin_nd_all.to(dev)
    d_out = torch.empty((S, MAX_TRACKS, 8), dtype=torch.float64, device=dev)
    d_nout = torch.empty((S,), dtype=torch.int32, device=dev)
    trk = BatchedTracker("bytetrack", S, max_tracks=MAX_TRACKS, max_dets=MAX_DETS, device=local_rank, **PARAMS)
    stream = torch.cuda.Stream(device=dev)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---------------- device-resident leg -------------------------------------------------
    with torch.cuda.stream(stream):
        for f in range(args.warmup):
            trk.step_device(d_dets[f], d_nd[f], d_out, d_nout, stream=stream.cuda_stream)
    barrier()
    tu0, l0 = trk.track_updates(), trk.launches()
    rows = 0
    sampler = ClockSampler(local_rank)
    sampler.start()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    nout_acc = torch.zeros((), dtype=torch.int64, device=dev)
    barrier()
    with torch.cuda.stream(stream):
        evs[0].record(stream)
        for k in range(args.steps):
            f = args.warmup + k
            trk.step_device(d_dets[f], d_nd[f], d_out, d_nout, stream=stream.cuda_stream)
            evs[k + 1].record(stream)
    barrier()
    sampler.stop_flag = True
    sampler.join()
    step_ms = np.array([evs[k].elapsed_time(evs[k + 1]) for k in range(args.steps)])
    total_ms = evs[0].elapsed_time(evs[-1])
    tu_dev = trk.track_updates() - tu0
    launches = trk.launches() - l0
    dets_timed = int(nd_h[args.warmup:].sum())
    # output rows of the timed region are re-counted in the e2e leg (same frames, same results)

    # ---------------- end-to-end leg: pinned host buffers through the C-ABI ----------------
    trk.reset()
    nslot = trk.host_slots
    pin_out = [torch.empty((S, MAX_TRACKS, 8), dtype=torch.float64).pin_memory() for _ in range(nslot)]
    pin_nout = [torch.empty((S,), dtype=torch.int32).pin_memory() for _ in range(nslot)]

    def submit(f, slot):
        trk.submit(slot, dets_h[f], nd_h[f], pin_out[slot].numpy(), pin_nout[slot].numpy())

    for f in range(args.warmup):
        submit(f, f % nslot)
        trk.wait(f % nslot)
    barrier()
    tu1 = trk.track_updates()
    h2d = d2h = 0
    t0 = time.perf_counter()
    for k in range(args.steps):
        slot = k % nslot
        if k >= nslot:
            trk.wait(slot)
            rows += int(pin_nout[slot].numpy().sum())      # device->host read of the step's result
        submit(args.warmup + k, slot)
        h2d += int(nd_h[args.warmup + k].max()) * 48 * S + 4 * S
        d2h += S * MAX_TRACKS * 64 + 4 * S
    for k in range(max(0, args.steps - nslot), args.steps):
        trk.wait(k % nslot)
        rows += int(pin_nout[k % nslot].numpy().sum())
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    tu_e2e = trk.track_updates() - tu1
    trk.sync()

    # ---------------- reduce over ranks ----------------------------------------------------
    vals = torch.tensor([total_ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    sums = torch.tensor([tu_dev, tu_e2e, launches, rows, dets_timed], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    total_ms_max, e2e_ms_max = vals.tolist()
    tu_all, tu_e2e_all, launches_all, rows_all, dets_all = sums.tolist()
    if rank == 0:
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
                peaks = json.load(fh)
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
                tr = json.load(fh)["bytetrack_step_kernel"]
            if tr["streams"] == S:
                traffic = tr["dram_bytes_per_launch"]
        except Exception:
            pass
        # rank-0 kernel: algorithmic bytes per launch / mean launch duration (events on the launch stream)
        alg_bytes = (tu_dev * B_SLOT + dets_timed * B_DET + rows * B_ROW) / args.steps     # rank-0 shard
        mean_ms = float(step_ms.mean())
        achieved = alg_bytes / (mean_ms * 1e-3) / 1e9
        line = {
            "metric": "track-updates/s", "value": tu_all / (total_ms_max * 1e-3), "unit": "track-updates/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "config5: ByteTrack multi-stream (bytetrack.yaml: track_thresh 0.5, match_thresh 0.8, "
                                   "track_buffer 30), 200 objects/stream, sharded by stream",
                       "streams_per_gpu": S, "streams_total": S * world, "objects_per_stream": N_OBJECTS,
                       "max_tracks": MAX_TRACKS, "max_dets": MAX_DETS,
                       "l2": "per-step working set (state + detections + outputs) is ~%.0f MB > 126 MB L2; "
                             "every step reads new detections" % ((S * MAX_TRACKS * 200 * 2 + S * 200 * 48 + S * MAX_TRACKS * 64) / 1e6),
                       "data_gen_s": round(t_gen, 1)},
            "p50_step_ms": float(np.percentile(step_ms, 50)), "p99_step_ms": float(np.percentile(step_ms, 99)),
            "e2e": {"value": tu_e2e_all / (e2e_ms_max * 1e-3), "unit": "track-updates/s",
                    "h2d_bytes_per_step": h2d // args.steps, "d2h_bytes_per_step": d2h // args.steps,
                    "ms_per_step": e2e_ms_max / args.steps, "pipeline_depth": nslot,
                    "output_rows_per_step": rows_all / args.steps},
            "gpu_launches": int(launches_all),
            "roofline": {"bound": "hbm", "kernel": "bytetrack_step_kernel", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650",
                         "alg_bytes_per_launch": alg_bytes, "mean_launch_ms": mean_ms,
                         "bytes_per_track_update": alg_bytes * args.steps / max(1, tu_dev)},
            "cpu_baseline": cpu_base,
            "clocks": sampler.summary(),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--streams", type=int, default=STREAMS_PER_GPU, help="streams per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
