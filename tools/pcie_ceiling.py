#!/usr/bin/env python
"""Aggregate pinned host<->device copy ceiling of one box with N GPUs copying at once.

  python tools/pcie_ceiling.py [--gpus 1,2,4,8] [--in-mb 19.7] [--out-mb 31.5] [--iters 40]

One process per GPU; every process allocates pinned host buffers of the bench's per-step sizes (packed interface:
~19.7 MB in, ~31.5 MB out at 4096 streams x 200 objects) and issues plain linear cudaMemcpyAsync copies (torch's
non_blocking copy_ between a pinned and a device tensor) on two streams - H2D only, D2H only, and both directions at once.
All processes start together (barrier) and the wall time of the slowest one is taken.  Each configuration is run with
the process left where the OS put it and with the process (and hence its first-touch pinned pages) bound to the CPUs of
the GPU's NUMA node, when sysfs exposes one.  One JSON line per measurement on stdout.

bench.py's e2e leg at N GPUs is bounded by the `both` line at that N: e2e ms/step >= (in + out bytes per rank) * N / aggregate.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time


def numa_cpus_of_gpu(index):
    """CPUs of the NUMA node the GPU hangs off (sysfs), or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:          # nvml pads the domain to 8 hex digits, sysfs uses 4
            bus = bus[4:]
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None, node
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0)
        return sorted(cpus & allowed), node
    except Exception:
        return None, None


def worker(rank, world, in_bytes, out_bytes, iters, bind, barrier, q):
    node = None
    if bind:
        cpus, node = numa_cpus_of_gpu(rank)
        if cpus:
            os.sched_setaffinity(0, cpus)
    import torch
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    h_in = torch.empty(in_bytes, dtype=torch.uint8, pin_memory=True)
    h_in.fill_(1)                                  # first touch under the chosen affinity
    h_out = torch.empty(out_bytes, dtype=torch.uint8, pin_memory=True)
    h_out.fill_(1)
    d_in = torch.empty(in_bytes, dtype=torch.uint8, device=dev)
    d_out = torch.ones(out_bytes, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    res = {}
    for mode in ("h2d", "d2h", "both"):
        for timed in (False, True):
            torch.cuda.synchronize(dev)
            barrier.wait()
            t0 = time.perf_counter()
            for _ in range(iters if timed else 3):
                if mode in ("h2d", "both"):
                    with torch.cuda.stream(s1):
                        d_in.copy_(h_in, non_blocking=True)
                if mode in ("d2h", "both"):
                    with torch.cuda.stream(s2):
                        h_out.copy_(d_out, non_blocking=True)
            torch.cuda.synchronize(dev)
            dt = time.perf_counter() - t0
            barrier.wait()
        res[mode] = dt
    q.put((rank, res, node))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", default="1,2,4,8")
    ap.add_argument("--in-mb", type=float, default=19.7)
    ap.add_argument("--out-mb", type=float, default=31.5)
    ap.add_argument("--iters", type=int, default=40)
    args = ap.parse_args()
    import torch
    import torch.multiprocessing as mp
    have = torch.cuda.device_count()
    in_bytes, out_bytes = int(args.in_mb * 1e6), int(args.out_mb * 1e6)
    ctx = mp.get_context("spawn")
    print(json.dumps({"host_cpus": len(os.sched_getaffinity(0)), "gpus_visible": have,
                      "numa": {str(i): numa_cpus_of_gpu(i)[1] for i in range(have)}}), flush=True)
    for n in [int(x) for x in args.gpus.split(",")]:
        if n > have:
            continue
        for bind in (False, True):
            barrier = ctx.Barrier(n)
            q = ctx.Queue()
            procs = [ctx.Process(target=worker, args=(r, n, in_bytes, out_bytes, args.iters, bind, barrier, q)) for r in range(n)]
            for p in procs:
                p.start()
            out = [q.get() for _ in range(n)]
            for p in procs:
                p.join()
            line = {"n_gpus": n, "numa_bound": bind, "in_bytes": in_bytes, "out_bytes": out_bytes, "iters": args.iters}
            for mode, nbytes in (("h2d", in_bytes), ("d2h", out_bytes), ("both", in_bytes + out_bytes)):
                worst = max(r[1][mode] for r in out)
                line[mode + "_ms_per_iter"] = worst / args.iters * 1e3
                line[mode + "_aggregate_gbs"] = n * nbytes * args.iters / worst / 1e9
            line["numa_nodes"] = [r[2] for r in sorted(out)]
            print(json.dumps(line), flush=True)


if __name__ == "__main__":
    sys.exit(main())
