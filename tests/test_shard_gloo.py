"""The N>1 path on CPU: two gloo ranks each own a block of streams (no collective in the frame step),
the timing reduction and the optional output gather reproduce the single-process run."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from yolo_tracking_b200.shard import gather_outputs, owner_of, reduce_timing, shard_bounds

S_TOTAL, FRAMES, CAP = 6, 12, 64


def _run_block(lo, hi):
    """Oracle trackers stand in for the per-rank device context (CPU test): out[S_r, CAP, 8], nout[S_r]."""
    from oracle.bytetrack import ByteTrackOracle
    from yolo_tracking_b200.synth import make_stream
    out = np.zeros((hi - lo, CAP, 8))
    nout = np.zeros((hi - lo,), dtype=np.int32)
    units = 0
    for k, s in enumerate(range(lo, hi)):
        dets, nd, _ = make_stream(1, 500 + s, 12, FRAMES)
        trk = ByteTrackOracle(0.5, 0.8, 30, 30)
        for f in range(FRAMES):
            rows = trk.update(dets[f, :nd[f]]).reshape(-1, 8)
        out[k, :len(rows)] = rows
        nout[k] = len(rows)
        units += trk.track_updates
    return out, nout, units


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_bounds(S_TOTAL, rank, world)
    out, nout, units = _run_block(lo, hi)
    ms, (tot_units, tot_rows) = reduce_timing(10.0 + rank, [units, int(nout.sum())])
    g_out, g_nout = gather_outputs(torch.from_numpy(out), torch.from_numpy(nout))
    if rank == 0:
        q.put((ms, tot_units, tot_rows, g_out.numpy(), g_nout.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_bounds_cover_every_stream_once():
    for total in (1, 5, 8, 4096, 4099):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                lo, hi = shard_bounds(total, r, world)
                seen.extend(range(lo, hi))
                assert all(owner_of(s, total, world) == r for s in (lo, hi - 1) if lo < hi)
            assert seen == list(range(total))
            sizes = [shard_bounds(total, r, world) for r in range(world)]
            assert max(h - l for l, h in sizes) - min(h - l for l, h in sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def test_two_gloo_ranks_reproduce_single_process():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ms, tot_units, tot_rows, g_out, g_nout = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    out, nout, units = _run_block(0, S_TOTAL)
    assert ms == 11.0                                  # max over ranks
    assert tot_units == units and tot_rows == int(nout.sum())
    assert np.array_equal(g_nout, nout) and np.array_equal(g_out, out)


def test_single_process_reduction_is_identity():
    assert reduce_timing(3.5, [7, 9]) == (3.5, [7.0, 9.0])
