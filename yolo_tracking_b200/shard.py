"""Multi-GPU plumbing: streams shard by index, one process per GPU, no collective in the frame step.

The reference runs one tracker object per stream (examples/track.py:42-57) and one subprocess per
sequence (examples/val.py:189-226); streams never interact, so rank r of G simply owns the contiguous
block [r*S/G, (r+1)*S/G).  torch.distributed (NCCL on GPUs, gloo in the CPU tests) is only used for the
timing reduction of bench.py and for the optional final gather of the padded outputs."""
from __future__ import annotations


def shard_bounds(total: int, rank: int, world: int):
    """Contiguous block of stream indices owned by `rank`; sizes differ by at most one."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(int(total), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def owner_of(stream: int, total: int, world: int) -> int:
    base, extra = divmod(int(total), int(world))
    edge = extra * (base + 1)
    return stream // (base + 1) if stream < edge else extra + (stream - edge) // max(base, 1)


def reduce_timing(local_ms: float, local_counts, device=None):
    """(max over ranks of the device time, element-wise sum of the counters).  Single process: identity."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(local_ms)], dtype=torch.float64, device=device)
    c = torch.tensor([float(v) for v in local_counts], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
    return float(t.item()), c.tolist()


def gather_outputs(out, nout):
    """Optional final gather: every rank's padded out[S_r, max_tracks, 8] / nout[S_r] (equal S_r) concatenated
    in stream order on every rank.  One collective per tensor - the capacity padding makes shapes equal."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return out, nout
    world = dist.get_world_size()
    outs = [torch.empty_like(out) for _ in range(world)]
    nouts = [torch.empty_like(nout) for _ in range(world)]
    dist.all_gather(outs, out.contiguous())
    dist.all_gather(nouts, nout.contiguous())
    return torch.cat(outs, dim=0), torch.cat(nouts, dim=0)
