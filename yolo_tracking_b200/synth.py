"""Synthetic MOT-shaped detection streams.

Every stream is generated from its own ``numpy.random.default_rng(1000 * config + stream)``
so the oracle, the GPU parity tests and ``bench.py`` all see the same arrays.  All values are
continuous draws (no rounding) so the inputs are tie-free: no two association costs coincide
and no confidence sits exactly on a threshold.

Scene model: ``n_objects`` constant-velocity boxes on a 1920x1080 canvas (3840x2160 above 64
objects); a detection is the true box plus N(0, 1 px) noise per coordinate; each object is
missed with probability 0.05 per frame; Poisson(1) false positives per frame; detection
rows are shuffled per frame.  Optional ``occlusion`` drops an object for a run of 2-10
frames (exercises the OCSORT re-update path); optional ``emb_dim`` attaches a noisy copy
of a per-object prototype embedding to every detection (BoTSORT configs).
"""
from __future__ import annotations

import numpy as np

__all__ = ["make_stream", "make_batch", "stream_seed"]


def stream_seed(config: int, stream: int) -> int:
    return 1000 * int(config) + int(stream)


def make_stream(config: int, stream: int, n_objects: int, n_frames: int, *,
                dmax: int | None = None, miss_prob: float = 0.05, fp_rate: float = 1.0,
                occlusion: bool = False, emb_dim: int = 0):
    """Return ``(dets[F, dmax, 6] f64, ndets[F] i32, embs[F, dmax, emb_dim] f32 | None)``.

    Rows past ``ndets[f]`` are zero.  Columns are ``x1, y1, x2, y2, conf, cls``.
    """
    rng = np.random.default_rng(stream_seed(config, stream))
    N, F = int(n_objects), int(n_frames)
    W, H = (1920.0, 1080.0) if N <= 64 else (3840.0, 2160.0)
    margin = 100.0
    cx = rng.uniform(margin, W - margin, N)
    cy = rng.uniform(margin, H - margin, N)
    bw = rng.uniform(30.0, 90.0, N)
    bh = rng.uniform(60.0, 220.0, N)
    vx = rng.normal(0.0, 2.0, N)
    vy = rng.normal(0.0, 2.0, N)
    t = np.arange(F, dtype=np.float64)[:, None]
    px = cx[None, :] + vx[None, :] * t
    py = cy[None, :] + vy[None, :] * t
    noise = rng.normal(0.0, 1.0, (F, N, 4))
    boxes = np.stack([px - bw / 2, py - bh / 2, px + bw / 2, py + bh / 2], axis=-1) + noise
    seen = rng.random((F, N)) >= miss_prob
    hi = rng.random((F, N)) < 0.85
    conf = np.where(hi, rng.uniform(0.5, 0.99, (F, N)), rng.uniform(0.1, 0.5, (F, N)))
    if occlusion:
        start = rng.random((F, N)) < 0.02
        length = rng.integers(2, 11, (F, N))
        hidden = np.zeros((F, N), dtype=bool)
        for f, n in zip(*np.nonzero(start)):
            hidden[f:f + length[f, n], n] = True
        seen &= ~hidden
    n_fp = rng.poisson(fp_rate, F)
    max_fp = int(n_fp.max()) if F else 0
    fcx = rng.uniform(margin, W - margin, (F, max_fp))
    fcy = rng.uniform(margin, H - margin, (F, max_fp))
    fw = rng.uniform(30.0, 90.0, (F, max_fp))
    fh = rng.uniform(60.0, 220.0, (F, max_fp))
    fconf = rng.uniform(0.1, 0.6, (F, max_fp))
    keys = rng.random((F, N + max_fp))

    # all candidate rows [F, N + max_fp, 6]; invalid ones are sorted to the back
    cand = np.zeros((F, N + max_fp, 6))
    cand[:, :N, :4] = boxes
    cand[:, :N, 4] = conf
    cand[:, N:, 0] = fcx - fw / 2
    cand[:, N:, 1] = fcy - fh / 2
    cand[:, N:, 2] = fcx + fw / 2
    cand[:, N:, 3] = fcy + fh / 2
    cand[:, N:, 4] = fconf
    valid = np.concatenate([seen, np.arange(max_fp)[None, :] < n_fp[:, None]], axis=1)
    order = np.argsort(np.where(valid, keys, 2.0), axis=1, kind="stable")
    cand = np.take_along_axis(cand, order[:, :, None], axis=1)
    counts = valid.sum(1)
    cand *= (np.arange(N + max_fp)[None, :] < counts[:, None])[:, :, None]

    cap = int(dmax) if dmax is not None else (int(counts.max()) if F else 0)
    if F and counts.max() > cap:
        raise ValueError(f"stream {stream}: {int(counts.max())} detections exceed dmax={cap}")
    dets = np.zeros((F, cap, 6), dtype=np.float64)
    k = min(cap, N + max_fp)
    dets[:, :k] = cand[:, :k]
    ndets = counts.astype(np.int32)
    embs = None
    if emb_dim:
        proto = rng.standard_normal((N, emb_dim))
        e = np.empty((F, N + max_fp, emb_dim), dtype=np.float32)
        for f in range(F):                      # frame at a time: bounds the fp64 temporaries
            src = np.concatenate([proto, rng.standard_normal((max_fp, emb_dim))], axis=0)
            ef = src + 0.3 * rng.standard_normal(src.shape)
            e[f] = ef[order[f]].astype(np.float32)
        e *= (np.arange(N + max_fp)[None, :] < counts[:, None])[:, :, None]
        embs = np.zeros((F, cap, emb_dim), dtype=np.float32)
        embs[:, :k] = e[:, :k]
    return dets, ndets, embs


def make_batch(config: int, n_streams: int, n_objects: int, n_frames: int, *, dmax: int,
               first_stream: int = 0, **kw):
    """Stack ``n_streams`` streams: ``dets[F, S, dmax, 6]``, ``ndets[F, S]``, ``embs`` or None."""
    dets = np.zeros((n_frames, n_streams, dmax, 6), dtype=np.float64)
    ndets = np.zeros((n_frames, n_streams), dtype=np.int32)
    embs = None
    for i in range(n_streams):
        d, n, e = make_stream(config, first_stream + i, n_objects, n_frames, dmax=dmax, **kw)
        dets[:, i] = d
        ndets[:, i] = n
        if e is not None:
            if embs is None:
                embs = np.zeros((n_frames, n_streams, dmax, e.shape[-1]), dtype=np.float32)
            embs[:, i] = e
    return dets, ndets, embs
