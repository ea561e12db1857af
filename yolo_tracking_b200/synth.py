"""Synthetic MOT-shaped detection streams.

Every stream is generated from its own ``numpy.random.default_rng(1000 * config + stream)``
so the oracle, the golden fixtures, the GPU parity tests and ``bench.py`` all see the same
arrays.  All values are continuous draws (no rounding) so the inputs are tie-free: no two
association costs coincide and no confidence sits exactly on a threshold.

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
    shuffle_keys = rng.random((F, N + max_fp))
    proto = emb = None
    if emb_dim:
        proto = rng.standard_normal((N, emb_dim))
        fp_proto = rng.standard_normal((F, max_fp, emb_dim))
        emb_noise_seed = rng.integers(0, 2**31 - 1)

    counts = seen.sum(1) + n_fp
    cap = int(dmax) if dmax is not None else int(counts.max()) if F else 0
    if F and counts.max() > cap:
        raise ValueError(f"stream {stream}: {int(counts.max())} detections exceed dmax={cap}")
    dets = np.zeros((F, cap, 6), dtype=np.float64)
    ndets = counts.astype(np.int32)
    embs = np.zeros((F, cap, emb_dim), dtype=np.float32) if emb_dim else None
    if emb_dim:
        erng = np.random.default_rng(int(emb_noise_seed))
    for f in range(F):
        idx = np.nonzero(seen[f])[0]
        k = int(n_fp[f])
        rows = np.empty((len(idx) + k, 6))
        rows[:len(idx), :4] = boxes[f, idx]
        rows[:len(idx), 4] = conf[f, idx]
        rows[len(idx):, 0] = fcx[f, :k] - fw[f, :k] / 2
        rows[len(idx):, 1] = fcy[f, :k] - fh[f, :k] / 2
        rows[len(idx):, 2] = fcx[f, :k] + fw[f, :k] / 2
        rows[len(idx):, 3] = fcy[f, :k] + fh[f, :k] / 2
        rows[len(idx):, 4] = fconf[f, :k]
        rows[:, 5] = 0.0
        keys = np.concatenate([shuffle_keys[f, idx], shuffle_keys[f, N:N + k]])
        order = np.argsort(keys, kind="stable")
        dets[f, :len(rows)] = rows[order]
        if emb_dim:
            e = np.concatenate([proto[idx], fp_proto[f, :k]], axis=0)
            e = e + 0.3 * erng.standard_normal(e.shape)
            embs[f, :len(rows)] = e[order].astype(np.float32)
    return dets, ndets, embs


def make_batch(config: int, n_streams: int, n_objects: int, n_frames: int, *, dmax: int,
               first_stream: int = 0, **kw):
    """Stack ``n_streams`` streams: ``dets[F, S, dmax, 6]``, ``ndets[F, S]``, ``embs`` or None."""
    D, Nd, E = [], [], []
    for s in range(first_stream, first_stream + n_streams):
        d, n, e = make_stream(config, s, n_objects, n_frames, dmax=dmax, **kw)
        D.append(d)
        Nd.append(n)
        E.append(e)
    dets = np.ascontiguousarray(np.stack(D, axis=1))
    ndets = np.ascontiguousarray(np.stack(Nd, axis=1))
    embs = np.ascontiguousarray(np.stack(E, axis=1)) if E and E[0] is not None else None
    return dets, ndets, embs
