"""Many independent video streams tracked at once by one device context.

This is the throughput API; the reference-shaped single-stream objects in
``yolo_tracking_b200.trackers`` are batches of one built on the same calls.  The frame step
itself is the C-ABI ``b200track_step*`` (include/b200track.h); numpy / torch are only used to
hand buffers over.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

_KINDS = {"bytetrack": _lib.BYTETRACK, "ocsort": _lib.OCSORT, "botsort": _lib.BOTSORT, "deepocsort": _lib.DEEPOCSORT,
          "strongsort": _lib.STRONGSORT, "hybridsort": _lib.HYBRIDSORT}


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    return C.c_void_p(a.data_ptr())          # torch tensor


class BatchedTracker:
    """``n_streams`` trackers of one kind with identical hyper-parameters.

    ``params`` are the reference constructor arguments (tracker_zoo.py:43-81), e.g. for
    ByteTrack ``track_thresh, match_thresh, track_buffer, frame_rate``.
    """

    def __init__(self, kind: str, n_streams: int, max_tracks: int = 256, max_dets: int = 256,
                 device: int = 0, feat_dim: int = 0, **params):
        if kind not in _KINDS:
            raise ValueError(f"No such tracker: {kind}")
        self.kind = kind
        self.n_streams, self.max_tracks, self.max_dets, self.feat_dim = n_streams, max_tracks, max_dets, feat_dim
        self.device = device
        cfg = _lib.Config()
        cfg.kind = _KINDS[kind]
        cfg.n_streams, cfg.max_tracks, cfg.max_dets, cfg.feat_dim, cfg.device = n_streams, max_tracks, max_dets, feat_dim, device
        if kind == "bytetrack":
            # BYTETracker defaults (byte_tracker.py:115-117)
            tt = params.get("track_thresh", 0.45)
            cfg.track_thresh, cfg.track_low_thresh, cfg.new_track_thresh = tt, 0.1, tt
            cfg.match_thresh = params.get("match_thresh", 0.8)
            cfg.track_buffer = int(params.get("track_buffer", 25))
            cfg.frame_rate = int(params.get("frame_rate", 30))
        elif kind == "botsort":
            # BoTSORT defaults (bot_sort.py:185-201)
            cfg.track_thresh = params.get("track_high_thresh", 0.5)
            cfg.track_low_thresh = params.get("track_low_thresh", 0.1)
            cfg.new_track_thresh = params.get("new_track_thresh", 0.6)
            cfg.match_thresh = params.get("match_thresh", 0.8)
            cfg.proximity_thresh = params.get("proximity_thresh", 0.5)
            cfg.appearance_thresh = params.get("appearance_thresh", 0.25)
            cfg.track_buffer = int(params.get("track_buffer", 30))
            cfg.frame_rate = int(params.get("frame_rate", 30))
            cfg.with_reid = int(params.get("with_reid", True))
            cfg.fuse_first_associate = int(bool(params.get("fuse_first_associate", False)))
            cfg.camera_motion = int(bool(params.get("camera_motion", False)))
        elif kind == "deepocsort":
            # DeepOCSort defaults (deep_ocsort.py:309-330)
            cfg.det_thresh = params.get("det_thresh", 0.3)
            cfg.max_age = int(params.get("max_age", 30))
            cfg.min_hits = int(params.get("min_hits", 3))
            cfg.iou_thresh = params.get("iou_threshold", params.get("iou_thresh", 0.3))
            cfg.delta_t = int(params.get("delta_t", 3))
            cfg.asso_func = _lib.SIM[params.get("asso_func", "iou")]
            cfg.inertia = params.get("inertia", 0.2)
            cfg.w_association_emb = params.get("w_association_emb", 0.5)
            cfg.alpha_fixed_emb = params.get("alpha_fixed_emb", 0.95)
            cfg.aw_param = params.get("aw_param", 0.5)
            cfg.embedding_off = int(bool(params.get("embedding_off", False)))
            cfg.aw_off = int(bool(params.get("aw_off", False)))
            if cfg.embedding_off:
                self.feat_dim = cfg.feat_dim = 0
        elif kind == "hybridsort":
            # HybridSORT defaults (hybridsort.py:337-338); everything else is fixed by its constructor
            cfg.det_thresh = params["det_thresh"]
            cfg.max_age = int(params.get("max_age", 30))
            cfg.min_hits = int(params.get("min_hits", 3))
            cfg.iou_thresh = params.get("iou_threshold", params.get("iou_thresh", 0.3))
            cfg.delta_t = int(params.get("delta_t", 3))
            cfg.asso_func = _lib.SIM[params.get("asso_func", "iou")]
            cfg.inertia = params.get("inertia", 0.2)
            cfg.use_byte = int(bool(params.get("use_byte", False)))
        elif kind == "strongsort":
            # StrongSORT defaults (strong_sort.py:14-25)
            cfg.max_dist = params.get("max_dist", 0.2)
            cfg.max_iou_dist = params.get("max_iou_dist", 0.7)
            cfg.max_age = int(params.get("max_age", 30))
            cfg.n_init = int(params.get("n_init", 1))
            cfg.nn_budget = int(params.get("nn_budget", 100))
            cfg.mc_lambda = params.get("mc_lambda", 0.995)
            cfg.ema_alpha = params.get("ema_alpha", 0.9)
        else:
            # OCSort defaults (ocsort.py:191-203)
            cfg.det_thresh = params.get("det_thresh", 0.2)
            cfg.max_age = int(params.get("max_age", 30))
            cfg.min_hits = int(params.get("min_hits", 3))
            cfg.iou_thresh = params.get("asso_threshold", params.get("iou_thresh", 0.3))
            cfg.delta_t = int(params.get("delta_t", 3))
            cfg.asso_func = _lib.SIM[params.get("asso_func", "iou")]
            cfg.inertia = params.get("inertia", 0.2)
            cfg.use_byte = int(params.get("use_byte", False))
        self._cfg = cfg
        self._lib = _lib.load()
        self._ctx = C.c_void_p()
        _lib.check(self._lib.b200track_create(C.byref(cfg), C.byref(self._ctx)))

    # ------------------------------------------------------------------ lifetime
    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx:
            self._lib.b200track_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self):
        _lib.check(self._lib.b200track_reset(self._ctx))

    # ------------------------------------------------------------------ frame step
    def step_device(self, d_dets, d_ndets, d_out, d_nout, d_feats=None, img_hw=(0, 0), stream=None, d_warps=None):
        """Asynchronous step on device tensors (torch, contiguous): dets [S, max_dets, 6] f64,
        ndets [S] i32 -> out [S, max_tracks, 8] f64, nout [S] i32; d_warps [S, 6] f64 camera motion (optional)."""
        st = C.c_void_p(stream) if stream else None
        if d_warps is not None:
            _lib.check(self._lib.b200track_step_cam(self._ctx, _ptr(d_dets), _ptr(d_ndets), _ptr(d_feats), _ptr(d_warps),
                                                    int(img_hw[0]), int(img_hw[1]), _ptr(d_out), _ptr(d_nout), st))
            return
        _lib.check(self._lib.b200track_step(self._ctx, _ptr(d_dets), _ptr(d_ndets), _ptr(d_feats),
                                            int(img_hw[0]), int(img_hw[1]), _ptr(d_out), _ptr(d_nout), st))

    def _host_args(self, dets, ndets, feats, out, nout):
        S, D, T = self.n_streams, self.max_dets, self.max_tracks
        assert dets.dtype == np.float64 and dets.shape == (S, D, 6) and dets.flags.c_contiguous
        assert ndets.dtype == np.int32 and ndets.shape == (S,) and ndets.flags.c_contiguous
        assert out.dtype == np.float64 and out.shape == (S, T, 8) and out.flags.c_contiguous
        assert nout.dtype == np.int32 and nout.shape == (S,) and nout.flags.c_contiguous
        if feats is not None:
            assert feats.dtype == np.float32 and feats.shape == (S, D, self.feat_dim) and feats.flags.c_contiguous

    def update_batch(self, dets, ndets, feats=None, img_hw=(0, 0), out=None, nout=None):
        """Synchronous step on HOST arrays (copies in, steps, copies out)."""
        if out is None:
            out = np.empty((self.n_streams, self.max_tracks, 8), dtype=np.float64)
        if nout is None:
            nout = np.empty((self.n_streams,), dtype=np.int32)
        self._host_args(dets, ndets, feats, out, nout)
        _lib.check(self._lib.b200track_step_host(self._ctx, _ptr(dets), _ptr(ndets), _ptr(feats),
                                                 int(img_hw[0]), int(img_hw[1]), _ptr(out), _ptr(nout)))
        return out, nout

    @property
    def host_slots(self):
        return self._lib.b200track_host_slots(self._ctx)

    def submit(self, slot, dets, ndets, out, nout, feats=None, img_hw=(0, 0)):
        """Pipelined host step: returns immediately; ``wait(slot)`` before reading out/nout."""
        _lib.check(self._lib.b200track_submit_host(self._ctx, slot, _ptr(dets), _ptr(ndets), _ptr(feats),
                                                   int(img_hw[0]), int(img_hw[1]), _ptr(out), _ptr(nout)))

    def wait(self, slot):
        _lib.check(self._lib.b200track_wait_host(self._ctx, slot))

    def sync(self):
        _lib.check(self._lib.b200track_sync(self._ctx))

    # ------------------------------------------------------------------ packed frames (one copy per direction)
    _ROW_DTYPES = {
        "bytetrack": np.dtype([("box", "<f8", 4), ("id", "<i4"), ("det_ind", "<i4")]),
        "botsort": np.dtype([("box", "<f8", 4), ("id", "<i4"), ("det_ind", "<i4"), ("cls", "<f4"), ("conf", "<f4")]),
        "ocsort": np.dtype([("id", "<i4"), ("det_ind", "<i4")]),
    }
    _ROW_DTYPES["deepocsort"] = _ROW_DTYPES["bytetrack"]

    _EXC_DTYPE = np.dtype([("row", "<i4"), ("reserved", "<i4"), ("box", "<f8", 4)])

    @property
    def _has_feats(self):
        return (self.kind == "botsort" and bool(self._cfg.with_reid)) or (self.kind == "deepocsort" and not self._cfg.embedding_off)

    @property
    def _has_warps(self):
        return self.kind in ("botsort", "deepocsort")

    def frame_layout(self, n_rows: int, dtype=np.float32):
        L = _lib.Layout()
        _lib.check(self._lib.b200track_frame_layout(self._ctx, int(n_rows), _lib.F32 if np.dtype(dtype) == np.float32 else _lib.F64,
                                                    C.byref(L)))
        return L

    def frame_buffers(self, max_rows: int | None = None, pinned: bool = True):
        """(input block, result block) as uint8 arrays big enough for ``max_rows`` detection rows in either dtype."""
        L = self.frame_layout(self.n_streams * self.max_dets if max_rows is None else max_rows, np.float64)
        if pinned:
            import torch
            keep = (torch.empty(int(L.in_bytes), dtype=torch.uint8, pin_memory=True),
                    torch.empty(int(L.out_bytes), dtype=torch.uint8, pin_memory=True))
            bufs = tuple(t.numpy() for t in keep)
            self._pinned = getattr(self, "_pinned", []) + [keep]          # the arrays borrow the tensors' memory
            return bufs
        return np.empty(int(L.in_bytes), dtype=np.uint8), np.empty(int(L.out_bytes), dtype=np.uint8)

    def pack(self, block, dets, ndets=None, feats=None, warps=None, dtype=np.float32):
        """Fill an input block.  ``dets``: a list of per-stream [n_s, 6] arrays, or a padded [S, D, 6] array with
        ``ndets`` [S].  ``feats`` likewise ([n_s, F] per stream or [S, D, F]); ``warps`` [S, 2, 3].  Returns
        ``(n_rows, flags)``; the views ``offsets``, ``dets`` of the block come from ``frame_views``."""
        S = self.n_streams
        dtype = np.dtype(dtype)
        if ndets is None:
            counts = np.fromiter((len(d) for d in dets), dtype=np.int64, count=S)
        else:
            counts = np.asarray(ndets, dtype=np.int64)
        if counts.max(initial=0) > self.max_dets:
            raise ValueError(f"{int(counts.max())} detections exceed max_dets={self.max_dets}")
        off = np.zeros(S + 1, dtype=np.int32)
        np.cumsum(counts, out=off[1:])
        R = int(off[S])
        v = self.frame_views(block, None, R, dtype)
        v["offsets"][:] = off
        if ndets is None:
            if R:
                np.concatenate([np.asarray(d).reshape(-1, 6) for d in dets], axis=0, out=v["dets"], casting="same_kind")
            if feats is not None and v["feats"] is not None and R:
                np.concatenate([np.asarray(f).reshape(-1, self.feat_dim) for f in feats], axis=0, out=v["feats"], casting="same_kind")
        else:
            mask = np.arange(self.max_dets)[None, :] < counts[:, None]
            v["dets"][:] = np.asarray(dets)[mask]
            if feats is not None and v["feats"] is not None:
                v["feats"][:] = np.asarray(feats)[mask]
        flags = 0
        if warps is not None:
            if v["warps"] is None:
                raise ValueError("only BoT-SORT and DeepOCSORT contexts take camera-motion warps")
            v["warps"][:] = np.asarray(warps, dtype=np.float64).reshape(S, 6)
            flags |= _lib.FRAME_HAS_WARPS
        return R, flags

    def frame_views(self, in_block, out_block, n_rows: int, dtype=np.float32):
        """Typed views into the blocks of a frame with ``n_rows`` detection rows."""
        dtype = np.dtype(dtype)
        L = self.frame_layout(n_rows, dtype)
        S, R = self.n_streams, int(n_rows)
        v = {"layout": L}
        if in_block is not None:
            v["offsets"] = in_block[L.in_off_offsets:L.in_off_offsets + 4 * (S + 1)].view(np.int32)
            v["dets"] = in_block[L.in_off_dets:L.in_off_dets + R * 6 * dtype.itemsize].view(dtype).reshape(R, 6)
            v["warps"] = in_block[L.in_off_warps:L.in_off_warps + 48 * S].view(np.float64).reshape(S, 6) if self._has_warps else None
            v["feats"] = (in_block[L.in_off_feats:L.in_off_feats + R * self.feat_dim * 4].view(np.float32).reshape(R, self.feat_dim)
                          if self._has_feats else None)
        if out_block is not None:
            v["header"] = out_block[0:16].view(np.int32)
            v["nout"] = out_block[L.out_off_nout:L.out_off_nout + 4 * S].view(np.int32)
            v["rows"] = out_block[L.out_off_rows:L.out_off_rows + R * L.row_bytes].view(self._ROW_DTYPES[self.kind])
            v["exc"] = out_block[L.out_off_exc:L.out_off_exc + L.exc_capacity * 40].view(self._EXC_DTYPE)
        return v

    def submit_packed(self, slot, in_block, out_block, dtype=np.float32, flags=0, img_hw=(0, 0)):
        _lib.check(self._lib.b200track_submit_packed(self._ctx, slot, _ptr(in_block), _lib.F32 if np.dtype(dtype) == np.float32 else _lib.F64,
                                                     int(flags), int(img_hw[0]), int(img_hw[1]), _ptr(out_block)))

    def wait_packed(self, slot):
        _lib.check(self._lib.b200track_wait_packed(self._ctx, slot))

    def step_packed_device(self, d_in, n_rows, d_result, dtype=np.float32, flags=0, img_hw=(0, 0), stream=None):
        """Asynchronous step on DEVICE blocks (torch uint8 tensors laid out like the host blocks)."""
        st = C.c_void_p(stream) if stream else None
        _lib.check(self._lib.b200track_step_packed(self._ctx, _ptr(d_in), int(n_rows), _lib.F32 if np.dtype(dtype) == np.float32 else _lib.F64,
                                                   int(flags), int(img_hw[0]), int(img_hw[1]), _ptr(d_result), st))

    def expand(self, in_block, out_block, n_rows: int, dtype=np.float32):
        """The reference's result arrays rebuilt from a finished frame: ``(rows[M, 8] f64, stream_of_row[M])`` with
        rows = [x1, y1, x2, y2, id, conf, cls, det_ind] in stream order, each stream in the reference's row order."""
        v = self.frame_views(in_block, out_block, n_rows, dtype)
        off, nout, rows, dets = v["offsets"], v["nout"], v["rows"], v["dets"]
        S = self.n_streams
        stream_of = np.repeat(np.arange(S, dtype=np.int32), nout)
        first = np.repeat(off[:S].astype(np.int64), nout)
        within = np.arange(len(stream_of), dtype=np.int64) - np.repeat(np.cumsum(nout, dtype=np.int64) - nout, nout)
        r = rows[first + within]
        di = r["det_ind"] & ~(_lib.ROW_OC_NEW | _lib.ROW_OC_STATE)
        src = dets[first + di].astype(np.float64, copy=False)
        out = np.empty((len(r), 8), dtype=np.float64)
        if self.kind == "ocsort":
            out[:, 0:4] = src[:, 0:4]
            new = (r["det_ind"] & _lib.ROW_OC_NEW) != 0
            if new.any():                               # KalmanBoxTracker.get_state of a fresh tracker (ocsort.py:24-62)
                b = src[new, 0:4]
                w, h = b[:, 2] - b[:, 0], b[:, 3] - b[:, 1]
                x, y, s_, r_ = b[:, 0] + w / 2.0, b[:, 1] + h / 2.0, w * h, w / (h + 1e-6)
                w2 = np.sqrt(s_ * r_)
                h2 = s_ / w2
                out[new, 0], out[new, 1], out[new, 2], out[new, 3] = x - w2 / 2.0, y - h2 / 2.0, x + w2 / 2.0, y + h2 / 2.0
            state = np.nonzero(r["det_ind"] & _lib.ROW_OC_STATE)[0]
            if len(state):                              # rows that report the filter's box (ocsort.py:355-358): exception area
                exc = v["exc"][:int(v["header"][1])]
                where = {int(e["row"]): e["box"] for e in exc}
                for k in state:
                    out[k, 0:4] = where[int(first[k] + within[k])]
        else:
            out[:, 0:4] = r["box"]
        out[:, 4] = r["id"]
        out[:, 5] = src[:, 4]
        out[:, 6] = r["cls"] if self.kind == "botsort" else src[:, 5]
        out[:, 7] = di
        return out, stream_of

    def update_frames(self, dets, feats=None, warps=None, img_hw=(0, 0), dtype=None):
        """Synchronous step on a list of per-stream detection arrays -> list of per-stream result arrays [M_s, 8]
        (the reference's ``tracker.update`` output for every stream).  fp32 detections travel as fp32."""
        if dtype is None:
            dtype = np.float32 if all(np.asarray(d).dtype == np.float32 for d in dets) else np.float64
        if not hasattr(self, "_frame_bufs"):
            self._frame_bufs = self.frame_buffers()
        bi, bo = self._frame_bufs
        R, flags = self.pack(bi, dets, feats=feats, warps=warps, dtype=dtype)
        self.submit_packed(0, bi, bo, dtype, flags, img_hw)
        self.wait_packed(0)
        out, stream_of = self.expand(bi, bo, R, dtype)
        cuts = np.cumsum(self.frame_views(None, bo, R, dtype)["nout"])[:-1]
        return np.split(out, cuts)

    # ------------------------------------------------------------------ probes
    def track_updates(self) -> int:
        v = C.c_uint64()
        _lib.check(self._lib.b200track_track_updates(self._ctx, C.byref(v)))
        return v.value

    def counters(self):
        """Event counters over all streams (DeepOCSORT): assignment-solved first associations, recovery rounds, re-updates."""
        buf = (C.c_uint64 * 8)()
        _lib.check(self._lib.b200track_counters(self._ctx, C.byref(buf)))
        return dict(lap_frames=int(buf[0]), ocr_frames=int(buf[1]), oru=int(buf[2]), gallery_rows=int(buf[3]))

    def launches(self) -> int:
        v = C.c_uint64()
        _lib.check(self._lib.b200track_launch_count(self._ctx, C.byref(v)))
        return v.value

    def phase_cycles(self, reset=True):
        """Per-phase cycle counters of the step kernel (first call enables them)."""
        buf = (C.c_uint64 * 16)()
        _lib.check(self._lib.b200track_phase_cycles(self._ctx, C.byref(buf), int(reset)))
        return list(buf)

    def footprint(self):
        a, b = C.c_uint64(), C.c_uint64()
        _lib.check(self._lib.b200track_footprint(self._ctx, C.byref(a), C.byref(b)))
        return dict(state_bytes_per_stream=a.value, smem_bytes=b.value)

    def live_classes(self, stream: int = 0):
        """HybridSORT contexts: the classes of the live trackers of one stream in list order (what the reference's
        PerClassDecorator reads before every frame) - one small state read, no embeddings."""
        T = self.max_tracks
        counts = np.zeros(4, dtype=np.int32)
        aux = np.zeros((T, 3))
        _lib.check(self._lib.b200track_get_state_hybridsort(self._ctx, int(stream), _ptr(counts), None, None, None, None, None, _ptr(aux)))
        return aux[:int(counts[0]), 1].copy()

    def state(self, stream: int = 0):
        """Track records of one stream in list order, in the reference's dense form."""
        T = self.max_tracks
        counts = np.zeros(4, dtype=np.int32)
        rec = np.zeros((T, 6), dtype=np.int32)
        if self.kind == "hybridsort":
            x, P, vel, last, aux = np.zeros((T, 9)), np.zeros((T, 81)), np.zeros((T, 8)), np.zeros((T, 5)), np.zeros((T, 3))
            feat = np.zeros((T, self.feat_dim), dtype=np.float32)
            _lib.check(self._lib.b200track_get_state_hybridsort(self._ctx, int(stream), _ptr(counts), _ptr(rec), _ptr(x), _ptr(P),
                                                                _ptr(vel), _ptr(last), _ptr(aux)))
            _lib.check(self._lib.b200track_get_features(self._ctx, int(stream), _ptr(feat)))
            n = int(counts[0])
            return dict(n=n, id_count=int(counts[2]), frame_count=int(counts[3]),
                        track_id=rec[:n, 0].copy(), age=rec[:n, 1].copy(), time_since_update=rec[:n, 2].copy(),
                        hits=rec[:n, 3].copy(), hit_streak=rec[:n, 4].copy(), observed=rec[:n, 5].copy(),
                        x=x[:n].copy(), P=P[:n].reshape(n, 9, 9).copy(), velocity=vel[:n].reshape(n, 4, 2).copy(),
                        last_observation=last[:n].copy(), conf=aux[:n, 0].copy(), cls=aux[:n, 1].copy(),
                        det_ind=aux[:n, 2].copy(), smooth_feat=feat[:n].copy())
        mean = np.zeros((T, 8))
        cov = np.zeros((T, 64))
        aux = np.zeros((T, 3))
        _lib.check(self._lib.b200track_get_state(self._ctx, int(stream), _ptr(counts), _ptr(rec), _ptr(mean),
                                                 _ptr(cov), _ptr(aux)))
        n = int(counts[0] + counts[1])
        if self.kind == "strongsort":
            feat = np.zeros((T, self.feat_dim), dtype=np.float32)
            _lib.check(self._lib.b200track_get_features(self._ctx, int(stream), _ptr(feat)))
            return dict(n=n, next_id=int(counts[2]), frame_count=int(counts[3]),
                        track_id=rec[:n, 0].copy(), state=rec[:n, 1].copy(), hits=rec[:n, 2].copy(), age=rec[:n, 3].copy(),
                        time_since_update=rec[:n, 4].copy(), gallery=rec[:n, 5].copy(), mean=mean[:n].copy(),
                        cov=cov[:n].reshape(n, 8, 8).copy(), conf=aux[:n, 0].copy(), cls=aux[:n, 1].copy(),
                        det_ind=aux[:n, 2].copy(), feature=feat[:n].copy())
        if self.kind == "deepocsort":
            extra = np.zeros((T, 8))
            F = self.feat_dim
            emb = np.zeros((T, max(F, 1)))
            _lib.check(self._lib.b200track_get_track_extras(self._ctx, int(stream), _ptr(extra), _ptr(emb) if F else None))
            return dict(n=n, id_count=int(counts[2]), frame_count=int(counts[3]),
                        track_id=rec[:n, 0].copy(), age=rec[:n, 1].copy(), time_since_update=rec[:n, 2].copy(),
                        hits=rec[:n, 3].copy(), hit_streak=rec[:n, 4].copy(), observed=rec[:n, 5] & 1, frozen=(rec[:n, 5] >> 1) & 1,
                        x=mean[:n].copy(), P=cov[:n].reshape(n, 8, 8).copy(), velocity=extra[:n, 0:2].copy(),
                        last_observation=extra[:n, 2:7].copy(), conf=aux[:n, 0].copy(), cls=aux[:n, 1].copy(),
                        det_ind=aux[:n, 2].copy(), emb=emb[:n].copy() if F else np.ones((n, 1)))
        if self.kind == "ocsort":
            c = cov[:n]
            return dict(n=n, id_count=int(counts[2]), frame_count=int(counts[3]),
                        track_id=rec[:n, 0].copy(), age=rec[:n, 1].copy(), time_since_update=rec[:n, 2].copy(),
                        hits=rec[:n, 3].copy(), hit_streak=rec[:n, 4].copy(), observed=rec[:n, 5].copy(),
                        x=mean[:n, :7].copy(), P=c[:, :49].reshape(n, 7, 7).copy(), velocity=c[:, 49:51].copy(),
                        last_observation=c[:, 51:56].copy(), conf=aux[:n, 0].copy(), cls=aux[:n, 1].copy(),
                        det_ind=aux[:n, 2].copy())
        st = dict(n_tracked=int(counts[0]), n_lost=int(counts[1]), id_count=int(counts[2]), frame_id=int(counts[3]),
                  track_id=rec[:n, 0].copy(), state=rec[:n, 1].copy(), is_activated=rec[:n, 2].copy(),
                  frame_id_t=rec[:n, 3].copy(), start_frame=rec[:n, 4].copy(), tracklet_len=rec[:n, 5].copy(),
                  mean=mean[:n].copy(), cov=cov[:n].reshape(n, 8, 8).copy(), score=aux[:n, 0].copy(),
                  cls=aux[:n, 1].copy(), det_ind=aux[:n, 2].copy())
        if self.kind == "botsort" and self._cfg.with_reid:
            feat = np.zeros((T, self.feat_dim), dtype=np.float32)
            _lib.check(self._lib.b200track_get_features(self._ctx, int(stream), _ptr(feat)))
            st["smooth_feat"] = feat[:n].copy()
        return st
