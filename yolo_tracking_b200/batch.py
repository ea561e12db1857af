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

_KINDS = {"bytetrack": _lib.BYTETRACK, "ocsort": _lib.OCSORT, "botsort": _lib.BOTSORT}


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
    def step_device(self, d_dets, d_ndets, d_out, d_nout, d_feats=None, img_hw=(0, 0), stream=None):
        """Asynchronous step on device tensors (torch, contiguous): dets [S, max_dets, 6] f64,
        ndets [S] i32 -> out [S, max_tracks, 8] f64, nout [S] i32."""
        st = C.c_void_p(stream) if stream else None
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

    # ------------------------------------------------------------------ probes
    def track_updates(self) -> int:
        v = C.c_uint64()
        _lib.check(self._lib.b200track_track_updates(self._ctx, C.byref(v)))
        return v.value

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

    def state(self, stream: int = 0):
        """Track records of one stream in list order, in the reference's dense form."""
        T = self.max_tracks
        counts = np.zeros(4, dtype=np.int32)
        rec = np.zeros((T, 6), dtype=np.int32)
        mean = np.zeros((T, 8))
        cov = np.zeros((T, 64))
        aux = np.zeros((T, 3))
        _lib.check(self._lib.b200track_get_state(self._ctx, int(stream), _ptr(counts), _ptr(rec), _ptr(mean),
                                                 _ptr(cov), _ptr(aux)))
        n = int(counts[0] + counts[1])
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
