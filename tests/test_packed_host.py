"""CPU checks of the packed frame interface's host side: block layout arithmetic (mirrors b200track_frame_layout),
packing of per-stream detection lists, typed views, and the rebuild of the reference's [M, 8] result rows from compact
rows.  No device work: the context call is replaced by the same arithmetic in Python."""
import ctypes as C

import numpy as np
import pytest

from yolo_tracking_b200 import _lib
from yolo_tracking_b200.batch import BatchedTracker


class _HostOnly(BatchedTracker):
    def __init__(self, kind, S, D, F=0):
        self.kind, self.n_streams, self.max_dets, self.max_tracks, self.feat_dim = kind, S, D, D, F

        class Cfg:
            with_reid = 1
        self._cfg = Cfg()

    def frame_layout(self, n_rows, dtype=np.float32):           # csrc/api.cu: b200track_frame_layout
        L, S, R = _lib.Layout(), self.n_streams, int(n_rows)

        def a16(v):
            return (v + 15) & ~15
        det_row = 24 if np.dtype(dtype) == np.float32 else 48
        L.in_off_offsets = 0
        L.in_off_warps = a16(4 * (S + 1))
        L.in_off_dets = a16(L.in_off_warps + (48 * S if self.kind == "botsort" else 0))
        L.in_off_feats = a16(L.in_off_dets + det_row * R)
        L.in_bytes = a16(L.in_off_feats + (R * self.feat_dim * 4 if self.kind == "botsort" else 0))
        L.row_bytes = {"bytetrack": 40, "botsort": 48, "ocsort": 8}[self.kind]
        L.out_off_nout = 16
        L.out_off_rows = a16(16 + 4 * S)
        L.out_off_exc = a16(L.out_off_rows + R * L.row_bytes)
        L.exc_capacity = 64 + R // 128 if self.kind == "ocsort" else 0
        L.out_bytes = a16(L.out_off_exc + L.exc_capacity * 40)
        return L

    def __del__(self):
        pass


def test_row_structs_match_the_header():
    assert BatchedTracker._ROW_DTYPES["bytetrack"].itemsize == 40
    assert BatchedTracker._ROW_DTYPES["botsort"].itemsize == 48
    assert BatchedTracker._ROW_DTYPES["ocsort"].itemsize == 8
    assert C.sizeof(_lib.Layout) == 9 * 8 + 8
    with open(_lib.HEADER_PATH) as f:
        text = f.read()
    for needle in ("b200track_row;", "b200track_row_bot;", "b200track_row_oc;", "B200TRACK_FRAME_HAS_WARPS 1", "(1 << 30)"):
        assert needle in text


@pytest.mark.parametrize("kind", ["bytetrack", "ocsort", "botsort"])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_pack_views_expand_roundtrip(kind, dtype):
    rng = np.random.default_rng(3)
    S, D, F = 4, 8, 4 if kind == "botsort" else 0
    t = _HostOnly(kind, S, D, F)
    bi, bo = t.frame_buffers(pinned=False)
    counts = (2, 0, 5, 1)
    dets = [rng.uniform(0, 500, (n, 6)).astype(dtype) for n in counts]
    feats = [rng.standard_normal((n, F)).astype(np.float32) for n in counts] if F else None
    warps = rng.standard_normal((S, 2, 3)) if kind == "botsort" else None
    R, flags = t.pack(bi, dets, feats=feats, warps=warps, dtype=dtype)
    v = t.frame_views(bi, bo, R, dtype)
    assert R == 8 and list(v["offsets"]) == [0, 2, 2, 7, 8]
    assert v["dets"].dtype == dtype and np.array_equal(v["dets"], np.concatenate(dets))
    assert flags == (_lib.FRAME_HAS_WARPS if kind == "botsort" else 0)
    if F:
        assert np.array_equal(v["feats"], np.concatenate(feats)) and np.array_equal(v["warps"], warps.reshape(S, 6))
    # the padded form packs to the same block
    pad = np.zeros((S, D, 6), dtype=dtype)
    for s, d in enumerate(dets):
        pad[s, :len(d)] = d
    bi2, _ = t.frame_buffers(pinned=False)
    R2, _ = t.pack(bi2, pad, ndets=np.array(counts, dtype=np.int32), dtype=dtype)
    L = t.frame_layout(R, dtype)
    assert R2 == R and np.array_equal(bi2[L.in_off_dets:L.in_off_dets + R * 6 * np.dtype(dtype).itemsize],
                                      bi[L.in_off_dets:L.in_off_dets + R * 6 * np.dtype(dtype).itemsize])
    # a result block as the device would write it: stream 0 -> 1 row, stream 2 -> 2 rows, stream 3 -> 1 row
    v["header"][:] = 0
    v["nout"][:] = [1, 0, 2, 1]
    rows = v["rows"]
    picks = [(0, 0, 11, 1), (2, 2, 12, 4), (3, 2, 13, 0), (7, 3, 14, 0)]          # (row slot, stream, id, det_ind)
    for slot, s, tid, di in picks:
        rows["id"][slot] = tid
        rows["det_ind"][slot] = di | (_lib.ROW_OC_NEW if kind == "ocsort" and tid == 13 else 0)
        if kind != "ocsort":
            rows["box"][slot] = rng.uniform(0, 100, 4)
        if kind == "botsort":
            rows["cls"][slot] = 3.0
    out, stream_of = t.expand(bi, bo, R, dtype)
    assert out.shape == (4, 8) and list(stream_of) == [0, 2, 2, 3]
    assert list(out[:, 4]) == [11, 12, 13, 14] and list(out[:, 7]) == [1, 4, 0, 0]
    for k, (slot, s, tid, di) in enumerate(picks):
        src = dets[s][di].astype(np.float64)
        assert out[k, 5] == src[4]
        assert out[k, 6] == (3.0 if kind == "botsort" else src[5])
        if kind == "ocsort" and tid != 13:
            assert np.array_equal(out[k, :4], src[:4])
        elif kind == "ocsort":
            # a new tracker reports the detection's round trip through the filter state (ocsort.py:24-62)
            w, h = src[2] - src[0], src[3] - src[1]
            x, y, s_, r_ = src[0] + w / 2.0, src[1] + h / 2.0, w * h, w / (h + 1e-6)
            w2 = np.sqrt(s_ * r_)
            h2 = s_ / w2
            assert np.array_equal(out[k, :4], [x - w2 / 2.0, y - h2 / 2.0, x + w2 / 2.0, y + h2 / 2.0])
        else:
            assert np.array_equal(out[k, :4], rows["box"][slot])


def test_pack_rejects_more_detections_than_max_dets():
    t = _HostOnly("bytetrack", 2, 8)
    bi, _ = t.frame_buffers(pinned=False)
    with pytest.raises(ValueError):
        t.pack(bi, [np.zeros((9, 6)), np.zeros((0, 6))], dtype=np.float64)
