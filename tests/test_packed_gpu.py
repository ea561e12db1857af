"""GPU parity of the packed frame interface (b200track_submit_packed / wait_packed / step_packed): one input block and
one result block per frame, compact result rows.  The results must be the padded interface's, bit for bit, and the
reference's goldens must replay through it."""
import numpy as np
import pytest

from _util import assert_close, load_golden

pytestmark = pytest.mark.gpu

BYTE = dict(track_thresh=0.5, match_thresh=0.8, track_buffer=30, frame_rate=30)
OC = dict(det_thresh=0.3, max_age=30, min_hits=1, asso_threshold=0.3, delta_t=3, asso_func="giou", inertia=0.2)
BOT = dict(track_high_thresh=0.5, track_low_thresh=0.1, new_track_thresh=0.6, track_buffer=30, match_thresh=0.8,
           proximity_thresh=0.5, appearance_thresh=0.25, frame_rate=30)


def _inputs(kind, S, N, F, D, fp32):
    from yolo_tracking_b200.synth import make_batch
    dets, nd, embs = make_batch(7, S, N, F, dmax=D, emb_dim=128 if kind == "botsort" else 0,
                                occlusion=(kind == "ocsort"), miss_prob=0.1, fp_rate=2.0)
    if fp32:
        dets = dets.astype(np.float32).astype(np.float64)
    feats = None if embs is None else np.ascontiguousarray(embs / 8.0)
    return dets, nd, feats


@pytest.mark.parametrize("kind,cfg", [("bytetrack", BYTE), ("ocsort", OC), ("botsort", BOT)])
@pytest.mark.parametrize("fp32", [False, True])
def test_packed_rows_equal_padded_rows(kind, cfg, fp32):
    from yolo_tracking_b200.batch import BatchedTracker
    S, N, F, D = 5, 30, 40, 64
    dets, nd, feats = _inputs(kind, S, N, F, D, fp32)
    fd = 128 if kind == "botsort" else 0
    T = 128 if kind == "ocsort" else D          # false positives live max_age frames as OC-SORT trackers
    a = BatchedTracker(kind, S, max_tracks=T, max_dets=D, feat_dim=fd, **cfg)
    b = BatchedTracker(kind, S, max_tracks=T, max_dets=D, feat_dim=fd, **cfg)
    hw = (1080, 1920)
    for f in range(F):
        out, nout = a.update_batch(np.ascontiguousarray(dets[f]), np.ascontiguousarray(nd[f]),
                                   feats=None if feats is None else np.ascontiguousarray(feats[f]), img_hw=hw)
        dl = [dets[f, s, :nd[f, s]].astype(np.float32 if fp32 else np.float64) for s in range(S)]
        fl = None if feats is None else [feats[f, s, :nd[f, s]] for s in range(S)]
        got = b.update_frames(dl, feats=fl, img_hw=hw)
        assert len(got) == S
        for s in range(S):
            ref = out[s, :nout[s]]
            assert got[s].shape == ref.shape, (f, s, got[s].shape, ref.shape)
            assert np.array_equal(got[s], ref), f"{kind} frame {f} stream {s}:\n{got[s] - ref}"
    assert a.track_updates() == b.track_updates()
    for s in range(S):                                   # same device state afterwards
        sa, sb = a.state(s), b.state(s)
        for k in sa:
            assert np.array_equal(np.asarray(sa[k]), np.asarray(sb[k])), (s, k)
    a.close(); b.close()


def test_packed_ocsort_rows_that_report_the_filter_box():
    """Boxes in negative coordinates: last_observation sums below zero, so a matched tracker reports the filter's box
    (ocsort.py:355-358) - it travels in the exception area of the result block."""
    from yolo_tracking_b200.batch import BatchedTracker
    S, N, F, D = 3, 20, 25, 64
    dets, nd, _ = _inputs("ocsort", S, N, F, D, False)
    dets = dets.copy()
    dets[..., 0:4] -= 5000.0
    dets[..., 0:4] *= (np.arange(D)[None, None, :, None] < nd[:, :, None, None])
    a = BatchedTracker("ocsort", S, max_tracks=128, max_dets=D, **OC)
    b = BatchedTracker("ocsort", S, max_tracks=128, max_dets=D, **OC)
    flagged = 0
    for f in range(F):
        out, nout = a.update_batch(np.ascontiguousarray(dets[f]), np.ascontiguousarray(nd[f]), img_hw=(1080, 1920))
        got = b.update_frames([dets[f, s, :nd[f, s]] for s in range(S)], img_hw=(1080, 1920))
        flagged += int(b.frame_views(None, b._frame_bufs[1], int(nd[f].sum()), np.float64)["header"][1])
        for s in range(S):
            assert np.array_equal(got[s], out[s, :nout[s]]), (f, s)
    assert flagged > 50
    a.close(); b.close()


def test_packed_bytetrack_replays_reference_golden():
    from yolo_tracking_b200.batch import BatchedTracker
    g = load_golden("bytetrack_churn")
    p = g["params"]
    dets, nd = g["dets"], g["ndets"]
    trk = BatchedTracker("bytetrack", 1, max_tracks=128, max_dets=128, track_thresh=p[0], match_thresh=p[1],
                         track_buffer=int(p[2]), frame_rate=int(p[3]))
    for f in range(dets.shape[0]):
        got = trk.update_frames([dets[f, :nd[f]]])[0]
        ref = g["out"][g["out_offs"][f]:g["out_offs"][f + 1]].reshape(-1, 8)
        assert got.shape == ref.shape, f
        assert np.array_equal(got[:, 4:], ref[:, 4:]), f"frame {f}: id/conf/cls/det_ind"
        assert_close(got[:, :4], ref[:, :4], what=f"frame {f} boxes")
    trk.close()


@pytest.mark.parametrize("name", ["ocsort_c2", "ocsort_churn", "ocsort_byte"])
def test_packed_ocsort_replays_reference_golden(name):
    from yolo_tracking_b200.batch import BatchedTracker
    g = load_golden(name)
    p = g["params"]
    dets, nd = g["dets"], g["ndets"]
    hw = tuple(int(v) for v in g["img_hw"])
    trk = BatchedTracker("ocsort", 1, max_tracks=128, max_dets=128, det_thresh=p[0], max_age=int(p[1]), min_hits=int(p[2]),
                         asso_threshold=p[3], delta_t=int(p[4]), asso_func="giou", inertia=p[5],
                         use_byte=bool(p[6]) if len(p) > 6 else False)
    for f in range(dets.shape[0]):
        got = trk.update_frames([dets[f, :nd[f]]], img_hw=hw)[0]
        ref = g["out"][g["out_offs"][f]:g["out_offs"][f + 1]].reshape(-1, 8)
        assert got.shape == ref.shape, f
        assert np.array_equal(got[:, 4:], ref[:, 4:]), f"{name} frame {f}"
        assert_close(got[:, :4], ref[:, :4], what=f"{name} frame {f} boxes")
    trk.close()


def test_packed_pipeline_three_slots_and_empty_streams():
    """Three frames in flight through submit_packed / wait_packed give the sequential results; streams without
    detections and a frame without any detection at all are fine."""
    from yolo_tracking_b200.batch import BatchedTracker
    S, N, F, D = 6, 25, 30, 64
    dets, nd, _ = _inputs("bytetrack", S, N, F, D, True)
    nd = nd.copy()
    nd[:, 2] = 0                     # a stream that never sees a detection
    nd[7] = 0                        # a frame without detections
    seq = BatchedTracker("bytetrack", S, max_tracks=D, max_dets=D, **BYTE)
    ref = [seq.update_frames([dets[f, s, :nd[f, s]].astype(np.float32) for s in range(S)]) for f in range(F)]
    pipe = BatchedTracker("bytetrack", S, max_tracks=D, max_dets=D, **BYTE)
    nslot = pipe.host_slots
    bufs = [pipe.frame_buffers() for _ in range(nslot)]
    rows_of = [0] * nslot
    got = [None] * F
    for f in range(F + nslot):
        slot = f % nslot
        if f >= nslot:
            pipe.wait_packed(slot)
            out, stream_of = pipe.expand(bufs[slot][0], bufs[slot][1], rows_of[slot], np.float32)
            got[f - nslot] = [out[stream_of == s] for s in range(S)]
        if f < F:
            rows_of[slot], flags = pipe.pack(bufs[slot][0], dets[f], ndets=nd[f], dtype=np.float32)
            pipe.submit_packed(slot, bufs[slot][0], bufs[slot][1], np.float32, flags)
    for f in range(F):
        for s in range(S):
            assert np.array_equal(got[f][s], ref[f][s]), (f, s)
    seq.close(); pipe.close()


def test_packed_capacity_overflow_is_reported_by_wait():
    from yolo_tracking_b200 import _lib
    from yolo_tracking_b200.batch import BatchedTracker
    S, D = 2, 32
    trk = BatchedTracker("bytetrack", S, max_tracks=32, max_dets=D, **BYTE)
    rng = np.random.default_rng(0)

    def frame(n):
        xy = rng.uniform(0, 3000, (n, 2))
        return np.concatenate([xy, xy + rng.uniform(20, 60, (n, 2)), rng.uniform(0.6, 0.9, (n, 1)), np.zeros((n, 1))], axis=1)
    # 30 + 30 distinct objects on consecutive frames: the second frame needs 60 slots of 32
    trk.update_frames([frame(30), frame(3)])
    with pytest.raises(_lib.B200TrackError) as e:
        trk.update_frames([frame(30), frame(3)])
    assert e.value.code == _lib.ERR_CAPACITY
    # the padded host interface reports it from wait_host as well
    trk.reset()
    d = np.zeros((S, D, 6)); n = np.zeros(S, dtype=np.int32)
    d[0, :30] = frame(30); n[0] = 30
    trk.update_batch(d, n)
    d[0, :30] = frame(30)
    with pytest.raises(_lib.B200TrackError) as e:
        trk.update_batch(d, n)
    assert e.value.code == _lib.ERR_CAPACITY
    trk.close()


def test_packed_device_blocks():
    """b200track_step_packed on device-resident blocks (the form a detector running on the same GPU would use)."""
    import torch
    from yolo_tracking_b200.batch import BatchedTracker
    S, N, F, D = 4, 20, 12, 32
    dets, nd, _ = _inputs("bytetrack", S, N, F, D, True)
    a = BatchedTracker("bytetrack", S, max_tracks=64, max_dets=D, **BYTE)
    b = BatchedTracker("bytetrack", S, max_tracks=64, max_dets=D, **BYTE)
    bi, bo = b.frame_buffers(pinned=False)
    for f in range(F):
        ref = a.update_frames([dets[f, s, :nd[f, s]].astype(np.float32) for s in range(S)])
        R, flags = b.pack(bi, dets[f], ndets=nd[f], dtype=np.float32)
        d_in = torch.from_numpy(bi).cuda()
        d_out = torch.empty(len(bo), dtype=torch.uint8, device="cuda")
        b.step_packed_device(d_in, R, d_out, np.float32, flags)
        torch.cuda.synchronize()
        bo[:] = d_out.cpu().numpy()
        out, stream_of = b.expand(bi, bo, R, np.float32)
        for s in range(S):
            assert np.array_equal(out[stream_of == s], ref[s]), (f, s)
    a.close(); b.close()
