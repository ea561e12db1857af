"""GPU edge cases of the batched frame step (SURVEY.md A.6 quirks and the boundary's error behaviour): threshold holes,
ragged / empty streams, capacity overflow reporting, the largest kernel variant, reset, the pipelined host API and the
device-pointer API on a caller's stream."""
import numpy as np
import pytest

from _util import assert_close

pytestmark = pytest.mark.gpu
BT = dict(track_thresh=0.5, match_thresh=0.8, track_buffer=30, frame_rate=30)


def _oracle_rows(orc, d):
    return orc.update(d).reshape(-1, 8)


def test_threshold_holes_and_ragged_streams():
    """conf == track_thresh and conf <= 0.1 fall in neither band (byte_tracker.py:151-158); streams without detections
    still age their tracks; every stream matches its own oracle."""
    from oracle.bytetrack import ByteTrackOracle
    from yolo_tracking_b200.batch import BatchedTracker
    from yolo_tracking_b200.synth import make_batch
    S, F, cap = 5, 40, 64
    dets, nd, _ = make_batch(1, S, 20, F, dmax=cap, first_stream=300, miss_prob=0.1)
    rng = np.random.default_rng(3)
    for f in range(F):
        for s in range(S):
            n = nd[f, s]
            if n >= 4:
                dets[f, s, 0, 4] = 0.5            # exactly track_thresh: ignored
                dets[f, s, 1, 4] = 0.1            # exactly the low bound: ignored
                dets[f, s, 2, 4] = 0.05           # below it: ignored
        if 10 <= f < 14:
            nd[f, 1] = 0                          # a stream that goes dark for a while ...
        if f >= 20:
            nd[f, 3] = 0                          # ... and one that ends (its tracks age out)
    nd[:, 4] = 0                                  # a stream that never sees anything
    trk = BatchedTracker("bytetrack", S, max_tracks=cap, max_dets=cap, **BT)
    orc = [ByteTrackOracle(0.5, 0.8, 30, 30) for _ in range(S)]
    for f in range(F):
        out, nout = trk.update_batch(np.ascontiguousarray(dets[f]), np.ascontiguousarray(nd[f]))
        for s in range(S):
            ref = _oracle_rows(orc[s], dets[f, s, :nd[f, s]])
            assert nout[s] == len(ref), (f, s)
            assert np.array_equal(out[s, :nout[s], 4:], ref[:, 4:]), (f, s)
            assert_close(out[s, :nout[s], :4], ref[:, :4])
    st = trk.state(4)
    assert st["n_tracked"] == 0 and st["n_lost"] == 0 and st["frame_id"] == F
    trk.sync()
    assert trk.track_updates() == sum(o.track_updates for o in orc)
    trk.close()


def test_capacity_overflow_is_reported_not_silent():
    from yolo_tracking_b200 import _lib
    from yolo_tracking_b200.batch import BatchedTracker
    cap = 32
    rng = np.random.default_rng(0)

    def boxes(n):
        c = np.stack([rng.uniform(50, 1800, n), rng.uniform(50, 1000, n)], axis=1)
        return np.concatenate([c - 15, c + 15, np.full((n, 1), 0.9), np.zeros((n, 1))], axis=1)
    trk = BatchedTracker("bytetrack", 1, max_tracks=cap, max_dets=cap, **BT)
    d = np.zeros((1, cap, 6))
    d[0] = boxes(cap)
    # the step that overflowed reports it when its results are collected (b200track_wait_host), and b200track_sync repeats
    # it for callers of the asynchronous device interface
    with pytest.raises(_lib.B200TrackError) as e:
        trk.update_batch(d, np.array([cap + 5], dtype=np.int32))      # claims more detections than the buffer holds
    assert e.value.code == _lib.ERR_CAPACITY and "detections" in str(e.value)
    with pytest.raises(_lib.B200TrackError) as e:
        trk.sync()
    assert e.value.code == _lib.ERR_CAPACITY and "detections" in str(e.value)
    trk.reset()
    with pytest.raises(_lib.B200TrackError) as e:
        for _ in range(3):                                            # 3 x 32 well separated boxes -> > 32 tracks (lost + new)
            d[0] = boxes(cap)
            trk.update_batch(d, np.array([cap], dtype=np.int32))
    assert e.value.code == _lib.ERR_CAPACITY and "tracks" in str(e.value)
    with pytest.raises(_lib.B200TrackError) as e:
        trk.sync()
    assert e.value.code == _lib.ERR_CAPACITY and "tracks" in str(e.value)
    trk.close()
    with pytest.raises(_lib.B200TrackError):
        BatchedTracker("bytetrack", 1, max_tracks=1024, max_dets=64, **BT)          # no kernel variant that large


def test_largest_variant_300_objects():
    from oracle.bytetrack import ByteTrackOracle
    from yolo_tracking_b200.batch import BatchedTracker
    from yolo_tracking_b200.synth import make_batch
    dets, nd, _ = make_batch(5, 2, 300, 12, dmax=512, first_stream=40)
    trk = BatchedTracker("bytetrack", 2, max_tracks=512, max_dets=512, **BT)
    orc = [ByteTrackOracle(0.5, 0.8, 30, 30) for _ in range(2)]
    for f in range(12):
        out, nout = trk.update_batch(np.ascontiguousarray(dets[f]), np.ascontiguousarray(nd[f]))
        for s in range(2):
            ref = _oracle_rows(orc[s], dets[f, s, :nd[f, s]])
            assert nout[s] == len(ref) and np.array_equal(out[s, :nout[s], 4:], ref[:, 4:]), (f, s)
            assert_close(out[s, :nout[s], :4], ref[:, :4])
    trk.sync()
    trk.close()


def test_reset_pipelined_and_device_apis_agree():
    """b200track_reset, b200track_submit_host / _wait_host (3 slots in flight) and b200track_step on device buffers with a
    caller-supplied stream all give the rows of the synchronous host call."""
    import torch
    from yolo_tracking_b200.batch import BatchedTracker
    from yolo_tracking_b200.synth import make_batch
    S, F, cap = 6, 9, 64
    dets, nd, _ = make_batch(1, S, 25, F, dmax=cap, first_stream=700)
    trk = BatchedTracker("bytetrack", S, max_tracks=cap, max_dets=cap, **BT)
    ref = []
    for f in range(F):
        out, nout = trk.update_batch(np.ascontiguousarray(dets[f]), np.ascontiguousarray(nd[f]))
        ref.append([out[s, :nout[s]].copy() for s in range(S)])
    base_updates = trk.track_updates()
    # pipelined host API
    trk.reset()
    assert trk.track_updates() == 0 and trk.state(0)["frame_id"] == 0
    nslot = trk.host_slots
    pin = [(torch.empty((S, cap, 8), dtype=torch.float64).pin_memory(), torch.empty((S,), dtype=torch.int32).pin_memory()) for _ in range(nslot)]
    got = [None] * F

    def collect(f):
        o, n = pin[f % nslot]
        got[f] = [o.numpy()[s, :n.numpy()[s]].copy() for s in range(S)]
    hold = []
    for f in range(F):
        slot = f % nslot
        if f >= nslot:
            trk.wait(slot)
            collect(f - nslot)
        a, b = np.ascontiguousarray(dets[f]), np.ascontiguousarray(nd[f])
        hold.append((a, b))                                           # host buffers must outlive the wait
        trk.submit(slot, a, b, pin[slot][0].numpy(), pin[slot][1].numpy())
    for f in range(max(0, F - nslot), F):
        trk.wait(f % nslot)
        collect(f)
    for f in range(F):
        for s in range(S):
            assert np.array_equal(got[f][s], ref[f][s]), (f, s)
    assert trk.track_updates() == base_updates
    # device-pointer API on a caller's stream
    trk.reset()
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(device=dev)
    d_dets, d_nd = torch.from_numpy(dets).to(dev), torch.from_numpy(nd).to(dev)
    d_out = torch.empty((S, cap, 8), dtype=torch.float64, device=dev)
    d_nout = torch.empty((S,), dtype=torch.int32, device=dev)
    with torch.cuda.stream(stream):
        for f in range(F):
            trk.step_device(d_dets[f], d_nd[f], d_out, d_nout, stream=stream.cuda_stream)
    stream.synchronize()
    o, n = d_out.cpu().numpy(), d_nout.cpu().numpy()
    for s in range(S):
        assert np.array_equal(o[s, :n[s]], ref[F - 1][s]), s
    trk.sync()
    trk.close()


@pytest.mark.parametrize("scale, cap", [(0.08, 64), (0.2, 64), (0.35, 64), (0.1, 224), (0.25, 224)])
def test_crowded_scenes_overflow_the_pair_and_edge_caches(scale, cap):
    """Objects squeezed into a corner of the canvas: every box overlaps many others, so the candidate pair list and / or
    the edge cache of the step kernel overflow and the solver runs on the bitmask form of the graph with recomputed
    costs (large connected components, general shortest-augmenting-path solver).  Results must still equal the oracle's."""
    from oracle.bytetrack import ByteTrackOracle
    from yolo_tracking_b200.batch import BatchedTracker
    from yolo_tracking_b200.synth import make_batch
    n_obj = 40 if cap == 64 else 150
    dets, nd, _ = make_batch(7, 2, n_obj, 25, dmax=cap, fp_rate=0.5)
    dets = dets.copy()
    ctr = 0.5 * (dets[..., :2] + dets[..., 2:4])
    half = 0.5 * (dets[..., 2:4] - dets[..., :2])
    dets[..., :2] = ctr * scale - half                  # centres pulled together, sizes kept
    dets[..., 2:4] = ctr * scale + half
    dets[np.arange(cap)[None, None, :] >= nd[:, :, None]] = 0.0
    trk = BatchedTracker("bytetrack", 2, max_tracks=cap, max_dets=cap, track_thresh=0.5, match_thresh=0.8, track_buffer=30,
                         frame_rate=30)
    oracles = [ByteTrackOracle(0.5, 0.8, 30, 30) for _ in range(2)]
    for f in range(25):
        out, nout = trk.update_batch(np.ascontiguousarray(dets[f]), np.ascontiguousarray(nd[f]))
        for s in range(2):
            ref = oracles[s].update(dets[f, s, :nd[f, s]], None).reshape(-1, 8)
            o = out[s, :nout[s]]
            assert o.shape == ref.shape, f"frame {f} stream {s}"
            assert np.array_equal(o[:, 4:], ref[:, 4:]), f"frame {f} stream {s}: ids"
            assert_close(o[:, :4], ref[:, :4], what=f"frame {f} stream {s} boxes")
    trk.close()
