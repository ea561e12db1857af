"""MOTChallenge formats either side of the path: real MOT17 public detections replayed through
the trackers, result rows as the reference's writer prints them (examples/utils.py:8-28).  The fixture
tests/golden/mot17_mini.npz holds the detection rows of three MOT17-mini sequences (first 300 frames) and the
integer MOT rows produced by the live reference's ByteTrack and OC-SORT."""
import numpy as np
import pytest

from _util import load_golden
from yolo_tracking_b200 import mot_io
from yolo_tracking_b200.replay import dense_frames, tracker_params

SEQS = ["MOT17-02-FRCNN", "MOT17-05-FRCNN", "MOT17-09-FRCNN"]


def _sequences(g):
    seqs = []
    for name in SEQS:
        frames, dets = mot_io.split_det_rows(g[name + "_det"])
        seqs.append(dense_frames(frames, dets, int(g[name + "_len"])))
    return seqs


def test_det_txt_round_trip(tmp_path):
    g = load_golden("mot17_mini")
    raw = g["MOT17-05-FRCNN_det"]
    path = tmp_path / "seq" / "det" / "det.txt"
    path.parent.mkdir(parents=True)
    np.savetxt(path, np.concatenate([raw, -np.ones((len(raw), 3))], axis=1), delimiter=",", fmt="%.10g")
    frames, dets = mot_io.read_det_txt(path)
    f2, d2 = mot_io.split_det_rows(raw)
    assert np.array_equal(frames, f2) and all(np.array_equal(a, b) for a, b in zip(dets, d2))
    assert frames[0] >= 1 and np.all(np.diff(frames) > 0)
    k = 3
    rows = raw[raw[:, 0] == frames[k]]
    assert np.array_equal(dets[k][:, 0], rows[:, 2]) and np.array_equal(dets[k][:, 2], rows[:, 2] + rows[:, 4])
    assert np.array_equal(dets[k][:, 4], rows[:, 6]) and np.all(dets[k][:, 5] == 0)
    # writer: frame is 1-based, ltwh, integers by truncation, trailing -1
    out = np.array([[10.7, 20.2, 50.9, 80.1, 3, 0.93, 0, 5], [1.5, 2.5, 4.0, 9.75, 7, 0.4, 2, 1]])
    txt = tmp_path / "res" / "a.txt"
    mot_io.write_mot_results(txt, out, 0)
    mot_io.write_mot_results(txt, out[:1], 1)
    got = np.loadtxt(txt, dtype=np.int64, ndmin=2)
    assert got.tolist() == [[1, 3, 10, 20, 40, 59, 0, 0, -1], [1, 7, 1, 2, 2, 7, 0, 2, -1], [2, 3, 10, 20, 40, 59, 0, 0, -1]]


@pytest.mark.parametrize("kind", ["bytetrack", "ocsort"])
def test_oracle_replay_matches_reference_files(kind):
    from oracle.bytetrack import ByteTrackOracle
    from oracle.ocsort import OCSortOracle
    g = load_golden("mot17_mini")
    p = tracker_params(kind)
    for name, seq in zip(SEQS, _sequences(g)):
        trk = ByteTrackOracle(**p) if kind == "bytetrack" else OCSortOracle(False, **p)
        rows = []
        for f, d in enumerate(seq):
            o = trk.update(d, (1080, 1920)) if kind == "ocsort" else trk.update(d)
            if o.size:
                rows.append(mot_io.mot_rows(o, f))
        assert np.array_equal(mot_io.as_int_rows(np.concatenate(rows)), g[f"{name}_{kind}"]), name


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["bytetrack", "ocsort"])
def test_cuda_replay_matches_reference_files(kind):
    """All three sequences as streams of one batched context; the files equal the reference's, row for row."""
    from yolo_tracking_b200.replay import replay
    g = load_golden("mot17_mini")
    res = replay(kind, _sequences(g), tracker_params(kind), img_hw=(1080, 1920))
    for name, rows in zip(SEQS, res):
        assert np.array_equal(mot_io.as_int_rows(rows), g[f"{name}_{kind}"]), name


@pytest.mark.gpu
def test_replay_cli_writes_result_files(tmp_path):
    from yolo_tracking_b200 import replay as rp
    g = load_golden("mot17_mini")
    for name in SEQS[1:]:
        d = tmp_path / "src" / name / "det"
        d.mkdir(parents=True)
        raw = g[name + "_det"]
        np.savetxt(d / "det.txt", np.concatenate([raw, -np.ones((len(raw), 3))], axis=1), delimiter=",", fmt="%.10g")
    rp.main(["--tracker", "bytetrack", "--source", str(tmp_path / "src"), "--out", str(tmp_path / "out")])
    for name in SEQS[1:]:
        got = np.loadtxt(tmp_path / "out" / (name + ".txt"), dtype=np.int64, ndmin=2)
        assert np.array_equal(got, g[f"{name}_bytetrack"]), name


def _deepocsort_replay(make, update):
    import sys
    from _util import GOLDEN
    if GOLDEN not in sys.path:
        sys.path.insert(0, GOLDEN)
    from scenarios import DEEPOCSORT_YAML, mot_feats
    g, ref = load_golden("mot17_mini"), load_golden("mot17_mini_deepocsort")
    for si, (name, seq) in enumerate(zip(SEQS, _sequences(g))):
        trk = make(**DEEPOCSORT_YAML)
        rows = []
        for f, d in enumerate(seq):
            raw = mot_feats(si, f, int((d[:, 4] > DEEPOCSORT_YAML["det_thresh"]).sum()))
            feats = raw / np.linalg.norm(raw) if len(raw) else raw
            o = update(trk, d, feats)
            if o.size:
                rows.append(mot_io.mot_rows(o, f))
        assert np.array_equal(mot_io.as_int_rows(np.concatenate(rows)), ref[name]), name


def test_deepocsort_oracle_replay_matches_reference_files():
    """Real MOT17 public detections (duplicate boxes, confidences down to 0.05) with seeded stand-in embeddings through the
    DeepOCSORT oracle: the result rows equal the live reference's (tests/golden/mot17_mini_deepocsort.npz)."""
    from oracle.deepocsort import DeepOCSortOracle
    _deepocsort_replay(lambda **kw: DeepOCSortOracle(**kw), lambda t, d, f: t.update(d, f, (1080, 1920)))


@pytest.mark.gpu
def test_deepocsort_cuda_replay_matches_reference_files():
    """GPU twin: the same streams through the fused CUDA DeepOCSORT step (the drop-in is a one-stream context of it)."""
    import yolo_tracking_b200 as pkg
    _deepocsort_replay(lambda **kw: pkg.DeepOCSORT(None, 0, False, False, **kw), lambda t, d, f: t.update(d, (1080, 1920), feats=f))


def _hybridsort_replay(make, update):
    import sys
    from _util import GOLDEN
    if GOLDEN not in sys.path:
        sys.path.insert(0, GOLDEN)
    from scenarios import HYBRIDSORT_YAML, mot_feats
    g, ref = load_golden("mot17_mini"), load_golden("mot17_mini_hybridsort")
    for si, (name, seq) in enumerate(zip(SEQS, _sequences(g))):
        trk = make(**HYBRIDSORT_YAML)
        rows = []
        for f, d in enumerate(seq):
            raw = mot_feats(si, f, len(d))
            feats = raw / np.linalg.norm(raw) if len(raw) else raw
            o = update(trk, d, feats)
            if o.size:
                rows.append(mot_io.mot_rows(o, f))
        assert np.array_equal(mot_io.as_int_rows(np.concatenate(rows)), ref[name]), name


def test_hybridsort_oracle_replay_matches_reference_files():
    """Real MOT17 public detections (duplicate boxes, confidences down to 0.05) with seeded stand-in embeddings through the
    HybridSORT oracle: the result rows equal the live reference's (tests/golden/mot17_mini_hybridsort.npz)."""
    from oracle.hybridsort import HybridSortOracle
    _hybridsort_replay(lambda **kw: HybridSortOracle(**kw), lambda t, d, f: t.update(d, f[d[:, 4] > 0], (1080, 1920)))


@pytest.mark.gpu
def test_hybridsort_cuda_replay_matches_reference_files():
    """GPU twin: the same streams through the fused CUDA HybridSORT step (the drop-in is a one-stream context of it)."""
    import yolo_tracking_b200 as pkg
    _hybridsort_replay(lambda **kw: pkg.HybridSORT(None, 0, False, **kw), lambda t, d, f: t.update(d, (1080, 1920), feats=f))


@pytest.mark.gpu
def test_deepocsort_cuda_replay_batched_streams():
    """The three sequences as three streams of ONE batched DeepOCSORT context (ragged frame counts: a finished sequence
    keeps sending empty frames), through the packed interface."""
    import sys
    from _util import GOLDEN
    if GOLDEN not in sys.path:
        sys.path.insert(0, GOLDEN)
    from scenarios import DEEPOCSORT_YAML, mot_feats
    from yolo_tracking_b200.batch import BatchedTracker
    g, ref = load_golden("mot17_mini"), load_golden("mot17_mini_deepocsort")
    seqs = _sequences(g)
    trk = BatchedTracker("deepocsort", len(seqs), max_tracks=256, max_dets=256, feat_dim=32, **DEEPOCSORT_YAML)
    rows = [[] for _ in seqs]
    for f in range(max(len(q) for q in seqs)):
        dets, feats = [], []
        for si, seq in enumerate(seqs):
            d = seq[f] if f < len(seq) else np.zeros((0, 6))
            keep = d[:, 4] > DEEPOCSORT_YAML["det_thresh"]
            raw = mot_feats(si, f, int(keep.sum()))
            full = np.zeros((len(d), 32), dtype=np.float32)
            if len(raw):
                full[keep] = raw / np.linalg.norm(raw)
            dets.append(d)
            feats.append(full)
        outs = trk.update_frames(dets, feats=feats, img_hw=(1080, 1920))
        for si, o in enumerate(outs):
            if o.size and f < len(seqs[si]):
                rows[si].append(mot_io.mot_rows(o, f))
    for name, r in zip(SEQS, rows):
        assert np.array_equal(mot_io.as_int_rows(np.concatenate(r)), ref[name]), name
    trk.close()


def test_strongsort_oracle_replay_matches_reference_files():
    """The same streams through the StrongSORT oracle (every detection row gets a stand-in embedding; scipy's tie
    behaviour on the clipped cost matrix decides ids on real data too): rows equal the live reference's."""
    import sys
    from _util import GOLDEN
    if GOLDEN not in sys.path:
        sys.path.insert(0, GOLDEN)
    from oracle.strongsort import StrongSORTOracle
    from scenarios import STRONGSORT_YAML, mot_feats
    g, ref = load_golden("mot17_mini"), load_golden("mot17_mini_strongsort")
    for si, (name, seq) in enumerate(zip(SEQS, _sequences(g))):
        trk = StrongSORTOracle(**STRONGSORT_YAML)
        rows = []
        for f, d in enumerate(seq):
            raw = mot_feats(si, f, len(d))
            o = trk.update(d, raw / np.linalg.norm(raw) if len(raw) else raw)
            if o.size:
                rows.append(mot_io.mot_rows(o, f))
        assert np.array_equal(mot_io.as_int_rows(np.concatenate(rows)), ref[name]), name


def test_botsort_oracle_replay_matches_reference_files():
    """The same streams through the BoT-SORT oracle (botsort.yaml; stand-in embeddings for the detections above
    track_high_thresh, normalised like the seam): rows equal the live reference's."""
    import sys
    from _util import GOLDEN
    if GOLDEN not in sys.path:
        sys.path.insert(0, GOLDEN)
    from oracle.botsort import BoTSORTOracle
    from scenarios import BOTSORT_YAML, mot_feats
    g, ref = load_golden("mot17_mini"), load_golden("mot17_mini_botsort")
    for si, (name, seq) in enumerate(zip(SEQS, _sequences(g))):
        trk = BoTSORTOracle(**BOTSORT_YAML)
        rows = []
        for f, d in enumerate(seq):
            raw = mot_feats(si, f, len(d))
            hi = np.nonzero(d[:, 4] > BOTSORT_YAML["track_high_thresh"])[0]
            feats = np.zeros_like(raw)
            if len(hi):
                feats[hi] = raw[hi] / np.linalg.norm(raw[hi])
            o = trk.update(d, feats)
            if o.size:
                rows.append(mot_io.mot_rows(o, f))
        assert np.array_equal(mot_io.as_int_rows(np.concatenate(rows)), ref[name]), name


@pytest.mark.gpu
def test_botsort_cuda_replay_matches_reference_files():
    """GPU twin of the BoT-SORT replay: the fused CUDA step on the MOT17-mini public detections."""
    import sys
    from _util import GOLDEN
    if GOLDEN not in sys.path:
        sys.path.insert(0, GOLDEN)
    import yolo_tracking_b200 as pkg
    from scenarios import BOTSORT_YAML, mot_feats
    g, ref = load_golden("mot17_mini"), load_golden("mot17_mini_botsort")
    for si, (name, seq) in enumerate(zip(SEQS, _sequences(g))):
        kw = {k: v for k, v in BOTSORT_YAML.items() if k != "cmc_method"}
        trk = pkg.BoTSORT(None, 0, False, feat_dim=128, max_tracks=256, max_dets=256, **kw)
        rows = []
        for f, d in enumerate(seq):
            raw = mot_feats(si, f, len(d))
            hi = np.nonzero(d[:, 4] > BOTSORT_YAML["track_high_thresh"])[0]
            feats = np.zeros((len(d), 128), dtype=np.float32)          # the 32-d stand-ins zero-padded to the kernel's row granularity
            if len(hi):
                feats[hi, :32] = raw[hi] / np.linalg.norm(raw[hi])
            o = trk.update(d, None, feats=feats)
            if o.size:
                rows.append(mot_io.mot_rows(o, f))
        assert np.array_equal(mot_io.as_int_rows(np.concatenate(rows)), ref[name]), name


@pytest.mark.gpu
def test_strongsort_cuda_replay_matches_reference_files():
    """GPU twin of the StrongSORT replay: the operator-backed drop-in on the MOT17-mini public detections."""
    import sys
    from _util import GOLDEN
    if GOLDEN not in sys.path:
        sys.path.insert(0, GOLDEN)
    import yolo_tracking_b200 as pkg
    from scenarios import STRONGSORT_YAML, mot_feats
    g, ref = load_golden("mot17_mini"), load_golden("mot17_mini_strongsort")
    for si, (name, seq) in enumerate(zip(SEQS, _sequences(g))):
        trk = pkg.StrongSORT(None, 0, False, **STRONGSORT_YAML)
        rows = []
        for f, d in enumerate(seq):
            raw = mot_feats(si, f, len(d))
            o = trk.update(d, np.zeros((1080, 1920, 3), dtype=np.uint8), feats=raw / np.linalg.norm(raw) if len(raw) else raw)
            if o.size:
                rows.append(mot_io.mot_rows(o, f))
        assert np.array_equal(mot_io.as_int_rows(np.concatenate(rows)), ref[name]), name
