"""GPU parity of the DeepOCSORT drop-in: (1) its operator kernels (csrc/kf8.cu) against the oracle's arithmetic on
seeded states, dense covariances included; (2) the three goldens of the live reference replayed through the CUDA
operators: ids, counters, observed / frozen flags exact, filter state / boxes / velocities to 1e-9 (fp64 tolerance of
BASELINE.json's north_star), embeddings to 1e-6 (the reference's float32 products)."""
import numpy as np
import pytest

from _util import OracleOps, assert_close, check_deepocsort_frame, deepocsort_scenario, heavy_offsets

pytestmark = pytest.mark.gpu


def _states(n, seed, dense):
    rng = np.random.default_rng(seed)
    x = np.concatenate([rng.uniform(100, 1800, (n, 2)), rng.uniform(20, 200, (n, 2)), rng.normal(0, 3, (n, 4))], axis=1)
    P = np.zeros((n, 8, 8))
    for i in range(n):
        if dense:
            a = rng.normal(0, 1, (8, 8))
            P[i] = a @ a.T + np.diag(rng.uniform(1, 30, 8))
        else:
            w, h = x[i, 2], x[i, 3]
            P[i] = np.diag([(w / 20) ** 2 * 4, (h / 20) ** 2 * 4, (w / 20) ** 2 * 4, (h / 20) ** 2 * 4,
                            (w / 160) ** 2 * 100, (h / 160) ** 2 * 100, (w / 160) ** 2 * 100, (h / 160) ** 2 * 100])
    return x, P


@pytest.mark.parametrize("dense", [False, True])
def test_kf8_operators_match_oracle(dense):
    from yolo_tracking_b200 import _ops
    n = 333                                                # not a multiple of the 64 tracks a CTA stages
    x, P = _states(n, 5 + dense, dense)
    rng = np.random.default_rng(77)
    for unit_q in (False, True):
        m, c = _ops.kf8_predict(x, P, unit_q)
        rm, rc = OracleOps.kf8_predict(x, P, unit_q)
        assert_close(m, rm, what="predict mean")
        assert_close(c, rc, abs_=1e-10, what="predict cov")
    z = x[:, :4] + rng.normal(0, 2, (n, 4))
    for wh in (None, x[:, 2:4] * rng.uniform(0.8, 1.2, (n, 2))):
        m, c = _ops.kf8_update(x, P, z, wh)
        rm, rc = OracleOps.kf8_update(x, P, z, wh)
        assert_close(m, rm, what="update mean")
        assert_close(c, rc, abs_=1e-10, what="update cov")
    # chained: several frames of predict + update stay within tolerance (no drift between the two implementations)
    m, c, rm, rc = x, P, x, P
    for _ in range(20):
        m, c = _ops.kf8_predict(m, c)
        rm, rc = OracleOps.kf8_predict(rm, rc)
        z = rm[:, :4] + rng.normal(0, 1, (n, 4))
        m, c = _ops.kf8_update(m, c, z, m[:, 2:4])
        rm, rc = OracleOps.kf8_update(rm, rc, z, rm[:, 2:4])
    assert_close(m, rm, what="chained mean")
    assert_close(c, rc, abs_=1e-10, what="chained cov")
    # the re-update: gaps 1..30, measurements read as [x, y, s, r]
    box1 = x[:, :4].copy()
    box2 = box1 + np.concatenate([rng.normal(0, 15, (n, 2)), rng.normal(0, 4, (n, 2))], axis=1)
    gap = rng.integers(1, 31, n)
    m, c, v = _ops.kf8_oru(x, P, box1, box2, gap)
    rm, rc, rv = OracleOps.kf8_oru(x, P, box1, box2, gap)
    assert_close(m, rm, what="oru mean")
    assert_close(c, rc, abs_=1e-10, what="oru cov")
    assert_close(v, rv, what="last virtual box")


def test_ocm_cost_and_dot_matrix_match_oracle():
    from yolo_tracking_b200 import _ops
    rng = np.random.default_rng(9)
    D, T, F = 57, 43, 130
    ctr = rng.uniform(100, 1000, (D, 2))
    dets5 = np.concatenate([ctr - 20, ctr + 20, rng.uniform(0.1, 1, (D, 1))], axis=1)
    pc = rng.uniform(100, 1000, (T, 2))
    prev5 = np.concatenate([pc - 25, pc + 25, rng.uniform(0.1, 1, (T, 1))], axis=1)
    prev5[::7] = -1                                                   # no previous observation
    vel = rng.normal(0, 1, (T, 2))
    vel /= np.linalg.norm(vel, axis=1, keepdims=True) + 1e-6
    vel[::5] = 0                                                       # no velocity yet: the term is exactly 0
    sim = rng.uniform(-1, 1, (D, T))
    sim[rng.random((D, T)) < 0.5] = 0.0                                # structural ties
    emb = rng.uniform(0, 1, (D, T))
    for e in (None, emb):
        got = _ops.ocm_cost(sim, dets5, vel, prev5, 0.2, e)
        ref = OracleOps.ocm_cost(sim, dets5, vel, prev5, 0.2, e)
        assert_close(got, ref, rel=1e-12, abs_=1e-15, what="ocm cost")
        tie = (sim == 0) & ((prev5[:, 4] < 0) | (vel[:, 0] == 0))[None, :]
        if e is None:
            assert np.array_equal(got[tie], ref[tie])                  # exact where only the tie-break decides
    assert np.array_equal(_ops.ocm_cost(sim), OracleOps.ocm_cost(sim))
    a, b = rng.normal(0, 1, (D, F)).astype(np.float32), rng.normal(0, 1, (T, F))
    assert_close(_ops.dot_matrix(a, b), a.astype(np.float64) @ b.T, rel=1e-12, abs_=1e-13, what="dot")
    assert _ops.dot_matrix(a[:0], b).shape == (0, T)


@pytest.mark.parametrize("name", ["deepocsort_c4", "deepocsort_churn", "deepocsort_cam", "deepocsort_noemb", "deepocsort_ciou"])
def test_deepocsort_replays_reference_golden(name):
    from yolo_tracking_b200 import DeepOCSORT
    sc, cfg, dets, nd, feats, g = deepocsort_scenario(name)
    trk = DeepOCSORT(None, 0, False, False, **cfg)
    img = np.zeros((1080, 1920, 3), dtype=np.uint8)
    heavy = heavy_offsets(g)
    for f in range(sc["n_frames"]):
        out = trk.update(dets[f, :nd[f]], img, feats=feats[f], warp=None if sc["warps"] is None else sc["warps"][f])
        check_deepocsort_frame(name, f, out, trk.state(), g, heavy)
    assert_close(trk.state()["emb"], g["final_emb"], rel=1e-6, what="embeddings")
    assert trk.stats["oru"] > 50 and trk.stats["lap_frames"] > 20 and trk.stats["ocr_frames"] >= 1


def test_deepocsort_factory_and_seam():
    from oracle.deepocsort import DeepOCSortOracle
    from yolo_tracking_b200 import create_tracker, get_tracker_config
    from yolo_tracking_b200.synth import make_stream
    dets, nd, embs = make_stream(4, 89, 12, 40, emb_dim=64, occlusion=True)

    class Seam:
        def get_features(self, xyxys, img):
            f = np.asarray(Seam.next, dtype=np.float32)
            assert len(f) == len(xyxys)
            return f / np.linalg.norm(f)
    trk = create_tracker("deepocsort", get_tracker_config("deepocsort"), None, 0, False, False, model=Seam())
    orc = DeepOCSortOracle(det_thresh=0, max_age=30, min_hits=1, iou_threshold=0.3, delta_t=3, asso_func="giou", inertia=0.2)
    img = np.zeros((1080, 1920, 3), dtype=np.uint8)
    assert trk.update(np.empty((0, 6)), img).size == 0
    orc.update(np.empty((0, 6)), np.zeros((0, 64), dtype=np.float32))
    for f in range(40):
        d = dets[f, :nd[f]]
        Seam.next = embs[f, :nd[f]]
        out = trk.update(d, img)
        raw = embs[f, :nd[f]].astype(np.float32)
        ref = orc.update(d, raw / np.linalg.norm(raw))
        assert out.shape == ref.shape
        if ref.size:
            assert np.array_equal(out[:, 4:], ref[:, 4:])
            assert_close(out[:, :4], ref[:, :4])
    with pytest.raises(AssertionError):
        trk.update(np.zeros((2, 5)), img)


def test_deepocsort_reference_known_answer():
    from _util import RandomReID, deepocsort_known_answer
    import yolo_tracking_b200 as pkg
    deepocsort_known_answer(lambda: pkg.create_tracker("deepocsort", pkg.get_tracker_config("deepocsort"), None, 0, False, False,
                                                       model=RandomReID()))


def test_deepocsort_edges():
    from _util import deepocsort_edge_replay
    from oracle.deepocsort import DeepOCSortOracle
    from yolo_tracking_b200 import DeepOCSORT
    deepocsort_edge_replay(lambda **kw: DeepOCSORT(None, 0, False, False, **kw), lambda **kw: DeepOCSortOracle(**kw))
