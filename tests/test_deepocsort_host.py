"""Host-side list logic of the DeepOCSORT drop-in (yolo_tracking_b200/trackers/deepocsort.py) on CPU: the operator
namespace is monkeypatched with the oracle's arithmetic (tests/_util.py::OracleOps, test-only), so what is checked here
against the live reference's goldens is everything the drop-in does OUTSIDE the CUDA operators - batching of the frame's
Kalman updates, freeze / re-update bookkeeping, observation aliasing under camera motion, ids and output order.  The
`-m gpu` twin (tests/test_deepocsort_gpu.py) runs the same replay through the CUDA operators."""
import types

import numpy as np
import pytest

from _util import OracleOps, assert_close, check_deepocsort_frame, deepocsort_scenario, heavy_offsets


@pytest.mark.parametrize("name", ["deepocsort_c4", "deepocsort_churn", "deepocsort_cam", "deepocsort_noemb", "deepocsort_ciou"])
def test_deepocsort_host_logic_replays_reference(name, monkeypatch):
    from yolo_tracking_b200.trackers import deepocsort as mod
    monkeypatch.setattr(mod, "_ops", OracleOps)
    monkeypatch.setattr(mod, "_lib", types.SimpleNamespace(load=lambda: None, SIM=mod._lib.SIM))
    sc, cfg, dets, nd, feats, g = deepocsort_scenario(name)
    trk = mod.DeepOCSort(None, 0, False, False, **cfg)
    heavy = heavy_offsets(g)
    for f in range(sc["n_frames"]):
        out = trk.update(dets[f, :nd[f]], (1080, 1920), feats=feats[f], warp=None if sc["warps"] is None else sc["warps"][f])
        check_deepocsort_frame(name, f, out, trk.state(), g, heavy)
    assert_close(trk.state()["emb"], g["final_emb"], rel=1e-6, what="embeddings")
    assert trk.stats["oru"] > 50 and trk.stats["lap_frames"] > 20 and trk.stats["ocr_frames"] >= 1


def test_deepocsort_without_cuda_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from yolo_tracking_b200 import create_tracker, get_tracker_config
    with pytest.raises((RuntimeError, ImportError)):
        create_tracker("deepocsort", get_tracker_config("deepocsort"), None, 0, False, False)


def test_deepocsort_reference_known_answer_host(monkeypatch):
    from _util import RandomReID, deepocsort_known_answer
    from yolo_tracking_b200.trackers import deepocsort as mod
    import yolo_tracking_b200 as pkg
    monkeypatch.setattr(mod, "_ops", OracleOps)
    monkeypatch.setattr(mod, "_lib", types.SimpleNamespace(load=lambda: None, SIM=mod._lib.SIM))
    deepocsort_known_answer(lambda: pkg.create_tracker("deepocsort", pkg.get_tracker_config("deepocsort"), None, 0, False, False,
                                                       model=RandomReID()))


def test_deepocsort_edges_host(monkeypatch):
    from _util import deepocsort_edge_replay
    from oracle.deepocsort import DeepOCSortOracle
    from yolo_tracking_b200.trackers import deepocsort as mod
    monkeypatch.setattr(mod, "_ops", OracleOps)
    monkeypatch.setattr(mod, "_lib", types.SimpleNamespace(load=lambda: None, SIM=mod._lib.SIM))
    deepocsort_edge_replay(lambda **kw: mod.DeepOCSort(None, 0, False, False, **kw), lambda **kw: DeepOCSortOracle(**kw))
