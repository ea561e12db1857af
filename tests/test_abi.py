"""CPU-only checks of the drop-in boundary: the C-ABI library loads and exports every symbol
include/b200track.h declares; the host-side factory mirrors the reference's names."""
import ctypes as C

import numpy as np
import pytest

from yolo_tracking_b200 import _lib


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    declared = _lib.declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/b200track.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature in _lib.py"
    assert set(_lib.SIGNATURES) == set(declared)
    assert lib.b200track_abi_version() == 3


def test_config_struct_matches_header_field_order():
    # field names of b200track_config in the header, in order
    import re
    text = open(_lib.HEADER_PATH).read()
    body = text[text.index("typedef struct {"):text.index("} b200track_config;")]
    names = re.findall(r"\b(?:int32_t|double)\s+([a-z_]+);", body)
    assert names == [f[0] for f in _lib.Config._fields_]


def test_bad_arguments_are_reported_without_a_gpu():
    lib = _lib.load()
    ctx = C.c_void_p()
    cfg = _lib.Config()
    cfg.kind, cfg.n_streams, cfg.max_tracks, cfg.max_dets = 0, 0, 64, 64
    assert lib.b200track_create(C.byref(cfg), C.byref(ctx)) == _lib.ERR_ARG
    assert b"n_streams" in lib.b200track_last_error()
    cfg.n_streams, cfg.max_tracks = 1, 65
    assert lib.b200track_create(C.byref(cfg), C.byref(ctx)) == _lib.ERR_ARG
    cfg.max_tracks, cfg.kind = 64, 7
    assert lib.b200track_create(C.byref(cfg), C.byref(ctx)) == _lib.ERR_ARG
    assert lib.b200track_reset(None) == _lib.ERR_ARG


def test_no_cpu_fallback(has_cuda):
    if has_cuda:
        pytest.skip("GPU present")
    from yolo_tracking_b200.batch import BatchedTracker
    with pytest.raises(_lib.B200TrackError):
        BatchedTracker("bytetrack", 1, max_tracks=64, max_dets=64)
    from yolo_tracking_b200 import _ops
    with pytest.raises(RuntimeError):
        _ops.iou_distance(np.zeros((1, 4)), np.zeros((1, 4)))


def test_factory_names_and_configs():
    import yaml
    import yolo_tracking_b200 as pkg
    assert pkg.TRACKERS == ["bytetrack", "botsort", "ocsort", "strongsort", "deepocsort", "hybridsort"]     # boxmot/__init__.py:14
    for name in pkg.TRACKERS:
        path = pkg.get_tracker_config(name)
        assert path.name == name + ".yaml" and path.exists()
        yaml.safe_load(open(path))
    bt = yaml.safe_load(open(pkg.get_tracker_config("bytetrack")))
    assert (bt["track_thresh"], bt["match_thresh"], bt["track_buffer"], bt["frame_rate"]) == (0.5, 0.8, 30, 30)
    with pytest.raises(ValueError):
        pkg.create_tracker("nosuch", pkg.get_tracker_config("bytetrack"), None, 0, False, False)
    hy = yaml.safe_load(open(pkg.get_tracker_config("hybridsort")))
    assert (hy["det_thresh"], hy["iou_thresh"], hy["asso_func"], hy["min_hits"], hy["use_byte"]) == (0, 0.3, "giou", 1, False)
