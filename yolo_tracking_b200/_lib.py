"""ctypes binding of libb200track.so (include/b200track.h).  There is no CPU fallback: if the
library is missing or a call fails, an exception is raised."""
from __future__ import annotations

import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200TRACK_LIB") or os.path.join(_HERE, "lib", "libb200track.so")   # override: A/B builds
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "b200track.h")

OK, ERR_ARG, ERR_CUDA, ERR_CAPACITY, ERR_STATE = 0, -1, -2, -3, -4
BYTETRACK, OCSORT, BOTSORT, DEEPOCSORT, STRONGSORT, HYBRIDSORT = 0, 1, 2, 3, 4, 5
KF_XYAH, KF_XYWH, KF_XYAH_CONF = 0, 1, 2
SIM = {"iou": 0, "giou": 1, "diou": 2, "ciou": 3, "centroid": 4}


class B200TrackError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"b200track error {code}: {msg}")
        self.code = code


class Config(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("n_streams", C.c_int32), ("max_tracks", C.c_int32), ("max_dets", C.c_int32),
        ("feat_dim", C.c_int32), ("device", C.c_int32),
        ("track_thresh", C.c_double), ("track_low_thresh", C.c_double), ("new_track_thresh", C.c_double),
        ("match_thresh", C.c_double), ("proximity_thresh", C.c_double), ("appearance_thresh", C.c_double),
        ("track_buffer", C.c_int32), ("frame_rate", C.c_int32),
        ("det_thresh", C.c_double), ("iou_thresh", C.c_double), ("inertia", C.c_double),
        ("max_age", C.c_int32), ("min_hits", C.c_int32), ("delta_t", C.c_int32), ("asso_func", C.c_int32),
        ("use_byte", C.c_int32), ("with_reid", C.c_int32), ("fuse_first_associate", C.c_int32),
        ("w_association_emb", C.c_double), ("alpha_fixed_emb", C.c_double), ("aw_param", C.c_double),
        ("embedding_off", C.c_int32), ("aw_off", C.c_int32), ("camera_motion", C.c_int32), ("reserved", C.c_int32),
        ("max_dist", C.c_double), ("max_iou_dist", C.c_double), ("mc_lambda", C.c_double), ("ema_alpha", C.c_double),
        ("n_init", C.c_int32), ("nn_budget", C.c_int32),
    ]


class Layout(C.Structure):
    _fields_ = [("in_bytes", C.c_uint64), ("in_off_offsets", C.c_uint64), ("in_off_warps", C.c_uint64),
                ("in_off_dets", C.c_uint64), ("in_off_feats", C.c_uint64),
                ("out_bytes", C.c_uint64), ("out_off_nout", C.c_uint64), ("out_off_rows", C.c_uint64),
                ("out_off_exc", C.c_uint64), ("row_bytes", C.c_int32), ("exc_capacity", C.c_int32)]


F32, F64 = 0, 1
FRAME_HAS_WARPS = 1
ROW_OC_NEW = 1 << 30
ROW_OC_STATE = 1 << 29

_P = C.c_void_p
_I = C.c_int32
_D = C.c_double

# name -> (restype, argtypes); must list every function declared in include/b200track.h
SIGNATURES = {
    "b200track_abi_version": (C.c_int, []),
    "b200track_last_error": (C.c_char_p, []),
    "b200track_create": (C.c_int, [C.POINTER(Config), C.POINTER(_P)]),
    "b200track_destroy": (None, [_P]),
    "b200track_reset": (C.c_int, [_P]),
    "b200track_step": (C.c_int, [_P, _P, _P, _P, _I, _I, _P, _P, _P]),
    "b200track_step_cam": (C.c_int, [_P, _P, _P, _P, _P, _I, _I, _P, _P, _P]),
    "b200track_get_track_extras": (C.c_int, [_P, _I, _P, _P]),
    "b200track_get_state_hybridsort": (C.c_int, [_P, _I, _P, _P, _P, _P, _P, _P, _P]),
    "b200track_counters": (C.c_int, [_P, C.POINTER(C.c_uint64 * 8)]),
    "b200track_step_host": (C.c_int, [_P, _P, _P, _P, _I, _I, _P, _P]),
    "b200track_host_slots": (C.c_int, [_P]),
    "b200track_submit_host": (C.c_int, [_P, _I, _P, _P, _P, _I, _I, _P, _P]),
    "b200track_wait_host": (C.c_int, [_P, _I]),
    "b200track_frame_layout": (C.c_int, [_P, C.c_int64, _I, C.POINTER(Layout)]),
    "b200track_step_packed": (C.c_int, [_P, _P, C.c_int64, _I, _I, _I, _I, _P, _P]),
    "b200track_submit_packed": (C.c_int, [_P, _I, _P, _I, _I, _I, _I, _P]),
    "b200track_wait_packed": (C.c_int, [_P, _I]),
    "b200track_sync": (C.c_int, [_P]),
    "b200track_track_updates": (C.c_int, [_P, C.POINTER(C.c_uint64)]),
    "b200track_launch_count": (C.c_int, [_P, C.POINTER(C.c_uint64)]),
    "b200track_phase_cycles": (C.c_int, [_P, C.POINTER(C.c_uint64 * 16), _I]),
    "b200track_footprint": (C.c_int, [_P, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "b200track_get_state": (C.c_int, [_P, _I, _P, _P, _P, _P, _P]),
    "b200track_get_features": (C.c_int, [_P, _I, _P]),
    "b200track_kf_initiate": (C.c_int, [_I, _I, _P, _P, _P, _P]),
    "b200track_kf_predict": (C.c_int, [_I, _I, _P, _P, _P]),
    "b200track_kf_apply_warp": (C.c_int, [_I, _P, _P, _P, _P, _P]),
    "b200track_aw_max_metric": (C.c_int, [_I, _I, _I, _P, _D, _D, _P, _P]),
    "b200track_kf_project": (C.c_int, [_I, _I, _P, _P, _P, _P, _P, _P]),
    "b200track_kf_update": (C.c_int, [_I, _I, _P, _P, _P, _P, _P]),
    "b200track_kf_gating_distance": (C.c_int, [_I, _I, _I, _P, _P, _P, _I, _I, _P, _P, _P]),
    "b200track_kf_gating_distance_batched": (C.c_int, [_I, _I, _I, _I, _P, _P, _P, _I, _I, _P, _P, _P]),
    "b200track_gate_cost": (C.c_int, [_I, _I, _I, _I, _P, _P, _P, _I, _I, _D, _P, _P, _P]),
    "b200track_box_similarity": (C.c_int, [_I, _I, _I, _P, _P, _D, _D, _P, _P]),
    "b200track_iou_distance": (C.c_int, [_I, _I, _P, _P, _P, _P, _P]),
    "b200track_embedding_distance": (C.c_int, [_I, _I, _I, _P, _P, _P, _P]),
    "b200track_appearance_cost_workspace": (C.c_int, [_I, _I, _I, _I, C.POINTER(C.c_uint64)]),
    "b200track_appearance_cost": (C.c_int, [_I, _I, _I, _I, _P, _P, _P, _D, _D, _D, _P, _P, C.c_uint64, _P, _P]),
    "b200track_gallery_cost_workspace": (C.c_int, [_I, _I, _I, _I, _I, _I, C.POINTER(C.c_uint64)]),
    "b200track_gallery_cost": (C.c_int, [_I, _I, _I, _I, _I, _P, _P, _P, _P, _D, _D, _P, _P, C.c_uint64, _P, _P]),
    "b200track_unit_bf16": (C.c_int, [C.c_int64, _I, _P, _P, _P]),
    "b200track_kf_xysr_predict": (C.c_int, [_I, _P, _P, _P, _P]),
    "b200track_kf_xysr_update": (C.c_int, [_I, _P, _P, _P, _P, _P]),
    "b200track_kf_xysr_unfreeze_update": (C.c_int, [_I, _P, _P, _P, _P, _P, _P, _P, _P]),
    "b200track_kf_xyscr_predict": (C.c_int, [_I, _P, _P, _P, _P]),
    "b200track_kf_xyscr_update": (C.c_int, [_I, _P, _P, _P, _P, _P]),
    "b200track_kf_xyscr_unfreeze_update": (C.c_int, [_I, _P, _P, _P, _P, _P, _P, _P, _P]),
    "b200track_gallery_append": (C.c_int, [_I, _I, _I, _P, _P, _P, _P, _P, _P]),
    "b200track_ema_unit_features": (C.c_int, [_I, _I, _P, _P, _D, _P]),
    "b200track_unit_features": (C.c_int, [_I, _I, _P, _P]),
    "b200track_camera_update_xyah": (C.c_int, [_I, _P, _P, _P]),
    "b200track_nn_cosine_distance": (C.c_int, [_I, _I, _I, _P, _P, _P, _P, _P]),
    "b200track_linear_sum_assignment": (C.c_int, [_I, _I, _I, _P, _P, _P, _P]),
    "b200track_lapjv": (C.c_int, [_I, _I, _I, _P, _D, _P, _P, _P]),
    "b200track_kf8_predict": (C.c_int, [_I, _P, _P, _I, _P]),
    "b200track_kf8_update": (C.c_int, [_I, _P, _P, _P, _P, _P]),
    "b200track_kf8_oru": (C.c_int, [_I, _P, _P, _P, _P, _P, _P, _P]),
    "b200track_ocm_cost": (C.c_int, [_I, _I, _P, _P, _P, _D, _P, _P, _P, _P]),
    "b200track_dot_matrix": (C.c_int, [_I, _I, _I, _P, _P, _P, _P]),
}

_lib = None


def declared_symbols():
    """Function names declared in include/b200track.h (used by the ABI export test)."""
    with open(HEADER_PATH) as f:
        text = f.read()
    return sorted(set(re.findall(r"\b(b200track_[a-z_0-9]+)\s*\(", text)))


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not built - run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(yolo_tracking_b200 has no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc):
    if rc != 0:
        msg = load().b200track_last_error()
        raise B200TrackError(rc, msg.decode() if msg else "")
    return rc
