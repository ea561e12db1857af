"""ORACLE (test infrastructure): the reference's constant-velocity Kalman filters, restated
as batched dense numpy (leading axis = tracks).  Dense 8x8 / 4x4 arithmetic on purpose: it
is independent of the CUDA path's block-sparse formulation.

Follows (reference file:line):
  boxmot/motion/kalman_filters/bytetrack_kf.py  initiate :55-86, predict :88-124,
      project :126-153, multi_predict :155-192, update :194-226, gating_distance :228-270
  boxmot/motion/kalman_filters/botsort_kf.py    same methods, noise from (w, h)
      :76-85, :107-116, :142-146, :170-179
  boxmot/motion/kalman_filters/strongsort_kf.py project(.., confidence) :124-155
      (std scaled by 1-confidence :148), update :157-189, gating_distance :191-233
Parity pinned by tests/golden/kf_*.npz (generated from the live reference).
"""
from __future__ import annotations

import numpy as np

W_POS = 1.0 / 20
W_VEL = 1.0 / 160

CHI2INV95 = {1: 3.8415, 2: 5.9915, 3: 7.8147, 4: 9.4877, 5: 11.070,
             6: 12.592, 7: 14.067, 8: 15.507, 9: 16.919}

F8 = np.eye(8)
F8[:4, 4:] = np.eye(4)
H48 = np.eye(4, 8)


def _std_terms(kind, ref, pos_scale, vel_scale, c_pos, c_vel):
    """ref: [n,4] (the measurement or mean[:, :4]).  Returns std[n,8]."""
    ref = np.asarray(ref, dtype=np.float64).reshape(-1, 4)
    n = ref.shape[0]
    std = np.empty((n, 8))
    if kind in ("xyah", "xyah_conf"):
        h = ref[:, 3]
        std[:, 0] = pos_scale * W_POS * h
        std[:, 1] = pos_scale * W_POS * h
        std[:, 2] = c_pos
        std[:, 3] = pos_scale * W_POS * h
        std[:, 4] = vel_scale * W_VEL * h
        std[:, 5] = vel_scale * W_VEL * h
        std[:, 6] = c_vel
        std[:, 7] = vel_scale * W_VEL * h
    elif kind == "xywh":
        w, h = ref[:, 2], ref[:, 3]
        std[:, 0] = pos_scale * W_POS * w
        std[:, 1] = pos_scale * W_POS * h
        std[:, 2] = pos_scale * W_POS * w
        std[:, 3] = pos_scale * W_POS * h
        std[:, 4] = vel_scale * W_VEL * w
        std[:, 5] = vel_scale * W_VEL * h
        std[:, 6] = vel_scale * W_VEL * w
        std[:, 7] = vel_scale * W_VEL * h
    else:
        raise ValueError(kind)
    return std


def _diag(v):
    n, k = v.shape
    out = np.zeros((n, k, k))
    idx = np.arange(k)
    out[:, idx, idx] = v
    return out


def initiate(kind, z):
    """z: [n,4] -> mean[n,8], cov[n,8,8]."""
    z = np.asarray(z, dtype=np.float64).reshape(-1, 4)
    mean = np.concatenate([z, np.zeros_like(z)], axis=1)
    std = _std_terms(kind, z, 2, 10, 1e-2, 1e-5)
    return mean, _diag(np.square(std))


def predict(kind, mean, cov):
    mean = np.asarray(mean, dtype=np.float64).reshape(-1, 8)
    cov = np.asarray(cov, dtype=np.float64).reshape(-1, 8, 8)
    std = _std_terms(kind, mean[:, :4], 1, 1, 1e-2, 1e-5)
    q = _diag(np.square(std))
    new_mean = mean @ F8.T
    new_cov = F8 @ cov @ F8.T + q
    return new_mean, new_cov


def project(kind, mean, cov, confidence=0.0):
    mean = np.asarray(mean, dtype=np.float64).reshape(-1, 8)
    cov = np.asarray(cov, dtype=np.float64).reshape(-1, 8, 8)
    std = _std_terms(kind, mean[:, :4], 1, 1, 1e-1, 0.0)[:, :4]
    if kind == "xyah_conf":
        std = (1 - np.asarray(confidence, dtype=np.float64).reshape(-1, 1)) * std
    pm = mean @ H48.T
    pc = H48 @ cov @ H48.T + _diag(np.square(std))
    return pm, pc


def update(kind, mean, cov, z, confidence=0.0):
    mean = np.asarray(mean, dtype=np.float64).reshape(-1, 8)
    cov = np.asarray(cov, dtype=np.float64).reshape(-1, 8, 8)
    z = np.asarray(z, dtype=np.float64).reshape(-1, 4)
    pm, pc = project(kind, mean, cov, confidence)
    # K = P H^T S^-1 ; the reference solves S K^T = (P H^T)^T by Cholesky (cho_factor/cho_solve)
    L = np.linalg.cholesky(pc)
    b = np.transpose(cov @ H48.T, (0, 2, 1))
    y1 = np.linalg.solve(L, b)
    kt = np.linalg.solve(np.transpose(L, (0, 2, 1)), y1)
    gain = np.transpose(kt, (0, 2, 1))
    innov = z - pm
    new_mean = mean + np.einsum("nk,nik->ni", innov, gain)
    new_cov = cov - gain @ pc @ np.transpose(gain, (0, 2, 1))
    return new_mean, new_cov


def gating_distance(kind, mean, cov, measurements, only_position=False, metric="maha",
                    confidence=0.0):
    """One track (mean[8], cov[8,8]) against measurements[D,4] -> d2[D]."""
    pm, pc = project(kind, np.asarray(mean).reshape(1, 8), np.asarray(cov).reshape(1, 8, 8),
                     confidence)
    pm, pc = pm[0], pc[0]
    m = np.asarray(measurements, dtype=np.float64).reshape(-1, 4)
    if only_position:
        pm, pc, m = pm[:2], pc[:2, :2], m[:, :2]
    d = m - pm
    if metric == "gaussian":
        return np.sum(d * d, axis=1)
    if metric != "maha":
        raise ValueError("invalid distance metric")
    L = np.linalg.cholesky(pc)
    zz = np.linalg.solve(L, d.T)
    return np.sum(zz * zz, axis=0)


def apply_warp(mean, cov, H):
    """STrack.multi_gmc (boxmot/trackers/botsort/bot_sort.py:95-111) on dense states [n, 8] / [n, 8, 8]:
    mean <- kron(I4, R) mean, mean[:2] += t, cov <- R8 cov R8^T with R = H[:2, :2], t = H[:2, 2]."""
    H = np.asarray(H, dtype=np.float64)
    R8 = np.kron(np.eye(4), H[:2, :2])
    mean = np.asarray(mean, dtype=np.float64) @ R8.T
    mean[:, :2] += H[:2, 2]
    cov = R8 @ np.asarray(cov, dtype=np.float64) @ R8.T
    return mean, cov
