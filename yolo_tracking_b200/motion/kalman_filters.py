"""The reference's functional Kalman-filter objects, computed on the GPU.

Same method names, argument meaning and return shapes as
boxmot/motion/kalman_filters/{bytetrack_kf,botsort_kf,strongsort_kf}.py `KalmanFilter`
(initiate :55, predict :88, project :126, multi_predict :155, update :194,
gating_distance :228); every call goes through the C-ABI operator kernels.
"""
from __future__ import annotations

import numpy as np

from .. import _lib, _ops

chi2inv95 = {1: 3.8415, 2: 5.9915, 3: 7.8147, 4: 9.4877, 5: 11.070, 6: 12.592, 7: 14.067, 8: 15.507, 9: 16.919}


class _KalmanFilterBase:
    _kind = _lib.KF_XYAH

    def initiate(self, measurement):
        m, c = _ops.kf_initiate(self._kind, np.asarray(measurement, dtype=np.float64).reshape(1, 4))
        return m[0], c[0]

    def predict(self, mean, covariance):
        m, c = _ops.kf_predict(self._kind, np.asarray(mean).reshape(1, 8), np.asarray(covariance).reshape(1, 8, 8))
        return m[0], c[0]

    def multi_predict(self, mean, covariance):
        return _ops.kf_predict(self._kind, mean, covariance)

    def project(self, mean, covariance):
        m, c = _ops.kf_project(self._kind, np.asarray(mean).reshape(1, 8), np.asarray(covariance).reshape(1, 8, 8))
        return m[0], c[0]

    def update(self, mean, covariance, measurement):
        m, c = _ops.kf_update(self._kind, np.asarray(mean).reshape(1, 8), np.asarray(covariance).reshape(1, 8, 8),
                              np.asarray(measurement).reshape(1, 4))
        return m[0], c[0]

    def multi_update(self, mean, covariance, measurement):
        """Batched update (no reference analogue: the reference loops over tracks)."""
        return _ops.kf_update(self._kind, mean, covariance, measurement)

    def multi_gmc(self, mean, covariance, H=None):
        """STrack.multi_gmc (bot_sort.py:95-111) in array form: apply a 2x3 camera-motion warp to [n, 8] / [n, 8, 8] states."""
        H = np.eye(2, 3) if H is None else H
        return _ops.kf_apply_warp(mean, covariance, H)

    def gating_distance(self, mean, covariance, measurements, only_position=False, metric="maha"):
        return _ops.kf_gating_distance(self._kind, np.asarray(mean).reshape(1, 8), np.asarray(covariance).reshape(1, 8, 8),
                                       measurements, only_position, metric)[0]


class KalmanFilterXYAH(_KalmanFilterBase):
    """bytetrack_kf.py:23 (state x, y, a, h, vx, vy, va, vh)."""
    _kind = _lib.KF_XYAH


class KalmanFilterXYWH(_KalmanFilterBase):
    """botsort_kf.py:23 (state x, y, w, h, vx, vy, vw, vh)."""
    _kind = _lib.KF_XYWH


class KalmanFilterXYAHConf(_KalmanFilterBase):
    """strongsort_kf.py:21: measurement noise scaled by (1 - confidence)."""
    _kind = _lib.KF_XYAH_CONF

    def project(self, mean, covariance, confidence=0.0):
        m, c = _ops.kf_project(self._kind, np.asarray(mean).reshape(1, 8), np.asarray(covariance).reshape(1, 8, 8), confidence)
        return m[0], c[0]

    def update(self, mean, covariance, measurement, confidence=0.0):
        m, c = _ops.kf_update(self._kind, np.asarray(mean).reshape(1, 8), np.asarray(covariance).reshape(1, 8, 8),
                              np.asarray(measurement).reshape(1, 4), confidence)
        return m[0], c[0]

    def gating_distance(self, mean, covariance, measurements, only_position=False):
        return _ops.kf_gating_distance(self._kind, np.asarray(mean).reshape(1, 8), np.asarray(covariance).reshape(1, 8, 8),
                                       measurements, only_position, "maha", 0.0)[0]
