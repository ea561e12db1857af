"""MOTChallenge text formats either side of the tracking path.

Reads public-detection files (`<seq>/det/det.txt`: frame, -1, left, top, width, height, conf, ...) into the
`dets[N, 6] = (x1, y1, x2, y2, conf, cls)` arrays `tracker.update` takes, and writes tracker output rows the way the
reference's `write_mot_results` does (examples/utils.py:8-28: `frame, id, left, top, width, height, conf, cls, -1`,
`np.savetxt(fmt='%d')`, frames 1-based), so the files feed TrackEval exactly like the reference's (`examples/val.py:239-259`).
"""
from __future__ import annotations

import configparser
import os

import numpy as np


def read_det_txt(path, cls: float = 0.0):
    """-> (frames[int64, sorted unique], list of dets[N_f, 6] float64 per frame, in file order within a frame)."""
    raw = np.loadtxt(path, delimiter=",", ndmin=2, dtype=np.float64)
    return split_det_rows(raw, cls)


def split_det_rows(raw, cls: float = 0.0):
    raw = np.asarray(raw, dtype=np.float64).reshape(-1, raw.shape[-1])
    frame = raw[:, 0].astype(np.int64)
    order = np.argsort(frame, kind="stable")
    raw, frame = raw[order], frame[order]
    frames, start = np.unique(frame, return_index=True)
    bounds = list(start) + [len(frame)]
    out = []
    for k in range(len(frames)):
        r = raw[bounds[k]:bounds[k + 1]]
        d = np.empty((len(r), 6))
        d[:, 0] = r[:, 2]
        d[:, 1] = r[:, 3]
        d[:, 2] = r[:, 2] + r[:, 4]
        d[:, 3] = r[:, 3] + r[:, 5]
        d[:, 4] = r[:, 6]
        d[:, 5] = cls
        out.append(d)
    return frames, out


def read_seqinfo(seq_dir):
    cp = configparser.ConfigParser()
    cp.read(os.path.join(seq_dir, "seqinfo.ini"))
    s = cp["Sequence"]
    return dict(name=s.get("name"), length=s.getint("seqLength"), width=s.getint("imWidth"), height=s.getint("imHeight"),
                frame_rate=s.getfloat("frameRate"))


def mot_rows(rows, frame_idx: int):
    """Tracker output rows [M, 8] (x1, y1, x2, y2, id, conf, cls, det_ind) of 0-based frame `frame_idx`
    -> MOT rows [M, 9] (frame + 1, id, left, top, width, height, conf, cls, -1), examples/utils.py:8-20."""
    rows = np.asarray(rows, dtype=np.float64).reshape(-1, 8)
    m = np.empty((len(rows), 9))
    m[:, 0] = frame_idx + 1
    m[:, 1] = rows[:, 4]
    m[:, 2] = rows[:, 0]
    m[:, 3] = rows[:, 1]
    m[:, 4] = rows[:, 2] - rows[:, 0]
    m[:, 5] = rows[:, 3] - rows[:, 1]
    m[:, 6] = rows[:, 5]
    m[:, 7] = rows[:, 6]
    m[:, 8] = -1
    return m


def write_mot_results(txt_path, rows, frame_idx: int):
    """Append one frame's rows to `txt_path` as integers (np.savetxt fmt='%d' truncates like the reference)."""
    os.makedirs(os.path.dirname(os.path.abspath(str(txt_path))) or ".", exist_ok=True)
    with open(str(txt_path), "ab+") as f:
        np.savetxt(f, mot_rows(rows, frame_idx), fmt="%d")


def as_int_rows(m):
    """The integers `fmt='%d'` prints (truncation towards zero)."""
    return np.trunc(np.asarray(m, dtype=np.float64)).astype(np.int64)
