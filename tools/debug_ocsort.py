"""Debug aid: replay an OC-SORT golden through the CUDA step and the oracle side by side and
print the first frame where they part.  usage: python tools/debug_ocsort.py [golden-name]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ocsort import OCSortOracle  # noqa: E402
from yolo_tracking_b200.batch import BatchedTracker  # noqa: E402

np.set_printoptions(linewidth=220, precision=6, suppress=True)
name = sys.argv[1] if len(sys.argv) > 1 else "ocsort_c2"
g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
p = g["params"]
cfg = dict(det_thresh=p[0], max_age=int(p[1]), min_hits=int(p[2]), asso_threshold=p[3], delta_t=int(p[4]),
           asso_func="giou", inertia=p[5])
hw = tuple(int(v) for v in g["img_hw"])
trk = BatchedTracker("ocsort", 1, max_tracks=128, max_dets=128, **cfg)
orc = OCSortOracle(False, use_byte=False, **cfg)
dets, nd = g["dets"], g["ndets"]
for f in range(dets.shape[0]):
    d = np.zeros((1, 128, 6))
    d[0, :nd[f]] = dets[f, :nd[f]]
    out, nout = trk.update_batch(d, np.array([nd[f]], dtype=np.int32), img_hw=hw)
    before = dict(orc.stats)
    ref = orc.update(dets[f, :nd[f]], hw).reshape(-1, 8)
    st, sn = trk.state(0), orc.snapshot()
    mine = np.stack([st[k] for k in ("track_id", "age", "time_since_update", "hits", "hit_streak", "observed")], 1)
    theirs = np.stack([sn[k] for k in ("track_id", "age", "time_since_update", "hits", "hit_streak", "observed")], 1)
    ok = nout[0] == len(ref) and mine.shape == theirs.shape and np.array_equal(mine, theirs)
    if ok:
        o = out[0, :nout[0]]
        ok = np.array_equal(o[:, 4:], ref[:, 4:]) and np.allclose(o[:, :4], ref[:, :4], rtol=1e-9, atol=1e-12)
        ok = ok and np.allclose(st["x"], sn["x"], rtol=1e-9, atol=1e-12) and np.allclose(st["P"], sn["P"], rtol=1e-9, atol=1e-9)
        ok = ok and np.allclose(st["velocity"], sn["velocity"], rtol=1e-9, atol=1e-12)
        ok = ok and np.allclose(st["last_observation"], sn["last_observation"], rtol=1e-9, atol=1e-12)
    print(f"frame {f}: nd {nd[f]} rows {nout[0]}/{len(ref)} trackers {len(mine)}/{len(theirs)} "
          f"lap {orc.stats['lap_frames'] - before['lap_frames']} ocr {orc.stats['ocr_frames'] - before['ocr_frames']} "
          f"oru {orc.stats['oru'] - before['oru']} {'ok' if ok else 'MISMATCH'}")
    if not ok:
        print("cuda records (id age tsu hits streak observed):\n", mine.T)
        print("oracle records:\n", theirs.T)
        print("cuda out ids/det_ind:\n", out[0, :nout[0]][:, [4, 7]].T)
        print("oracle out ids/det_ind:\n", ref[:, [4, 7]].T)
        n = min(len(mine), len(theirs))
        dx = np.abs(st["x"][:n] - sn["x"][:n]).max(axis=1)
        print("max |dx| per tracker:\n", dx)
        dv = np.abs(st["velocity"][:n] - sn["velocity"][:n]).max(axis=1)
        print("max |dvel| per tracker:\n", dv)
        dl = np.abs(st["last_observation"][:n] - sn["last_observation"][:n]).max(axis=1)
        print("max |dlast| per tracker:\n", dl)
        dP = np.abs(st["P"][:n] - sn["P"][:n]).reshape(n, -1).max(axis=1)
        print("max |dP| per tracker:\n", dP)
        break
