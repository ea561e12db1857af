"""Per-phase cycle breakdown of the ByteTrack / BoT-SORT step kernel (thread 0 of every CTA, clock64).
usage: python tools/phase_profile.py [streams] [frames] [bytetrack|botsort]
B200_STEP_SMEM_PAD=170000 leaves one CTA per SM: the cycles of a CTA running alone (latency, not contention)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
F = int(sys.argv[2]) if len(sys.argv) > 2 else 40
KIND = sys.argv[3] if len(sys.argv) > 3 else "bytetrack"
bench.select_workload(KIND)
dets, nd, feats = bench.generate(S, 0, F, bench.host_cores())
import torch  # noqa: E402
from yolo_tracking_b200.batch import BatchedTracker  # noqa: E402

dev = torch.device("cuda", 0)
d_dets = torch.from_numpy(dets).to(dev).to(torch.float64)
d_nd = torch.from_numpy(nd).to(dev)
d_out = torch.empty((S, bench.MAX_TRACKS, 8), dtype=torch.float64, device=dev)
d_nout = torch.empty((S,), dtype=torch.int32, device=dev)
d_feats = torch.from_numpy(feats).to(dev) if feats is not None else None
trk = BatchedTracker(KIND, S, max_tracks=bench.MAX_TRACKS, max_dets=bench.MAX_DETS, feat_dim=bench.W["emb"], **bench.PARAMS)
warm = F // 2
for f in range(warm):
    trk.step_device(d_dets[f], d_nd[f], d_out, d_nout, d_feats=d_feats[f] if d_feats is not None else None, img_hw=bench.W['img_hw'])
trk.phase_cycles(reset=True)          # enable + zero
for f in range(warm, F):
    trk.step_device(d_dets[f], d_nd[f], d_out, d_nout, d_feats=d_feats[f] if d_feats is not None else None, img_hw=bench.W['img_hw'])
c = trk.phase_cycles()
names = {1: "load dets/means", 2: "det+track prep, lap_prepare", 3: "cell masks", 4: "graph pass 1", 5: "solve pass 1: augmentations",
         6: "pass-2 setup", 7: "graph pass 2", 8: "deferred KF + lifecycle", 9: "lost-list scan", 10: "solve pass 2: augmentations",
         11: "lost boxes + dedupe", 12: "graph pass 1: candidate walk", 13: "solve pass 1: classify + small", 14: "solve pass 2: classify + small", 15: "final scan + writes"}
if KIND == "ocsort":
    names = {1: "loads, motion step, k-previous observations", 2: "row / column lists", 3: "dense cost fill (similarity + direction term)", 4: "permutation test, row-reduction start",
             5: "dense augmentations (or shortcut)", 6: "BYTE / recovery rounds", 7: "Kalman update, freeze / unfreeze", 8: "births, output scan"}
n = c[0]
tot = sum(c[1:])
print(f"CTAs {n}, mean cycles/CTA {tot / n:.0f}")
for k in range(1, 16):
    if c[k]:
        print(f"  phase {k:2d} {names.get(k, ''):32s} {c[k] / n:9.0f} cyc  {100 * c[k] / tot:5.1f} %")
