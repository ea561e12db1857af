"""Attribute the warp-stall samples of an ncu --set full capture to CUDA source lines.
ncu's CSV source page is SASS-only, so the line table comes from `nvdisasm -g` of the same cubin and is joined by
instruction order.  usage: python tools/ncu_lines.py <rep | source.csv[.gz]> <cubin-name-substring> <mangled-kernel-substring> [top]"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

rep, cubin_sub, kern_sub = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "yolo_tracking_b200", "lib", "libb200track.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
cands = sorted(f for f in os.listdir(tmp) if cubin_sub in f and f.count("-") == 0)
cubin = ([f for f in cands if f.startswith(cubin_sub)] or cands)[0]          # "ocsort_step" must not pick deepocsort_step
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
# instruction index -> (file, line) for the wanted function
lines, cur, infn = [], ("?", 0), False
for l in dis:
    if l.startswith(".text."):
        infn = kern_sub in l
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        lines.append((cur, l.strip()))
if rep.endswith(".csv.gz"):                       # the source page exported on the GPU box (tools/profile_round.sh)
    import gzip
    txt = gzip.open(rep, "rt").read()
elif rep.endswith(".csv"):
    txt = open(rep).read()
else:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = rows[1]
ci = {n: i for i, n in enumerate(hdr)}
body = rows[2:]
print(f"{len(body)} SASS rows in the report, {len(lines)} instructions disassembled")
agg = defaultdict(lambda: defaultdict(float))
stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
tot = 0
for k, r in enumerate(body):
    if k >= len(lines):
        break
    key = lines[k][0]
    s = float(r[ci["# Samples"]] or 0)
    agg[key]["samples"] += s
    agg[key]["inst"] += float(r[ci["Instructions Executed"]] or 0)
    agg[key]["tinst"] += float(r[ci["Thread Instructions Executed"]] or 0)
    tot += s
    for c in stall_cols:
        agg[key][c] += float(r[ci[c]] or 0)
src_cache = {}


def src(f, n):
    if f not in src_cache:
        p = os.path.join(ROOT, "yolo_tracking_b200", "csrc", f)
        src_cache[f] = open(p).read().splitlines() if os.path.exists(p) else []
    s = src_cache[f]
    return s[n - 1].strip()[:90] if 0 < n <= len(s) else ""


print(f"total samples {tot:.0f}")
if os.environ.get("BY_INST"):
    ti = sum(a["inst"] for a in agg.values())
    print(f"total warp instructions {ti / 1e6:.1f} M; by instruction count:")
    byfile = defaultdict(lambda: [0.0, 0.0])
    for key, a in agg.items():
        byfile[key[0]][0] += a["inst"]
        byfile[key[0]][1] += a["tinst"]
    for f, (i, t) in sorted(byfile.items(), key=lambda kv: -kv[1][0]):
        print(f"  {f:32s} {i / 1e6:8.2f} M inst  {100 * i / ti:5.1f} %   avg lanes {t / max(i, 1):5.1f}")
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1]["inst"])[:top]:
        print(f"{100 * a['inst'] / ti:5.1f}%  inst {a['inst'] / 1e6:7.2f}M  lanes {a['tinst'] / max(a['inst'], 1):5.1f}  {key[0]}:{key[1]:<4d} {src(*key)}")
    sys.exit(0)
for key, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    st = sorted(((c[6:], a[c]) for c in stall_cols if a[c] > 0), key=lambda x: -x[1])[:3]
    print(f"{100 * a['samples'] / tot:5.1f}%  inst {a['inst'] / 1e6:7.2f}M  {key[0]}:{key[1]:<4d} {src(*key)}\n        " +
          ", ".join(f"{n} {100 * v / max(a['samples'], 1):.0f}%" for n, v in st))
