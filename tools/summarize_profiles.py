"""Turn the ncu outputs of tools/profile_round.sh (gpurun_out/rNN_*) into the tracked summaries under profiles/:
  profiles/rNN_launches_<workload>.csv   copied launch lists
  profiles/rNN_ncu_summary.md            per-kernel share tables + key metrics of every --set full capture
  profiles/traffic.json                  DRAM bytes per launch of the step kernels (bench.py's roofline.traffic)
usage: python tools/summarize_profiles.py rNN"""
import csv
import io
import json
import os
import shutil
import subprocess
import sys
from collections import defaultdict

R = sys.argv[1] if len(sys.argv) > 1 else "r01"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "smsp__average_warp_latency_issue_stalled_barrier.pct",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]
md = [f"# Round {R[1:]} ncu summaries (B200, `--clock-control none`)\n",
      "Produced by `tools/profile_round.sh` under gpurun and summarised by `tools/summarize_profiles.py`.  Times under ncu are "
      "cold-cache and serialised: compare SHARES, not absolutes; the bench numbers are in the BENCH lines / DESIGN.md.\n"]


def short(name):
    name = name.replace("void ", "").replace("b200::<unnamed>::", "")
    return name[:110]


try:                                                  # optional per-workload remarks: profiles/rNN_notes.json {workload: text}
    NOTES = json.load(open(os.path.join(P, f"{R}_notes.json")))
except Exception:
    NOTES = {}
for wl in ("bytetrack", "ocsort", "botsort", "deepocsort", "strongsort", "hybridsort", "ops"):
    src = os.path.join(G, f"{R}_launches_{wl}.csv")
    if not os.path.exists(src):
        continue
    shutil.copy(src, os.path.join(P, f"{R}_launches_{wl}.csv"))
    lines = [l for l in open(src) if l.startswith('"')]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    agg = defaultdict(lambda: [0, 0.0])
    unit = "ns"
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", unit)
        v_us = v / 1000.0 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000.0)
        a = agg[r["Kernel Name"]]
        a[0] += 1
        a[1] += v_us
    tot = sum(a[1] for a in agg.values()) or 1.0
    md.append(f"\n## Launch list: {wl} (`profiles/{R}_launches_{wl}.csv`)\n")
    if wl in NOTES:
        md.append(f"> NOTE: {NOTES[wl]}\n")
    md.append("| kernel | launches | mean us | share of GPU time |\n|---|---|---|---|")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]:
        md.append(f"| `{short(k)}` | {n} | {t / n:.1f} | {100 * t / tot:.2f} % |")

traffic = {}
try:
    traffic = json.load(open(os.path.join(P, "traffic.json")))
except Exception:
    pass
for cap in ("bytetrack", "ocsort", "botsort", "deepocsort", "strongsort", "hybridsort", "appearance", "gallery", "kf"):
    rep = os.path.join(G, f"{R}_full_{cap}.ncu-rep")
    raw = os.path.join(G, f"{R}_full_{cap}.raw.csv")          # exported on the GPU box by tools/profile_round.sh
    if os.path.exists(raw):
        txt = open(raw).read()
    elif os.path.exists(rep):
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    else:
        continue
    rows = list(csv.reader(io.StringIO(txt)))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    md.append(f"\n## `--set full` capture: {cap} (`gpurun_out/{R}_full_{cap}.ncu-rep`, not tracked; the raw-metric page is exported on the "
              f"GPU box, per-line source pages of the ByteTrack / OC-SORT / HybridSORT captures in `profiles/{R}_full_*.source.csv.gz`)\n")
    last = {}
    for row in rows[2:]:                                  # one entry per kernel: the last captured launch (after the warm-ups)
        last[dict(zip(hdr, row)).get('Kernel Name', '?')] = row
    for row in last.values():
        d = dict(zip(hdr, row))
        u = dict(zip(hdr, units))
        md.append(f"\n**`{short(d.get('Kernel Name', '?'))}`** grid {d.get('Grid Size', '?')} block {d.get('Block Size', '?')}\n")
        md.append("| metric | value |\n|---|---|")
        for k in KEYS:
            if k in d and d[k] != "":
                md.append(f"| {k} | {d[k]} {u.get(k, '')} |")
        kn = d.get("Kernel Name", "")
        try:
            rd = float(d["dram__bytes_read.sum"].replace(",", ""))
            wr = float(d["dram__bytes_write.sum"].replace(",", ""))
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            tot_b = rd * scale.get(u["dram__bytes_read.sum"], 1) + wr * scale.get(u["dram__bytes_write.sum"], 1)
            key = {"bytetrack": "bytetrack_step_kernel", "ocsort": "ocsort_step_kernel", "botsort": "bytetrack_step_kernel<BOT>",
                   "deepocsort": "deepocsort_step_kernel", "hybridsort": "hybridsort_step_kernel"}.get(cap)
            if cap in NOTES and "experiment" in NOTES[cap]:
                key = None                               # a capture of a variant that is not the kernel in the tree
            if key:
                traffic[key] = {"streams": int(d.get("Grid Size", "0").replace(",", "").split()[0].strip("()")) if d.get("Grid Size") else 0,
                                "dram_bytes_per_launch": int(tot_b),
                                "source": f"profiles/{R}_ncu_summary.md ({cap}: ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)"}
        except Exception as e:
            md.append(f"(traffic not parsed: {e})")
open(os.path.join(P, f"{R}_ncu_summary.md"), "w").write("\n".join(md) + "\n")
json.dump(traffic, open(os.path.join(P, "traffic.json"), "w"), indent=1)
print("\n".join(md))
