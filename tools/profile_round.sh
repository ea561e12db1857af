#!/bin/bash
# ncu evidence for one round (run under gpurun on one B200):  bash tools/profile_round.sh rNN
# Launch lists (gpu__time_duration.sum, --clock-control none) of the three workloads and the operator bench, then one
# --set full capture of every top kernel.  Each command runs plain first (must exit 0) before it runs under ncu.
set -u
R=${1:-r01}
O=gpurun_out
mkdir -p $O
run() {   # name, ncu args..., -- , command...
    local name=$1; shift
    local ncu_args=()
    while [ "$1" != "--" ]; do ncu_args+=("$1"); shift; done
    shift
    "$@" > $O/${R}_${name}_plain.log 2>&1 || { echo "$name: plain run failed"; tail -5 $O/${R}_${name}_plain.log; return 1; }
    ncu "${ncu_args[@]}" "$@" > $O/${R}_${name}_ncu.log 2>&1 || { echo "$name: ncu failed"; tail -5 $O/${R}_${name}_ncu.log; }
    # gpurun brings back at most 64 MiB: a --set full report is exported to its raw-metric page (what
    # tools/summarize_profiles.py reads) and the per-line source page of the step kernels (tools/ncu_lines.py) on the box,
    # and the .ncu-rep itself is dropped unless KEEP_REP=1
    local rep=$O/${R}_${name#full_}.ncu-rep
    [ "${name#full_}" != "$name" ] && rep=$O/${R}_${name}.ncu-rep
    if [ -f "$rep" ]; then
        ncu -i "$rep" --page raw --csv > "${rep%.ncu-rep}.raw.csv" 2>/dev/null
        case "$name" in full_bytetrack|full_ocsort|full_botsort|full_deepocsort|full_hybridsort) ncu -i "$rep" --page source --csv > "${rep%.ncu-rep}.source.csv" 2>/dev/null; gzip -f "${rep%.ncu-rep}.source.csv";; esac
        [ -z "${KEEP_REP:-}" ] && rm -f "$rep"
    fi
}
LL="--metrics gpu__time_duration.sum --clock-control none -c 400 --csv"
BT="python bench.py --steps 20 --warmup 5 --no-cpu-baseline"
OC="python bench.py --workload ocsort --streams 2048 --steps 20 --warmup 5 --no-cpu-baseline"
BS="python bench.py --workload botsort --streams 1024 --steps 20 --warmup 5 --no-cpu-baseline"
DO="python bench.py --workload deepocsort --streams 256 --steps 5 --warmup 3 --no-cpu-baseline"
SS="python bench.py --workload strongsort --streams 64 --steps 3 --warmup 3 --no-cpu-baseline"
HY="python bench.py --workload hybridsort --streams 256 --steps 5 --warmup 3 --no-cpu-baseline"
OP="python tools/bench_ops.py --iters 3 --gallery-streams 64"
run ll_bytetrack $LL --log-file $O/${R}_launches_bytetrack.csv -- $BT
run ll_ocsort $LL --log-file $O/${R}_launches_ocsort.csv -- $OC
run ll_botsort $LL --log-file $O/${R}_launches_botsort.csv -- $BS
run ll_ops $LL --log-file $O/${R}_launches_ops.csv -- $OP
LLT="--metrics gpu__time_duration.sum --clock-control none -s 330 -c 160 --csv"     # skip the pre-roll's launches
run ll_deepocsort $LL --log-file $O/${R}_launches_deepocsort.csv -- $DO
run ll_strongsort $LLT --log-file $O/${R}_launches_strongsort.csv -- $SS
run ll_hybridsort $LL --log-file $O/${R}_launches_hybridsort.csv -- $HY
if [ "${2:-}" = "lists" ]; then ls -la $O | grep ${R}_; exit 0; fi
FULL="--set full --clock-control none --import-source on"
if [ -z "${SKIP_BT_OC:-}" ]; then      # SKIP_BT_OC=1: keep earlier captures of kernels that did not change
run full_bytetrack $FULL -k regex:bytetrack_step -s 12 -c 1 -o $O/${R}_full_bytetrack -f -- $BT
run full_ocsort $FULL -k regex:ocsort_step -s 12 -c 1 -o $O/${R}_full_ocsort -f -- $OC
fi
run full_botsort $FULL -k regex:bytetrack_step -s 12 -c 1 -o $O/${R}_full_botsort -f -- $BS
run full_deepocsort $FULL -k regex:deepocsort_step -s 12 -c 1 -o $O/${R}_full_deepocsort -f -- $DO
run full_strongsort $FULL -k regex:"gallery_cost|ss_match|ss_post" -s 126 -c 3 -o $O/${R}_full_strongsort -f -- $SS
run full_hybridsort $FULL -k regex:hybridsort_step -s 44 -c 1 -o $O/${R}_full_hybridsort -f -- $HY
if [ "${2:-}" = "steps" ]; then ls -la $O | grep ${R}_; exit 0; fi
run full_appearance $FULL -k regex:appearance_cost -s 1 -c 1 -o $O/${R}_full_appearance -f -- $OP --only appearance
run full_kf $FULL -k regex:"kf_(predict|update|project|gating)" -c 24 -o $O/${R}_full_kf -f -- python tools/bench_ops.py --iters 1 --only kf_predict,kf_update,kf_project,gating
run full_gallery $FULL -k regex:gallery_cost -s 1 -c 1 -o $O/${R}_full_gallery -f -- $OP --only gallery --gallery-streams 64 --kf-tracks 1000 --streams 8
ls -la $O | grep ${R}_
