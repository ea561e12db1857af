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
}
LL="--metrics gpu__time_duration.sum --clock-control none -c 400 --csv"
BT="python bench.py --steps 20 --warmup 5 --no-cpu-baseline"
OC="python bench.py --workload ocsort --streams 1024 --steps 20 --warmup 5 --no-cpu-baseline"
BS="python bench.py --workload botsort --streams 512 --steps 20 --warmup 5 --no-cpu-baseline"
OP="python tools/bench_ops.py --iters 3 --gallery-streams 64"
run ll_bytetrack $LL --log-file $O/${R}_launches_bytetrack.csv -- $BT
run ll_ocsort $LL --log-file $O/${R}_launches_ocsort.csv -- $OC
run ll_botsort $LL --log-file $O/${R}_launches_botsort.csv -- $BS
run ll_ops $LL --log-file $O/${R}_launches_ops.csv -- $OP
FULL="--set full --clock-control none --import-source on"
run full_bytetrack $FULL -k regex:bytetrack_step -s 12 -c 1 -o $O/${R}_full_bytetrack -f -- $BT
run full_ocsort $FULL -k regex:ocsort_step -s 12 -c 1 -o $O/${R}_full_ocsort -f -- $OC
run full_botsort $FULL -k regex:bytetrack_step -s 12 -c 1 -o $O/${R}_full_botsort -f -- $BS
run full_appearance $FULL -k regex:appearance_cost -s 1 -c 1 -o $O/${R}_full_appearance -f -- $OP --only appearance
run full_kf $FULL -k regex:"kf_(predict|update|project|gating)" -c 24 -o $O/${R}_full_kf -f -- python tools/bench_ops.py --iters 1 --only kf_predict,kf_update,kf_project,gating
run full_gallery $FULL -k regex:gallery_cost -s 1 -c 1 -o $O/${R}_full_gallery -f -- $OP --only gallery --gallery-streams 64 --kf-tracks 1000 --streams 8
ls -la $O | grep ${R}_
