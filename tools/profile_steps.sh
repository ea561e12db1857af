set -u
R=r01; O=gpurun_out
BT="python bench.py --steps 20 --warmup 5 --no-cpu-baseline"
OC="python bench.py --workload ocsort --streams 1024 --steps 20 --warmup 5 --no-cpu-baseline"
BS="python bench.py --workload botsort --streams 512 --steps 20 --warmup 5 --no-cpu-baseline"
LL="--metrics gpu__time_duration.sum --clock-control none -c 400 --csv"
FULL="--set full --clock-control none --import-source on"
ncu $LL --log-file $O/${R}_launches_bytetrack.csv $BT > $O/${R}_ll_bytetrack_ncu.log 2>&1
ncu $LL --log-file $O/${R}_launches_ocsort.csv $OC > $O/${R}_ll_ocsort_ncu.log 2>&1
ncu $LL --log-file $O/${R}_launches_botsort.csv $BS > $O/${R}_ll_botsort_ncu.log 2>&1
ncu $FULL -k regex:bytetrack_step -s 12 -c 1 -o $O/${R}_full_bytetrack -f $BT > $O/${R}_full_bytetrack_ncu.log 2>&1
ncu $FULL -k regex:ocsort_step -s 12 -c 1 -o $O/${R}_full_ocsort -f $OC > $O/${R}_full_ocsort_ncu.log 2>&1
ncu $FULL -k regex:bytetrack_step -s 12 -c 1 -o $O/${R}_full_botsort -f $BS > $O/${R}_full_botsort_ncu.log 2>&1
python bench.py > $O/bench_r1_final_bt.json 2>/dev/null
python bench.py --workload ocsort > $O/bench_r1_final_oc.json 2>/dev/null
python bench.py --workload ocsort --streams 2048 --no-cpu-baseline > $O/bench_r1_final_oc2048.json 2>/dev/null
python bench.py --workload botsort > $O/bench_r1_final_bs.json 2>/dev/null
python bench.py --workload botsort --streams 1024 --no-cpu-baseline > $O/bench_r1_final_bs1024.json 2>/dev/null
python tools/bench_ops.py > $O/r01_ops.jsonl 2>/dev/null
ls $O | grep -c r01_
