#!/bin/bash
# Round-1 measurement run: bench line (+ per-launch timeline) and ncu evidence.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
MODE=${MODE:-tc}
LT_BENCH_TIMELINE=$O/timeline_$MODE.json timeout 900 python bench.py --mode $MODE --steps 10 --warmup 3 2>&1 | tail -3 | tee $O/bench_$MODE.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_$MODE.csv python tools/profile_step.py --mode $MODE --repeat 1 > $O/ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:unproject_kernel -c 1 -f -o $O/prof_unproject python tools/profile_step.py --stage post --repeat 1 > $O/ncu_unproject.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:softargmax -c 3 -f -o $O/prof_softargmax python tools/profile_step.py --stage v2v --repeat 1 --mode $MODE > $O/ncu_softargmax.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -c 4 -f -o $O/prof_conv_v2v python tools/profile_step.py --stage v2v --repeat 1 --mode $MODE > $O/ncu_conv_v2v.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 60 -c 4 -f -o $O/prof_conv_bb python tools/profile_step.py --stage all --repeat 1 --mode $MODE > $O/ncu_conv_bb.log 2>&1
ls -la $O
