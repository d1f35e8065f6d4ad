#!/bin/bash
# Measurement run: bench line (+ per-launch timeline), ncu launch list with DRAM bytes, ncu --set full captures.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
MODE=${MODE:-tc}
TAG=${TAG:-r01b}
LT_BENCH_TIMELINE=$O/${TAG}_timeline_$MODE.json timeout 900 python bench.py --mode $MODE --steps 10 --warmup 3 2>&1 | tail -1 | tee $O/${TAG}_bench_$MODE.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | tee $O/${TAG}_bench_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/${TAG}_launches_$MODE.csv python tools/profile_step.py --mode $MODE --repeat 2 > $O/ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:unproject -c 1 -f -o $O/${TAG}_prof_unproject python tools/profile_step.py --stage post --repeat 1 > $O/ncu_unproject.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:softargmax -c 3 -f -o $O/${TAG}_prof_softargmax python tools/profile_step.py --stage v2v --repeat 1 --mode $MODE > $O/ncu_softargmax.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_fold_kernel -c 2 -f -o $O/${TAG}_prof_conv_fold python tools/profile_step.py --stage v2v --repeat 1 --mode $MODE > $O/ncu_conv_fold.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 60 -c 4 -f -o $O/${TAG}_prof_conv_tc python tools/profile_step.py --stage all --repeat 1 --mode $MODE > $O/ncu_conv_tc.log 2>&1
ls -la $O | tail -12
