#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_fold_kernel -c 3 -f -o $O/prof_fold python tools/profile_step.py --stage v2v --repeat 1 > $O/ncu_fold.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:unproject -c 1 -f -o $O/prof_unproject2 python tools/profile_step.py --stage post --repeat 1 > $O/ncu_unproject2.log 2>&1
LT_BENCH_TIMELINE=$O/timeline_tc6.json timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee $O/bench_tc6.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['step_breakdown_ms'], d['clocks'])"
