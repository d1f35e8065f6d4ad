#!/bin/bash
# Last GPU call of round 1: validate the issue-loop rewrites (conv_tc strength reduction, fold fast issue loop, B-resident
# 1x1 variant) with the whole GPU suite, then measure.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout 100 python -m pytest tests -q -m gpu -p no:cacheprovider --tb=line -x > $O/exp10_pytest_gpu.log 2>&1
ALL=$?; tail -4 $O/exp10_pytest_gpu.log | cut -c1-300
show='import sys,json; d=json.loads(sys.stdin.read()); print(round(d["value"],1), round(d["e2e"]["value"],1), d["step_breakdown_ms"], d["clocks"])'
LT_BENCH_TIMELINE=$O/exp10_timeline.json timeout 110 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2> $O/exp10_bench.err | tail -1 | tee $O/exp10_bench.json | python -c "$show"
if [ $ALL -ne 0 ]; then
  echo "-- suite with the new conv paths off"
  LT_TC_BRES=0 LT_FOLD_FAST_ISSUE=0 timeout 100 python -m pytest tests/test_gpu_tc.py -q -m gpu -p no:cacheprovider --tb=line -x 2>&1 | tail -3 | cut -c1-300
fi
LT_FOLD_FAST_ISSUE=0 timeout 60 python tools/fold_probe.py run 2>&1 | tail -1
timeout 60 python tools/fold_probe.py run 2>&1 | tail -1
echo "== done"
