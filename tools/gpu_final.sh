#!/bin/bash
# Round-end validation: full GPU suite, smoke, bench (both arms), with logs under gpurun_out/.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider --tb=short > $O/final_pytest_gpu.log 2>&1; tail -4 $O/final_pytest_gpu.log
timeout 600 python __graft_entry__.py --smoke 2>&1 | tail -3 | tee $O/final_smoke.log
LT_BENCH_TIMELINE=$O/r01c_timeline_tc.json timeout 900 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 | tee $O/r01c_bench_tc.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['step_breakdown_ms'], d['cpu_baseline'])"
timeout 300 python bench.py --impl reference --steps 1 --warmup 1 2>&1 | tail -1 | tee $O/r01c_bench_reference.json | cut -c1-300
