#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
echo "== fold tests"; timeout 600 python -m pytest tests/test_gpu_tc.py -q -m gpu -p no:cacheprovider --tb=short -s -k fold > $O/test_fold.log 2>&1; grep -E "passed|failed|conv_fold|Error|error|assert" $O/test_fold.log | tail -30
echo "== tc tests"; timeout 600 python -m pytest tests/test_gpu_tc.py -q -m gpu -p no:cacheprovider --tb=line -k "not fold" > $O/test_tc.log 2>&1; tail -2 $O/test_tc.log
echo "== forward"; timeout 1200 python -m pytest tests/test_gpu_forward.py -q -m gpu -s -p no:cacheprovider --tb=line > $O/test_forward.log 2>&1; grep -E "rel err|config2|passed|failed" $O/test_forward.log | tail -14
echo "== bench"; LT_BENCH_TIMELINE=$O/timeline_tc4.json timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee $O/bench_tc4.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['step_breakdown_ms'], d['roofline'])"
