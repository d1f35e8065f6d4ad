#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
echo "== tc tests"; timeout 600 python -m pytest tests/test_gpu_tc.py -q -m gpu -p no:cacheprovider --tb=short -s > $O/test_tc.log 2>&1; grep -E "passed|failed|Error|error" $O/test_tc.log | tail -8
echo "== bench wx16"; LT_FOLD_WX=16 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['step_breakdown_ms'])"
echo "== bench auto"; LT_BENCH_TIMELINE=$O/timeline_tc6.json timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee $O/bench_tc6.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['step_breakdown_ms'])"
echo "== forward"; timeout 1200 python -m pytest tests/test_gpu_forward.py -q -m gpu -s -p no:cacheprovider --tb=line > $O/test_forward.log 2>&1; grep -E "rel err|config2|passed|failed" $O/test_forward.log | tail -14
