#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
echo "== tc tests"; timeout 600 python -m pytest tests/test_gpu_tc.py -q -m gpu -p no:cacheprovider --tb=short -s > $O/test_tc.log 2>&1; grep -E "passed|failed|Error|error|splitk" $O/test_tc.log | tail -24
echo "== pipeline tests"; timeout 600 python -m pytest tests/test_gpu_pipeline.py -q -m gpu -p no:cacheprovider --tb=short > $O/test_pipe.log 2>&1; tail -15 $O/test_pipe.log
echo "== bench"; LT_BENCH_TIMELINE=$O/timeline_tc7.json timeout 900 python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee $O/bench_tc7.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['e2e']['sync_value'], d['step_breakdown_ms'])"
echo "== bench nosplit"; LT_TC_SPLITK=0 timeout 900 python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['e2e']['sync_value'], d['step_breakdown_ms'])"
echo "== forward"; timeout 1200 python -m pytest tests/test_gpu_forward.py -q -m gpu -s -p no:cacheprovider --tb=line > $O/test_forward.log 2>&1; grep -E "rel err|config2|passed|failed" $O/test_forward.log | tail -14
