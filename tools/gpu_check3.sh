#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
echo "== tc tests, TMA epilogue, stride mode 0"; timeout 600 python -m pytest tests/test_gpu_tc.py -q -m gpu -p no:cacheprovider --tb=line -s > $O/test_tc_m0.log 2>&1; grep -E "passed|failed|stride2|Error|error" $O/test_tc_m0.log | tail -12
echo "== stride mode 1"; LT_TMA_STRIDE_MODE=1 timeout 600 python -m pytest tests/test_gpu_tc.py -q -m gpu -p no:cacheprovider --tb=line -s -k stride2 > $O/test_tc_m1.log 2>&1; grep -E "passed|failed|stride2|Error|error" $O/test_tc_m1.log | tail -8
echo "== direct epilogue control"; LT_TC_EPILOGUE=direct timeout 600 python -m pytest tests/test_gpu_tc.py -q -m gpu -p no:cacheprovider --tb=line -k "not stride2" > $O/test_tc_direct.log 2>&1; tail -2 $O/test_tc_direct.log
echo "== ops"; timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu -p no:cacheprovider --tb=line > $O/test_ops.log 2>&1; tail -2 $O/test_ops.log
echo "== forward"; timeout 1200 python -m pytest tests/test_gpu_forward.py -q -m gpu -s -p no:cacheprovider --tb=line > $O/test_forward.log 2>&1; grep -E "rel err|config2|passed|failed" $O/test_forward.log | tail -14
echo "== bench"; LT_BENCH_TIMELINE=$O/timeline_tc3.json timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee $O/bench_tc3.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['step_breakdown_ms'], d.get('roofline_unproject'), d.get('roofline_softargmax'))"
