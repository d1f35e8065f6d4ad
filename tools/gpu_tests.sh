#!/bin/bash
# GPU unit + module parity tests, full logs under gpurun_out/.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu -p no:cacheprovider --tb=short > $O/test_ops.log 2>&1; tail -15 $O/test_ops.log
timeout 600 python -m pytest tests/test_gpu_tc.py -q -m gpu -p no:cacheprovider --tb=short > $O/test_tc.log 2>&1; tail -30 $O/test_tc.log
timeout 1200 python -m pytest tests/test_gpu_forward.py -q -m gpu -s -p no:cacheprovider --tb=line ${FWD_ARGS} > $O/test_forward.log 2>&1; grep -E "rel err|config2|passed|failed" $O/test_forward.log | tail -30
