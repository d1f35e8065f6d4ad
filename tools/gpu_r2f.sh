#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
T=${TAG:-r02f}
timeout 900 python -m pytest tests/test_gpu_tc.py -q -m gpu -p no:cacheprovider --tb=short -x -s -k "tail" > $O/${T}_test_tail.log 2>&1; grep -E "statistics|passed|failed|Error|error" $O/${T}_test_tail.log | tail -20
TAG=$T STAGES="all" bash tools/gpu_r2.sh
/usr/bin/time -v -o $O/${T}_benchfull.time bash -c "TAG=$T STAGES=benchfull bash tools/gpu_r2.sh"; grep -E "Elapsed" $O/${T}_benchfull.time
TAG=$T STAGES="benchref smoke" bash tools/gpu_r2.sh
