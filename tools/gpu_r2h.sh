#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
T=${TAG:-r02h}
timeout 900 python -m pytest tests/test_gpu_tc.py -q -m gpu -p no:cacheprovider --tb=short -x -k "pair or tail" > $O/${T}_test_pair.log 2>&1; tail -3 $O/${T}_test_pair.log
for cfg in "0 0" "1 0" "1 1" "1 2" "2 0"; do set -- $cfg
  echo "== pair_two_acc=$1 pair_long_k=$2"
  LT_OPT_PAIR_TWO_ACC=$1 LT_OPT_PAIR_LONG_K=$2 python tools/precision_probe.py --batch 2 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('   tc vs simt', d['tc vs simt']); print('   tc vs oracle', d['tc vs cpu_oracle'])"
  LT_OPT_PAIR_TWO_ACC=$1 LT_OPT_PAIR_LONG_K=$2 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-torch-gpu --no-config5 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('   bench', round(d['value'],1), round(d['ms_per_step'],3), d['step_breakdown_ms'])"
done 2>&1 | tee $O/${T}_two_acc_ab.log
