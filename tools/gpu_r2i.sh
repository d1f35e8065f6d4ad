#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
T=${TAG:-r02i}
for cfg in "1 1" "2 1" "0 0"; do set -- $cfg
  echo "== pair_two_acc=$1 pair_long_k=$2"
  LT_BENCH_TIMELINE=$O/${T}_timeline_$1$2.json LT_OPT_PAIR_TWO_ACC=$1 LT_OPT_PAIR_LONG_K=$2 python bench.py --steps 20 --warmup 3 --no-torch-gpu --no-config5 2>/dev/null | tail -1 > $O/${T}_bench_$1$2.json
  python -c "
import json
d=json.load(open('$O/${T}_bench_$1$2.json')); print('   bench', round(d['value'],1), round(d['ms_per_step'],3), d['step_breakdown_ms']); print('   parity', d['parity'])"
done 2>&1 | tee $O/${T}_two_acc_b8.log
