#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
T=${TAG:-r02j}
for comp in 1 0; do
  echo "== LT_TC_ACCUM_COMP=$comp"
  LT_TC_ACCUM_COMP=$comp python tools/precision_probe.py --batch 2 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('   B=2 tc vs simt', d['tc vs simt']); print('   B=2 tc vs oracle', d['tc vs cpu_oracle'])"
done 2>&1 | tee $O/${T}_accum_comp.log
LT_BENCH_TIMELINE=$O/${T}_timeline_tc.json python bench.py --steps 20 --warmup 3 --no-torch-gpu 2>/dev/null | tail -1 > $O/${T}_bench.json
python -c "
import json
d=json.load(open('$O/${T}_bench.json')); print('   bench', round(d['value'],1), round(d['e2e']['value'],1), round(d['ms_per_step'],3), d['step_breakdown_ms']); print('   parity', d['parity']); print('   config5', d.get('config5'))" | tee -a $O/${T}_accum_comp.log
TAG=$T STAGES="all" bash tools/gpu_r2.sh
