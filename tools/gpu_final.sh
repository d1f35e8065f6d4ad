#!/bin/bash
# round-closing single-GPU run: full GPU suite, smoke, the default bench line (wall-clocked), the reference arm, one training step
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
T=${TAG:-r02m}
TAG=$T STAGES="all smoke" bash tools/gpu_r2.sh
S0=$(date +%s)
LT_BENCH_TIMELINE=$O/${T}_timeline_tc.json timeout 1500 python bench.py 2> $O/${T}_bench_full.err | tail -1 > $O/${T}_bench_full.json
echo "default bench.py wall seconds: $(( $(date +%s) - S0 ))" | tee $O/${T}_bench_full.wall
python -c "
import json
d=json.load(open('$O/${T}_bench_full.json'))
print(round(d['value'],1), round(d['e2e']['value'],1), round(d['ms_per_step'],3), d['step_breakdown_ms'], {k: round(d[k]['frac'],3) for k in d if k.startswith('roofline') and d[k]}, d['clocks'])
print('parity', d.get('parity')); print({k: (round(d[k]['value'],1), round(d[k]['native_over_this'],2)) for k in d if k.startswith('torch_gpu')}); print('config5', d.get('config5', {}).get('value'))
"
TAG=$T STAGES="trainstep" bash tools/gpu_r2.sh
