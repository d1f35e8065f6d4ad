#!/bin/bash
# round-2 measurement run: default bench line (wall-clocked), ncu launch list, ncu --set full captures of every hot kernel
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
T=${TAG:-r02g}
S0=$(date +%s)
LT_BENCH_TIMELINE=$O/${T}_timeline_tc.json timeout 1500 python bench.py 2> $O/${T}_bench_full.err | tail -1 > $O/${T}_bench_full.json
echo "default bench.py wall seconds: $(( $(date +%s) - S0 ))" | tee $O/${T}_bench_full.wall
python -c "
import json,sys
d=json.load(open('$O/${T}_bench_full.json'))
print(round(d['value'],1), round(d['e2e']['value'],1), round(d['ms_per_step'],3), d['step_breakdown_ms'], {k: round(d[k]['frac'],3) for k in d if k.startswith('roofline') and d[k]}, d['clocks'])
print('parity', d.get('parity')); print('cpu', d.get('cpu_baseline')); print({k: d[k] for k in d if k.startswith('torch_gpu')}); print('config5', d.get('config5'))
"
TAG=$T STAGES="launches" bash tools/gpu_r2.sh
T2=$T
timeout 600 ncu --set full --clock-control none --import-source on -k regex:unproject -c 1 -f -o $O/${T2}_prof_unproject python tools/profile_step.py --stage post --repeat 1 > $O/ncu_unproject.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"stream_|softargmax|v2v_tail" -c 3 -f -o $O/${T2}_prof_softargmax python tools/profile_step.py --stage v2v --repeat 1 > $O/ncu_softargmax.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv_fold_kernel" -c 3 -f -o $O/${T2}_prof_conv_fold python tools/profile_step.py --stage v2v --repeat 1 > $O/ncu_conv_fold.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_pair_kernel -s 40 -c 6 -f -o $O/${T2}_prof_conv_pair python tools/profile_step.py --stage all --repeat 1 > $O/ncu_conv_pair.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_tc_kernel|conv_tc_persist" -s 6 -c 4 -f -o $O/${T2}_prof_conv_tc python tools/profile_step.py --stage all --repeat 1 > $O/ncu_conv_tc.log 2>&1
ls -la $O/*.ncu-rep
