#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
T=${TAG:-r02e}
timeout 900 python -m pytest tests/test_gpu_tc.py -q -m gpu -p no:cacheprovider --tb=short -x -k "fold" > $O/${T}_test_fold.log 2>&1; tail -5 $O/${T}_test_fold.log
for v in "1 1" "0 1" "1 0" "0 0"; do set -- $v; echo "fold_pair=$1 fold_direct=$2"; LT_OPT_FOLD_PAIR=$1 LT_OPT_FOLD_DIRECT=$2 python tools/fold_probe.py run; done 2>&1 | tee $O/${T}_fold_variants.log
LT_OPT_FOLD_DEBUG=16 python tools/fold_probe.py run 2>&1 | sort -u | tail -4 | tee -a $O/${T}_fold_variants.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_fold_kernel -s 3 -c 1 -f -o $O/${T}_src_fold_k3 python tools/fold_probe.py run > $O/ncu_src_fold3.log 2>&1
TAG=$T STAGES="bench" bash tools/gpu_r2.sh
