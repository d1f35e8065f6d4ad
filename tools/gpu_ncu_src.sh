#!/bin/bash
# ncu --set full with source-level stall samples for single layers (pair kernel 1x1 256->1024 +res at 24x24; kw-folded 3^3 and 7^3)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
T=${TAG:-r02c}
CONV_PROBE_PROF=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_pair_kernel -s 60 -c 1 -f -o $O/${T}_src_pair_256_1024 python tools/conv_probe.py 256,1024,1,1,24,32 > $O/ncu_src_pair.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_fold_kernel -s 3 -c 1 -f -o $O/${T}_src_fold_k3 python tools/fold_probe.py run > $O/ncu_src_fold3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_fold_kernel -s 9 -c 1 -f -o $O/${T}_src_fold_k7 python tools/fold_probe.py run > $O/ncu_src_fold7.log 2>&1
python tools/fold_probe.py 2>&1 | tee $O/${T}_fold_probe.log
ls -la $O/*.ncu-rep
