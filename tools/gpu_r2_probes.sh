#!/bin/bash
# First GPU call of round 2: the feed / pair probes that decide the conv_tc redesign.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
for p in tma_probe pair_probe; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/$p tools/$p.cu -lcuda > $O/${p}_build.log 2>&1 \
    && timeout 120 /tmp/$p > $O/$p.log 2>&1
  echo "$p exit $?"; cat $O/$p.log | head -40
done
