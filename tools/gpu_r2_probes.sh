#!/bin/bash
# First GPU call of round 2: the feed / pair probes that decide the conv_tc redesign.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
for p in cta2_probe tma_probe pair_probe; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/$p tools/$p.cu -lcuda > $O/${p}_build.log 2>&1 \
    && timeout 120 /tmp/$p > $O/$p.log 2>&1
  echo "$p exit $?"; grep -E "issue|ring|feed|b64|b128|a5d|cta_group|TIMEOUT|error" $O/$p.log | tail -40
done
echo "== hybrid (native backward kernels) gradient parity"
LT_TEST_HYBRID=1 timeout 600 python -m pytest tests/test_gpu_hybrid.py -q -m gpu -p no:cacheprovider --tb=short > $O/r2_hybrid.log 2>&1; tail -5 $O/r2_hybrid.log
echo "== secondary bar: the torch formulation through ATen/cuDNN on the same B200 (fp32, then TF32)"
timeout 600 python bench.py --impl torch_gpu --steps 5 --warmup 3 2>&1 | tail -1 | tee $O/r2_bench_torch_gpu_fp32.json | cut -c1-200
timeout 600 python bench.py --impl torch_gpu --tf32 --steps 5 --warmup 3 2>&1 | tail -1 | tee $O/r2_bench_torch_gpu_tf32.json | cut -c1-200
