#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu -p no:cacheprovider --tb=short -x > $O/test_ops.log 2>&1; tail -5 $O/test_ops.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 2>&1 | tail -3 | tee $O/bench_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --collective reduce_scatter 2>&1 | tail -3 | tee $O/bench_2gpu_rs.log
LT_BENCH_TIMELINE=$O/timeline_tc2.json timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -2 | tee $O/bench_tc2.log
