#!/bin/bash
# N-GPU weak-scaling bench (view-sharded), to be run with gpurun --gpus N
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=${N:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 8 --warmup 3 --no-cpu-baseline --collective ${COLL:-all_reduce} 2>&1 | tail -1 | tee gpurun_out/r01b_bench_tc_n${N}_${COLL:-all_reduce}.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['value'], d['e2e']['value'], d['ms_per_step'], d['config']['parallelism'])"
