#!/bin/bash
# N-GPU run (gpurun --gpus N): the 2-GPU view-sharded parity test, then the weak-scaling bench line (both exchanges measured).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=${N:-2}
T=${TAG:-r02}
if [ "${DIST_TEST:-1}" = "1" ]; then
  timeout 900 python -m pytest tests/test_gpu_dist.py -q -m gpu -s -p no:cacheprovider --tb=short > gpurun_out/${T}_test_dist_n${N}.log 2>&1; tail -6 gpurun_out/${T}_test_dist_n${N}.log
fi
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps ${STEPS:-8} --warmup 3 2> gpurun_out/${T}_bench_n${N}.err | tail -1 | tee gpurun_out/${T}_bench_tc_n${N}.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['n_gpus'], round(d['value'],1), round(d['e2e']['value'],1), round(d['ms_per_step'],3), d.get('exchanges'), d.get('config4'), d['config']['parallelism'])"
tail -5 gpurun_out/${T}_bench_n${N}.err
