#!/bin/bash
# Experiments 9: probe v2 (product MMA pattern with distinct operands, L2 feed rate), streaming soft-argmax rewrite,
# unproject register variants; then the full suite + bench with the new defaults.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
echo "== probe"
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/cta2_probe tools/cta2_probe.cu > $O/cta2_probe_build.log 2>&1 \
  && timeout 180 /tmp/cta2_probe > $O/cta2_probe2.log 2>&1
echo "probe exit $?"; grep -E "^ring|^feed|TIMEOUT|error" $O/cta2_probe2.log | head -40
echo "== ops tests"
timeout 900 python -m pytest tests/test_gpu_ops.py -q -m gpu -p no:cacheprovider --tb=short > $O/exp9_ops.log 2>&1
OPS=$?; tail -3 $O/exp9_ops.log; [ $OPS -ne 0 ] && grep -E "^FAILED|^ERROR|Error" $O/exp9_ops.log | head -20
echo "== post probe"
timeout 600 python tools/post_probe.py 2>&1 | tee $O/exp9_post_probe.log
echo "== full GPU suite"
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider --tb=short > $O/exp9_pytest_gpu.log 2>&1
ALL=$?; tail -3 $O/exp9_pytest_gpu.log; [ $ALL -ne 0 ] && grep -E "^FAILED|^ERROR" $O/exp9_pytest_gpu.log | head -20
show='import sys,json; d=json.loads(sys.stdin.read()); print(round(d["value"],1), round(d["e2e"]["value"],1), d["step_breakdown_ms"], {k: round(d[k]["frac"],3) for k in d if k.startswith("roofline_")}, d["gpu_launches"])'
echo "== bench"
LT_BENCH_TIMELINE=$O/exp9_timeline.json timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2> $O/exp9_bench.err | tail -1 | tee $O/exp9_bench.json | python -c "$show"
if [ $OPS -eq 0 ]; then
  echo "== ncu full: soft-argmax"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:stream_ -c 2 -f -o $O/r01d_prof_softargmax python tools/profile_step.py --stage v2v --repeat 1 > $O/ncu_softargmax.log 2>&1
  ls -la $O/r01d_prof_softargmax.ncu-rep
fi
echo "== done"
