#!/bin/bash
# Experiments: 2-CTA tcgen05 probe, fused soft-argmax + compact logits + unproject v2 (parity, then timing vs the
# previous kernels), and -- only if the whole GPU suite is green with the new defaults -- the ncu evidence for them.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L | head -2
echo "== cta2 probe"
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/cta2_probe tools/cta2_probe.cu > $O/cta2_probe_build.log 2>&1 \
  && timeout 120 /tmp/cta2_probe > $O/cta2_probe.log 2>&1
echo "probe exit $?"; cat $O/cta2_probe.log | head -60

echo "== ops tests (new defaults)"
timeout 900 python -m pytest tests/test_gpu_ops.py -q -m gpu -p no:cacheprovider --tb=short > $O/exp8_ops.log 2>&1
OPS=$?; tail -4 $O/exp8_ops.log
if [ $OPS -ne 0 ]; then
  grep -E "^FAILED|^ERROR" $O/exp8_ops.log | head -20
  for v in "LT_SOFTARGMAX_FUSED=0" "LT_UNPROJECT_V2=0"; do
    echo "-- ops tests with $v"; env $v timeout 900 python -m pytest tests/test_gpu_ops.py -q -m gpu -p no:cacheprovider --tb=line 2>&1 | tail -3
  done
fi

echo "== full GPU suite (new defaults)"
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider --tb=short > $O/exp8_pytest_gpu.log 2>&1
ALL=$?; tail -4 $O/exp8_pytest_gpu.log
if [ $ALL -ne 0 ]; then
  grep -E "^FAILED|^ERROR" $O/exp8_pytest_gpu.log | head -20
  echo "-- suite with compact logits off"; LT_LOGITS_COMPACT=0 timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider --tb=line -x 2>&1 | tail -3
  echo "-- suite with everything new off"; LT_LOGITS_COMPACT=0 LT_SOFTARGMAX_FUSED=0 LT_UNPROJECT_V2=0 timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider --tb=line -x 2>&1 | tail -3
fi

show='import sys,json; d=json.loads(sys.stdin.read()); print(round(d["value"],1), round(d["e2e"]["value"],1), d["step_breakdown_ms"], {k: round(d[k]["frac"],3) for k in d if k.startswith("roofline_")}, d["gpu_launches"])'
echo "== bench new defaults"
LT_BENCH_TIMELINE=$O/exp8_timeline_new.json timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2> $O/exp8_bench_new.err | tail -1 | tee $O/exp8_bench_new.json | python -c "$show"
echo "== bench previous kernels"
LT_LOGITS_COMPACT=0 LT_SOFTARGMAX_FUSED=0 LT_UNPROJECT_V2=0 timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2> $O/exp8_bench_old.err | tail -1 | tee $O/exp8_bench_old.json | python -c "$show"
echo "== bench fused soft-argmax on 32-wide logits"
LT_LOGITS_COMPACT=0 timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2> $O/exp8_bench_wide.err | tail -1 | tee $O/exp8_bench_wide.json | python -c "$show"

if [ $ALL -eq 0 ]; then
  echo "== ncu full: soft-argmax / unproject"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:softargmax -c 1 -f -o $O/r01d_prof_softargmax python tools/profile_step.py --stage v2v --repeat 1 > $O/ncu_softargmax.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:unproject -c 1 -f -o $O/r01d_prof_unproject python tools/profile_step.py --stage post --repeat 1 > $O/ncu_unproject.log 2>&1
  ls -la $O/*.ncu-rep 2>/dev/null
fi
echo "== done"
