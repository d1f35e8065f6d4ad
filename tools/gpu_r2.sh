#!/bin/bash
# Round-2 GPU session script: STAGES selects what runs (space separated): tc ops forward hybrid rest bench benchfull smoke
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
T=${TAG:-r02a}
STAGES=${STAGES:-"tc forward bench"}
show='import sys,json; d=json.loads(sys.stdin.read()); print(round(d["value"],1), round(d["e2e"]["value"],1), round(d["ms_per_step"],3), d["step_breakdown_ms"], {k: round(d[k]["frac"],3) for k in d if k.startswith("roofline") and d[k]}, d["clocks"], d.get("parity"), {k: d[k] for k in d if k.startswith("torch_gpu")})'
for s in $STAGES; do
  case $s in
    tc) timeout 900 python -m pytest tests/test_gpu_tc.py -q -m gpu -p no:cacheprovider --tb=short -x -s > $O/${T}_test_tc.log 2>&1; grep -E "conv_pair|passed|failed|Error|error" $O/${T}_test_tc.log | tail -40 ;;
    ops) timeout 900 python -m pytest tests/test_gpu_ops.py -q -m gpu -p no:cacheprovider --tb=short > $O/${T}_test_ops.log 2>&1; tail -5 $O/${T}_test_ops.log ;;
    forward) timeout 1500 python -m pytest tests/test_gpu_forward.py -q -m gpu -s -p no:cacheprovider --tb=short > $O/${T}_test_forward.log 2>&1; grep -E "rel err|config2|passed|failed|Error" $O/${T}_test_forward.log | tail -30 ;;
    hybrid) timeout 900 python -m pytest tests/test_gpu_hybrid.py -q -m gpu -p no:cacheprovider --tb=short > $O/${T}_test_hybrid.log 2>&1; tail -5 $O/${T}_test_hybrid.log ;;
    rest) timeout 1500 python -m pytest tests/test_gpu_algebraic.py tests/test_gpu_pipeline.py tests/test_gpu_variants.py -q -m gpu -p no:cacheprovider --tb=short > $O/${T}_test_rest.log 2>&1; tail -8 $O/${T}_test_rest.log ;;
    all) timeout 2400 python -m pytest tests -q -m gpu -p no:cacheprovider --tb=short > $O/${T}_pytest_gpu.log 2>&1; tail -8 $O/${T}_pytest_gpu.log ;;
    bench) LT_BENCH_TIMELINE=$O/${T}_timeline_tc.json timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-torch-gpu 2> $O/${T}_bench_tc.err | tail -1 | tee $O/${T}_bench_tc.json | python -c "$show" ;;
    benchfull) LT_BENCH_TIMELINE=$O/${T}_timeline_tc.json timeout 1200 python bench.py --steps 20 --warmup 3 2> $O/${T}_bench_tc.err | tail -1 | tee $O/${T}_bench_tc.json | python -c "$show" ;;
    benchref) timeout 600 python bench.py --impl reference --steps 1 --warmup 1 2>/dev/null | tail -1 | tee $O/${T}_bench_reference.json | cut -c1-300 ;;
    smoke) timeout 600 python __graft_entry__.py --smoke 2>&1 | tail -4 | tee $O/${T}_smoke.log ;;
    launches) timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/${T}_launches_tc.csv python tools/profile_step.py --mode tc --repeat 2 > $O/ncu_launches.log 2>&1; wc -l $O/${T}_launches_tc.csv ;;
    ncufull)
      timeout 600 ncu --set full --clock-control none --import-source on -k regex:unproject -c 1 -f -o $O/${T}_prof_unproject python tools/profile_step.py --stage post --repeat 1 > $O/ncu_unproject.log 2>&1
      timeout 600 ncu --set full --clock-control none --import-source on -k regex:"stream_|softargmax" -c 3 -f -o $O/${T}_prof_softargmax python tools/profile_step.py --stage v2v --repeat 1 > $O/ncu_softargmax.log 2>&1
      timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv_fold_kernel|v2v_tail" -c 3 -f -o $O/${T}_prof_conv_fold python tools/profile_step.py --stage v2v --repeat 1 > $O/ncu_conv_fold.log 2>&1
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_pair_kernel -s 40 -c 4 -f -o $O/${T}_prof_conv_pair python tools/profile_step.py --stage all --repeat 1 > $O/ncu_conv_pair.log 2>&1
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 20 -c 3 -f -o $O/${T}_prof_conv_tc python tools/profile_step.py --stage all --repeat 1 > $O/ncu_conv_tc.log 2>&1
      ls -la $O/*.ncu-rep ;;
    trainstep) timeout 900 python tools/train_step_bench.py --batch 4 --steps 4 2>&1 | tail -1 | tee $O/${T}_train_step.json | cut -c1-600 ;;
    convprobe) timeout 600 python tools/conv_probe.py ${PROBE_CASES} 2>&1 | tee $O/${T}_conv_probe.log | cut -c1-900 ;;
    ncupair)
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_pair_kernel -s 44 -c 3 -f -o $O/${T}_prof_conv_pair python tools/profile_step.py --stage all --repeat 1 > $O/ncu_conv_pair.log 2>&1
      ls -la $O/*.ncu-rep ;;
    postprobe) timeout 600 python tools/post_probe.py 2>&1 | tee $O/${T}_post_probe.log | cut -c1-300 ;;
    nopair) LT_TC_PAIR=0 LT_BENCH_TIMELINE=$O/${T}_timeline_nopair.json timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-torch-gpu 2> $O/${T}_bench_nopair.err | tail -1 | tee $O/${T}_bench_nopair.json | python -c "$show" ;;
  esac
  echo "== stage $s done"
done
