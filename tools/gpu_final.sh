#!/bin/bash
# Round-1 closing run: diagnostics for the conv kernels (what bounds them), then the validation artefacts.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
T=${TAG:-r01d}
echo "== fold diagnostics (LT_FOLD_DBG 0/1/2/3, wait counters)"
timeout 300 python tools/fold_probe.py 2>&1 | tee $O/${T}_fold_probe.log | cut -c1-400
echo "== pytest gpu"
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider --tb=short > $O/final_pytest_gpu.log 2>&1; tail -3 $O/final_pytest_gpu.log
echo "== smoke"
timeout 600 python __graft_entry__.py --smoke 2>&1 | tail -3 | tee $O/final_smoke.log
show='import sys,json; d=json.loads(sys.stdin.read()); print(round(d["value"],1), round(d["e2e"]["value"],1), d["step_breakdown_ms"], {k: round(d[k]["frac"],3) for k in d if k.startswith("roofline")}, d["clocks"], d.get("cpu_baseline"))'
echo "== bench"
LT_BENCH_TIMELINE=$O/${T}_timeline_tc.json timeout 900 python bench.py --steps 10 --warmup 3 2> $O/${T}_bench_tc.err | tail -1 | tee $O/${T}_bench_tc.json | python -c "$show"
echo "== bench reference arm"
timeout 300 python bench.py --impl reference --steps 1 --warmup 1 2>/dev/null | tail -1 | tee $O/${T}_bench_reference.json | cut -c1-260
echo "== bench tc1 (one product instead of three: same operand traffic, a third of the MMA work)"
LT_BENCH_TIMELINE=$O/${T}_timeline_tc1.json timeout 900 python bench.py --mode tc1 --steps 6 --warmup 3 --no-cpu-baseline 2> $O/${T}_bench_tc1.err | tail -1 | tee $O/${T}_bench_tc1.json | python -c "$show"
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/${T}_launches_tc.csv python tools/profile_step.py --mode tc --repeat 2 > $O/ncu_launches.log 2>&1; wc -l $O/${T}_launches_tc.csv
echo "== ncu full"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stream_normalize -c 1 -f -o $O/${T}_prof_normalize python tools/profile_step.py --stage v2v --repeat 1 > $O/ncu_normalize.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:unproject -c 1 -f -o $O/${T}_prof_unproject python tools/profile_step.py --stage post --repeat 1 > $O/ncu_unproject.log 2>&1
ls -la $O/*.ncu-rep
echo "== done"
