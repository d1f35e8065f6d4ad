#!/bin/bash
# A/B run of a kernel-selection knob on one box:  gpurun -- 'KNOB="LT_TC_WIDE=1" TESTS="tests/test_gpu_tc.py" bash tools/gpu_ab.sh'
#   1. parity tests with the knob set, 2. bench with and without the knob (same box, back to back), 3. per-layer diff.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
KNOB=${KNOB:?set KNOB="NAME=value [NAME2=value2]"}
TESTS=${TESTS:-tests/test_gpu_tc.py tests/test_gpu_ops.py}
TAG=${TAG:-ab}
echo "== parity tests with $KNOB"
env $KNOB timeout 900 python -m pytest $TESTS -q -m gpu -p no:cacheprovider --tb=short > $O/${TAG}_tests.log 2>&1; RC=$?
tail -3 $O/${TAG}_tests.log; [ $RC -ne 0 ] && grep -E "^FAILED|^ERROR" $O/${TAG}_tests.log | head -20
show='import sys,json; d=json.loads(sys.stdin.read()); print(round(d["value"],1), round(d["e2e"]["value"],1), d["step_breakdown_ms"], d["clocks"])'
echo "== bench baseline"
LT_BENCH_TIMELINE=$O/${TAG}_timeline_base.json timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2> $O/${TAG}_base.err | tail -1 | tee $O/${TAG}_bench_base.json | python -c "$show"
echo "== bench with $KNOB"
env $KNOB LT_BENCH_TIMELINE=$O/${TAG}_timeline_knob.json timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2> $O/${TAG}_knob.err | tail -1 | tee $O/${TAG}_bench_knob.json | python -c "$show"
python - <<PY
import json, collections
def agg(p):
    a = collections.OrderedDict()
    for r in json.load(open(p)):
        v = a.setdefault((r["kernel"], r["desc"]), [0, 0.0]); v[0] += 1; v[1] += r["ms"]
    return a
a, b = agg("$O/${TAG}_timeline_base.json"), agg("$O/${TAG}_timeline_knob.json")
rows = sorted(((b[k][1] - v[1], k, v[0], v[1], b[k][1]) for k, v in a.items() if k in b))
print("per-layer (knob - base), ms over all launches of the layer; total %.3f -> %.3f" % (sum(r[3] for r in rows), sum(r[4] for r in rows)))
for d, k, n, o, nw in rows[:8] + rows[-8:]:
    print("%-10s %-46s n=%3d base=%7.3f knob=%7.3f d=%+.3f" % (k[0], k[1], n, o, nw, d))
PY
