#!/bin/bash
# First-contact GPU run: each stage in its own process with a timeout so that a trap in one kernel
# cannot poison the rest; everything lands under gpurun_out/.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi > $O/nvidia_smi.txt 2>&1
python -c "import torch;print(torch.__version__, torch.cuda.get_device_name(0))" > $O/env.txt 2>&1
echo "== ops ==";   timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu -p no:cacheprovider 2>&1 | tail -40 | tee $O/test_ops.log
echo "== selftest =="; timeout 300 python -m pytest tests/test_gpu_tc.py -q -m gpu -k selftest -p no:cacheprovider 2>&1 | tail -40 | tee $O/test_tc_selftest.log
echo "== tc conv =="; timeout 600 python -m pytest tests/test_gpu_tc.py -q -m gpu -k "not selftest" -p no:cacheprovider 2>&1 | tail -60 | tee $O/test_tc.log
echo "== forward =="; timeout 1200 python -m pytest tests/test_gpu_forward.py -q -m gpu -s -p no:cacheprovider 2>&1 | tail -60 | tee $O/test_forward.log
echo "== smoke =="; timeout 600 python __graft_entry__.py --smoke 2>&1 | tail -20 | tee $O/smoke.log
echo "== bench simt =="; timeout 900 python bench.py --mode simt --steps 3 --warmup 3 2>&1 | tail -5 | tee $O/bench_simt.log
echo "== bench tc =="; timeout 900 python bench.py --mode tc --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -5 | tee $O/bench_tc.log
