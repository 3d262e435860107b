#!/bin/bash
# Round 2, first GPU call: smoke, full gpu test suite, bench (both arms).
TAG=${1:-r2a}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $OUT/gpu.txt 2>&1
nvidia-smi topo -m > $OUT/topo.txt 2>&1
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $OUT/smoke.log 2>&1
echo "smoke rc=$?" | tee -a $OUT/summary.txt
tail -8 $OUT/smoke.log
timeout 400 python -m pytest tests/test_gpu_session.py tests/test_gpu_tensor.py -x -q > $OUT/pytest_gate.log 2>&1
echo "gate rc=$?" | tee -a $OUT/summary.txt
tail -15 $OUT/pytest_gate.log
if [ "$2" != "skip-tests" ]; then
  timeout 1000 python -m pytest tests -m gpu -q --durations=12 > $OUT/pytest_gpu.log 2>&1
  echo "pytest rc=$?" | tee -a $OUT/summary.txt
  tail -30 $OUT/pytest_gpu.log
fi
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err
echo "bench rc=$?" | tee -a $OUT/summary.txt
cat $OUT/bench.json; tail -5 $OUT/bench.err
