#!/bin/bash
# One gpurun call: tests, bench (both arms), ncu launch list, full ncu capture
# of the two hot kernels, emulator timeline.  Outputs land in gpurun_out/.
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_round.sh TAG [skip-tests]'
TAG=${1:-run}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $OUT/gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $OUT/smoke.log 2>&1
echo "smoke rc=$?" | tee -a $OUT/summary.txt
# gate: the tensor-core and fused-cycle tests first, under a short timeout (a
# deadlocked kernel traps after 2^22 polls; do not spend the budget on it)
timeout 300 python -m pytest tests/test_gpu_tensor.py tests/test_gpu_session.py -x -q > $OUT/pytest_gate.log 2>&1
GATE=$?
echo "gate rc=$GATE" | tee -a $OUT/summary.txt
tail -5 $OUT/pytest_gate.log
if [ $GATE -ne 0 ]; then tail -40 $OUT/pytest_gate.log; exit 1; fi
if [ "$2" != "skip-tests" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q --durations=12 > $OUT/pytest_gpu.log 2>&1
  echo "pytest rc=$?" | tee -a $OUT/summary.txt
  tail -25 $OUT/pytest_gpu.log
fi
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err
echo "bench rc=$?" | tee -a $OUT/summary.txt
cat $OUT/bench.json
timeout 400 python bench.py --impl reference --steps 20 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
echo "bench reference rc=$?" | tee -a $OUT/summary.txt
cat $OUT/bench_reference.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv \
  --log-file $OUT/launches.csv python bench.py --steps 4 --warmup 3 --no-cpu > $OUT/ncu_launch.log 2>&1
echo "ncu launches rc=$?" | tee -a $OUT/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on \
  -k regex:'k_front|k_mlp_tf32' -s 4 -c 2 -f -o $OUT/hot_kernels \
  python bench.py --steps 4 --warmup 3 --no-cpu > $OUT/ncu_full.log 2>&1
echo "ncu full rc=$?" | tee -a $OUT/summary.txt
NB200_EXTRA_FLAGS=-DNB200_TIMELINE timeout 300 python tools/mlp_timeline.py > $OUT/mlp_timeline.txt 2>&1
echo "timeline rc=$?" | tee -a $OUT/summary.txt
tail -30 $OUT/mlp_timeline.txt
