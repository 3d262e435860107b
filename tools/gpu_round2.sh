#!/bin/bash
# Shorter GPU round: gate tests, full GPU tests, bench, optional extras.
# usage: bash tools/gpu_round2.sh TAG [extras...]   extras: dmma ncu e2e2 e2e1
TAG=${1:-run}; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $OUT/smoke.log 2>&1
echo "smoke rc=$?" | tee -a $OUT/summary.txt
timeout 300 python -m pytest tests/test_gpu_tensor.py tests/test_gpu_session.py -x -q > $OUT/pytest_gate.log 2>&1
GATE=$?
echo "gate rc=$GATE" | tee -a $OUT/summary.txt
if [ $GATE -ne 0 ]; then tail -40 $OUT/pytest_gate.log; exit 1; fi
timeout 900 python -m pytest tests -m gpu -x -q --durations=8 > $OUT/pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee -a $OUT/summary.txt
tail -15 $OUT/pytest_gpu.log
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err
echo "bench rc=$?" | tee -a $OUT/summary.txt
cat $OUT/bench.json
for x in "$@"; do
  case $x in
    dmma) ./tools/peak_dmma > $OUT/peak_dmma.json 2>&1; cat $OUT/peak_dmma.json;;
    ncu) timeout 600 ncu --set full --clock-control none --import-source on \
           -k regex:'k_front|k_mlp_tf32' -s 4 -c 2 -f -o $OUT/hot_kernels \
           python bench.py --steps 4 --warmup 3 --no-cpu > $OUT/ncu_full.log 2>&1
         echo "ncu rc=$?" | tee -a $OUT/summary.txt;;
    launches) timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv \
           --log-file $OUT/launches.csv python bench.py --steps 4 --warmup 3 --no-cpu > $OUT/ncu_launch.log 2>&1;;
    e2e2) timeout 600 python tools/run_config.py --config 2 --n-batch 1000 --profile > $OUT/e2e_cfg2.log 2>&1
          echo "e2e2 rc=$?" | tee -a $OUT/summary.txt; tail -45 $OUT/e2e_cfg2.log;;
    e2e1) timeout 300 python tools/run_config.py --config 1 > $OUT/e2e_cfg1.log 2>&1; tail -2 $OUT/e2e_cfg1.log;;
    ncufit) timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_mlp_fit -c 1 -f -o $OUT/fit_kernel python tools/bench_fit.py > $OUT/ncu_fit.log 2>&1; echo "ncufit rc=$?" | tee -a $OUT/summary.txt;;
    fitab) bash tools/fit_ab.sh $TAG;;
    mvee) python tools/bench_mvee.py > $OUT/mvee.json 2>&1; cat $OUT/mvee.json
          timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_mvee -c 1 -f -o $OUT/mvee_kernel python tools/bench_mvee.py > $OUT/ncu_mvee.log 2>&1;;
    fit) python tools/bench_fit.py --sklearn > $OUT/fit.json 2>&1; cat $OUT/fit.json;;
    e2e4) timeout 900 python tools/run_config.py --config 4 --n-batch 1000 > $OUT/e2e_cfg4.log 2>&1; tail -2 $OUT/e2e_cfg4.log;;
    timeline) NB200_EXTRA_FLAGS=-DNB200_TIMELINE timeout 300 python tools/mlp_timeline.py > $OUT/mlp_timeline.txt 2>&1;;
  esac
done
