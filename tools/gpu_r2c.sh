#!/bin/bash
# Round 2: full gpu tests, bench config 2 (default) and config 5 at N=1.
TAG=${1:-r2c}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $OUT/smoke.log 2>&1
echo "smoke rc=$?" | tee -a $OUT/summary.txt
timeout 1200 python -m pytest tests -m gpu -q --durations=8 > $OUT/pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee -a $OUT/summary.txt
tail -25 $OUT/pytest_gpu.log
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err
echo "bench rc=$?" | tee -a $OUT/summary.txt
cat $OUT/bench.json; tail -5 $OUT/bench.err
timeout 600 python bench.py --config 5 --no-cpu > $OUT/bench_cfg5_n1.json 2> $OUT/bench_cfg5.err
echo "bench cfg5 rc=$?" | tee -a $OUT/summary.txt
cat $OUT/bench_cfg5_n1.json; tail -5 $OUT/bench_cfg5.err
timeout 400 python bench.py --impl reference --steps 4 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
echo "bench reference rc=$?" | tee -a $OUT/summary.txt
cat $OUT/bench_reference.json
