#!/bin/bash
# Round 2, trainer iteration: UMMA peak, trainer tests + timing, config-2 run.
TAG=${1:-r2d}
OUT=gpurun_out/$TAG
mkdir -p $OUT
./tools/peak_umma > $OUT/peak_umma.json 2>&1; echo "peak rc=$?"; cat $OUT/peak_umma.json
timeout 900 python -m pytest tests/test_gpu_bounds_api.py tests/test_gpu_configs.py -m gpu -q -x --durations=5 > $OUT/pytest.log 2>&1
echo "pytest rc=$?"; tail -15 $OUT/pytest.log
timeout 300 python tools/bench_fit.py > $OUT/bench_fit.txt 2>&1; cat $OUT/bench_fit.txt
timeout 300 python tools/run_config.py --config 2 --n-eff 10000 --arith f16 --profile > $OUT/cfg2_run.txt 2>&1; tail -45 $OUT/cfg2_run.txt
