#!/bin/bash
# MVEE kernel: timing against the host restatement + the bound tests.
OUT=gpurun_out/${1:-mvee_ab}; mkdir -p $OUT
python tools/bench_mvee.py > $OUT/mvee.json 2>&1; cat $OUT/mvee.json
timeout 600 python -m pytest tests/test_gpu_bounds_api.py tests/test_gpu_configs.py -m gpu -q -x > $OUT/pytest.log 2>&1
echo "pytest rc=$?"; tail -5 $OUT/pytest.log
timeout 300 python tools/run_config.py --config 2 --n-eff 10000 --arith f16 > $OUT/cfg2_run.txt 2>&1; tail -2 $OUT/cfg2_run.txt
