#!/bin/bash
# Round 2 close-out: smoke, full gpu tests, default bench, config 5, reference
# arm, end-to-end config-2 run with a host profile.
TAG=${1:-r2e}
OUT=gpurun_out/$TAG
bash tools/gpu_r2c.sh $TAG
timeout 300 python tools/run_config.py --config 2 --n-eff 10000 --arith f16 --profile > $OUT/cfg2_run.txt 2>&1
echo "cfg2 run rc=$?" | tee -a $OUT/summary.txt
tail -42 $OUT/cfg2_run.txt
