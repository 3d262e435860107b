#!/bin/bash
# End-to-end config-2 run with a host profile.
OUT=gpurun_out/${1:-cfg2}; mkdir -p $OUT
timeout 300 python tools/run_config.py --config 2 --n-eff 40000 --arith f16 --profile > $OUT/cfg2_prof.txt 2>&1
head -40 $OUT/cfg2_prof.txt | cut -c1-150
tail -1 $OUT/cfg2_prof.txt | cut -c1-200
