#!/bin/bash
# End-to-end config-2 runs: two plain (run-to-run noise), one with a host profile.
OUT=gpurun_out/${1:-cfg2}; mkdir -p $OUT
for i in 1 2; do
  timeout 300 python tools/run_config.py --config 2 --n-eff 10000 --arith f16 > $OUT/cfg2_run$i.txt 2>&1
  tail -1 $OUT/cfg2_run$i.txt | cut -c1-120
done
timeout 300 python tools/run_config.py --config 2 --n-eff 10000 --arith f16 --profile > $OUT/cfg2_prof.txt 2>&1
head -36 $OUT/cfg2_prof.txt | cut -c1-150
grep -n "was called by" -A16 $OUT/cfg2_prof.txt | cut -c1-200 | head -40
tail -1 $OUT/cfg2_prof.txt | cut -c1-120
nproc; grep -m1 "model name" /proc/cpuinfo
