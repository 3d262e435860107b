#!/bin/bash
# Full ncu capture of the front kernel only (source-level), summaries as text.
TAG=${1:-r2_front}
OUT=gpurun_out/$TAG
mkdir -p $OUT
B="python bench.py --steps 4 --warmup 3 --no-cpu --no-logz --no-later"
timeout 400 ncu --set full --clock-control none --import-source on \
  -k regex:'k_front_mma' -s 6 -c 1 -f -o $OUT/front \
  $B > $OUT/ncu_full.log 2>&1
echo "ncu full rc=$?"
python tools/ncu_summary.py $OUT/front.ncu-rep > $OUT/ncu_front.txt 2>&1
python tools/ncu_hot_lines.py $OUT/front.ncu-rep k_front_mma 30 > $OUT/ncu_hot_front.txt 2>&1
for f in nb200_front_mma.cu nb200_rng.cuh nb200_device.cuh; do
  python tools/ncu_src_lines.py $OUT/front.ncu-rep k_front_mma $f 45 > $OUT/ncu_src_$f.txt 2>&1
done
head -40 $OUT/ncu_front.txt
head -24 $OUT/ncu_hot_front.txt
