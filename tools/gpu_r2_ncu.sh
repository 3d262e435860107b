#!/bin/bash
# Round-2 profiling pass: launch list of the default bench command, full ncu
# captures of the hot kernels (front, emulator, exclusion prep + grouped
# emulator), summaries as text.
TAG=${1:-r2_ncu}
OUT=gpurun_out/$TAG
mkdir -p $OUT
B="python bench.py --steps 4 --warmup 3 --no-cpu --no-logz"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
  --log-file $OUT/launches.csv $B > $OUT/ncu_launch.log 2>&1
echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on \
  -k regex:'k_front_mma|k_mlp_tf32' -s 6 -c 2 -f -o $OUT/hot_kernels \
  $B --no-later > $OUT/ncu_full.log 2>&1
echo "ncu full rc=$?"
python tools/ncu_summary.py $OUT/hot_kernels.ncu-rep > $OUT/ncu_front_and_mlp.txt 2>&1
python tools/ncu_hot_lines.py $OUT/hot_kernels.ncu-rep k_front_mma 25 > $OUT/ncu_hot_front.txt 2>&1
python tools/ncu_hot_lines.py $OUT/hot_kernels.ncu-rep k_mlp_tf32 25 > $OUT/ncu_hot_mlp.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on \
  -k regex:'k_excl_prep|k_mlp_tf32<1' -c 2 -f -o $OUT/exclusion \
  python tools/profile_later.py 47 f16 > $OUT/ncu_excl.log 2>&1
echo "ncu exclusion rc=$?"
python tools/ncu_summary.py $OUT/exclusion.ncu-rep > $OUT/ncu_exclusion.txt 2>&1
cat $OUT/ncu_front_and_mlp.txt | head -60
head -30 $OUT/ncu_hot_front.txt
head -12 $OUT/ncu_exclusion.txt
