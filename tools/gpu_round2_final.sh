#!/bin/bash
# Final pass of round 2: smoke, full GPU tests, the default bench line, the
# config-5 line, launch list and full ncu capture of the two hot kernels.
TAG=${1:-r2_final}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1
echo "smoke rc=$?" | tee -a $OUT/summary.txt
timeout 300 python -m pytest tests -m gpu -q --durations=6 > $OUT/pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee -a $OUT/summary.txt
tail -9 $OUT/pytest_gpu.log
timeout 400 python bench.py > $OUT/bench.json 2> $OUT/bench.err
echo "bench rc=$?" | tee -a $OUT/summary.txt
timeout 200 python bench.py --config 5 --steps 5 --warmup 3 --no-cpu > $OUT/bench_cfg5.json 2> $OUT/bench_cfg5.err
echo "bench cfg5 rc=$?" | tee -a $OUT/summary.txt
B="python bench.py --steps 4 --warmup 3 --no-cpu --no-logz"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
  --log-file $OUT/launches.csv $B > $OUT/ncu_launch.log 2>&1
echo "launch list rc=$?" | tee -a $OUT/summary.txt
timeout 400 ncu --set full --clock-control none --import-source on \
  -k regex:'k_front_mma|k_mlp_tf32' -s 6 -c 2 -f -o $OUT/hot_kernels \
  $B --no-later > $OUT/ncu_full.log 2>&1
echo "ncu full rc=$?" | tee -a $OUT/summary.txt
python tools/ncu_summary.py $OUT/hot_kernels.ncu-rep > $OUT/ncu_front_and_mlp.txt 2>&1
python tools/ncu_hot_lines.py $OUT/hot_kernels.ncu-rep k_front_mma 25 > $OUT/ncu_hot_front.txt 2>&1
python tools/ncu_src_lines.py $OUT/hot_kernels.ncu-rep k_front_mma nb200_front_mma.cu 40 > $OUT/ncu_src_front.txt 2>&1
python tools/ncu_hot_lines.py $OUT/hot_kernels.ncu-rep k_mlp_tf32 25 > $OUT/ncu_hot_mlp.txt 2>&1
python - <<'PY'
import json
for f in ('bench.json', 'bench_cfg5.json'):
    try:
        d = json.loads(open('gpurun_out/r2_final/' + f).read().strip().splitlines()[-1])
        print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['stages_ms'], d.get('delta_log_z'), (d.get('log_z_run') or {}).get('wall_s'), (d.get('cpu_baseline') or {}).get('value'), (d.get('later_bounds') or {}).get('value'))
    except Exception as e:
        print(f, 'unreadable', e)
PY
head -30 $OUT/ncu_front_and_mlp.txt
