#!/bin/bash
# log Z at N_eff = 1e5: config test and the bench line.
OUT=gpurun_out/${1:-spec}; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_configs.py -m gpu -q -s --durations=5 > $OUT/pytest.log 2>&1
echo "pytest rc=$?"; grep -n "config 2" $OUT/pytest.log; tail -6 $OUT/pytest.log
timeout 600 python bench.py --no-cpu --no-later > $OUT/bench.json 2> $OUT/bench.err
python - <<PY
import json
d=json.loads(open('$OUT/bench.json').read().strip().splitlines()[-1])
print(d['value'], d['delta_log_z'], d['log_z_run'])
PY
