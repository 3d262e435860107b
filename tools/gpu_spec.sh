#!/bin/bash
# Speculative fit: sampler / config / bound tests, the bench line and the config-2 run.
OUT=gpurun_out/${1:-spec}; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_bounds_api.py tests/test_gpu_configs.py tests/test_gpu_sampler.py -m gpu -q --durations=5 > $OUT/pytest.log 2>&1
echo "pytest rc=$?"; tail -12 $OUT/pytest.log
timeout 600 python bench.py --no-cpu --no-later > $OUT/bench.json 2> $OUT/bench.err
python - <<PY
import json
d=json.loads(open('$OUT/bench.json').read().strip().splitlines()[-1])
print(d['value'], d['delta_log_z'], d['log_z_run'])
PY
for i in 1 2; do
  timeout 300 python tools/run_config.py --config 2 --n-eff 10000 --arith f16 > $OUT/cfg2_run$i.txt 2>&1
  tail -1 $OUT/cfg2_run$i.txt | cut -c1-200
done
