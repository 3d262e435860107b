#!/bin/bash
# Scaling runs on ONE multi-GPU box: usage  bash tools/gpu_scale.sh TAG "1 2 4 8" [steps2] [steps5]
TAG=${1:-scale}
NS=${2:-"1 2"}
S2=${3:-400}
S5=${4:-10}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
run() {  # n extra-args out
  local n=$1; shift
  local out=$1; shift
  if [ "$n" = "1" ]; then
    timeout 900 python bench.py --gpus 1 "$@" > $out 2> $out.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n \
      --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n "$@" > $out 2> $out.err
  fi
  echo "n=$n $* rc=$?" | tee -a $OUT/summary.txt
  python - "$out" <<'PY'
import json, sys
try:
    line = [l for l in open(sys.argv[1]) if l.startswith('{')][-1]
    d = json.loads(line)
    print('  value %.4g  ms/step %.4g  e2e %.4g (%.3f of device)  scaling %s' % (
        d['value'], d['ms_per_step'], d['e2e']['value'],
        d['e2e'].get('frac_of_device', 0), d['scaling']))
except Exception as e:
    print('  (no line)', e)
PY
}
for n in $NS; do
  run $n $OUT/cfg2_n$n.json --steps $S2 --warmup 3 --no-cpu --no-logz --no-later
done
for n in $NS; do
  run $n $OUT/cfg5_n$n.json --config 5 --steps $S5 --warmup 3 --no-cpu
done
for f in $OUT/*.err; do tail -n 2 $f; done | tail -n 30
