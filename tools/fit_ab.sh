#!/bin/bash
# A/B of the trainer's cluster size on the GPU box (rebuilds the library).
OUT=gpurun_out/${1:-fit_ab}; mkdir -p $OUT
for c in 4 8; do
  NB200_EXTRA_FLAGS=-DNB200_FIT_CLUSTER=$c python -c "from nautilus_b200 import _lib; _lib.build(force=True)"
  python tools/bench_fit.py > $OUT/fit_cluster$c.json 2>&1; cat $OUT/fit_cluster$c.json
done
