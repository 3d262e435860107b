#!/bin/bash
# Trainer: tests with the shipped build, then stage clocks of one Adam step
# (debug build on the box).
OUT=gpurun_out/${1:-fitprof}; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_bounds_api.py -m gpu -q -x > $OUT/pytest.log 2>&1
echo "pytest rc=$?"; tail -8 $OUT/pytest.log
timeout 300 python tools/bench_fit.py > $OUT/bench_fit.txt 2>&1; cat $OUT/bench_fit.txt
NB200_EXTRA_FLAGS=-DNB200_FIT_PROF python -c "from nautilus_b200 import _lib; _lib.build(force=True)" > $OUT/build.log 2>&1
echo "build rc=$?"
timeout 300 python tools/bench_fit.py 2>&1 | grep -v ffma | head -4 > $OUT/bench_fit_prof.txt; cat $OUT/bench_fit_prof.txt
