#!/bin/bash
OUT=gpurun_out/${1:-last}; mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -4 | tee $OUT/gpu_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $OUT/smoke.log
timeout 300 python bench.py --no-cpu --steps 10 2>$OUT/bench.err | tail -1 | cut -c1-260 | tee $OUT/bench_short.json
