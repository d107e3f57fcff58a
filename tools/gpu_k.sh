#!/bin/bash
# usage: tools/gpu_k.sh <outdir-tag> <pytest -k expression> [extra shell command]
OUT=gpurun_out/${1:-k}; mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -q -m gpu -x -k "$2" 2>&1 | tail -25 | tee $OUT/tests.log
if [ -n "$3" ]; then bash -c "$3" 2>&1 | tail -20 | tee $OUT/extra.log; fi
