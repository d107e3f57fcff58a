#!/bin/bash
OUT=gpurun_out/${1:-tests}; mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -q -m gpu ${2:+-k "$2"} 2>&1 | tail -30 | tee $OUT/gpu_tests.log
