#!/bin/bash
OUT=gpurun_out/${1:-gemm}; mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== gemm tests"; timeout 240 python -m pytest tests/test_gpu_parity.py -q -x -k "gemm" 2>&1 | tail -25 | tee $OUT/gemm_tests.log
echo "== kbench gemm"; timeout 200 python tools/kbench.py --shape t224 --only gemm 2>&1 | tee $OUT/kbench_gemm.log
