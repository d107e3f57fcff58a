#!/bin/bash
OUT=gpurun_out/${1:-bwdk}; mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests -q -m gpu -x -k "backward or grads or bwd" 2>&1 | tail -4 | tee $OUT/tests.log
for sh in b224 t224; do
  timeout 200 python tools/kbench.py --shape $sh --only bwd 2>&1 | tail -3 | tee -a $OUT/kbench_bwd.jsonl
done
