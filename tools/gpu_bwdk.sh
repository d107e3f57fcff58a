#!/bin/bash
OUT=gpurun_out/${1:-bwdk}; mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests -q -m gpu -x -k "backward or grads or bwd or training" 2>&1 | tail -12 | tee $OUT/tests.log
for sh in b224 t224; do
  timeout 100 python tools/kbench.py --shape $sh --only bwd 2>&1 | tail -3 | cut -c1-130 | tee -a $OUT/kbench_bwd.jsonl
done
