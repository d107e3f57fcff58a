#!/bin/bash
OUT=gpurun_out/${1:-tt}; mkdir -p $OUT
export PYTHONUNBUFFERED=1
for ml in 8 4 2 1; do
  for sh in b224 t224; do
    FASTVIM_BWD_MAXLEN=$ml timeout 100 python tools/kbench.py --shape $sh --only bwd 2>&1 | grep -v conv_pool | cut -c1-110 | sed "s/^/maxlen=$ml /" | tee -a $OUT/kbench_tt.log
  done
done
