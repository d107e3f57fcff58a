#!/bin/bash
OUT=gpurun_out/${1:-stagger}; mkdir -p $OUT
for ns in 1 2 4 8; do
  echo "split $ns"; FV_BLOCK_STAGGER_NS=$ns timeout 200 python tools/kbench.py --shape t224 --only block 2>&1 | tee -a $OUT/split.log
done
