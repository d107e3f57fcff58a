#!/bin/bash
# ncu --set full of the three backward kernels at FastVim-B shape (one launch each)
OUT=gpurun_out/${1:-ncubwd}; mkdir -p $OUT
export PYTHONUNBUFFERED=1 KBENCH_EAGER=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"scan_bwd_short|gate_bwd|conv_pool_bwd_stream" \
   --launch-skip 9 --launch-count 3 -o $OUT/bwd_b224 -f python tools/kbench.py --shape b224 --only bwd --iters 2 > $OUT/ncu.log 2>&1
tail -3 $OUT/ncu.log
ls -la $OUT
