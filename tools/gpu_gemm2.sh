#!/bin/bash
OUT=gpurun_out/${1:-gemm2}; mkdir -p $OUT
export PYTHONUNBUFFERED=1
for sh in b224 s224; do timeout 200 python tools/kbench.py --shape $sh --only gemm 2>&1 | tee -a $OUT/kbench_gemm_$sh.log; done
echo "== ncu gemm"; timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -c 3 -o $OUT/gemm_tc env KBENCH_EAGER=1 python tools/kbench.py --shape t224 --only gemm --iters 1 > $OUT/ncu_gemm.log 2>&1
ls $OUT
