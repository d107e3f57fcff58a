#!/bin/bash
# ncu --set full of the tcgen05 GEMM kernels (forward, dgrad, wgrad at FastVim-B; out_proj + add + RMSNorm epilogue at FastVim-T)
OUT=gpurun_out/${1:-ncu_gemm}; mkdir -p $OUT
export PYTHONUNBUFFERED=1 KBENCH_EAGER=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc|gemm_out_norm" --launch-skip 3 --launch-count 40 -o $OUT/gemm_b224 python tools/kbench.py --shape b224 --only gemm --iters 1 > $OUT/ncu_b224.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm_out_norm" --launch-skip 3 --launch-count 2 -o $OUT/out_norm_t224 python tools/kbench.py --shape t224 --only gemm --iters 1 > $OUT/ncu_t224.log 2>&1
tail -3 $OUT/ncu_b224.log $OUT/ncu_t224.log; ls -la $OUT
