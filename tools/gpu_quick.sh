#!/bin/bash
# Quick iteration on the fused block kernel: parity tests, kernel timing, one ncu --set full capture.
TAG=${1:-quick}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== fused tests"; timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "block_fwd or fused" 2>&1 | tail -15 | tee $OUT/fused_tests.log
echo "== kbench"; timeout 300 python tools/kbench.py --shape t224 --only block 2>&1 | tee $OUT/kbench_t224.log
echo "== ncu full"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:block_fwd -c 1 -o $OUT/block_fwd env KBENCH_EAGER=1 python tools/kbench.py --shape t224 --only block --iters 2 > $OUT/ncu_full.log 2>&1
ls $OUT
