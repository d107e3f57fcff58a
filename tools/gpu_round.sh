#!/bin/bash
# One gpurun call: fused-kernel parity, kernel timings, full GPU suite, bench line, ncu launch list + full capture.
# Usage (from the repo root on the GPU box):  bash tools/gpu_round.sh [tag]
TAG=${1:-run}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
echo "== fused tests"; timeout 900 python -m pytest tests/test_gpu_parity.py -q -k "block_fwd or fused" 2>&1 | tail -40 | tee $OUT/fused_tests.log
if grep -q "failed\|error\|Error" $OUT/fused_tests.log || ! grep -q "passed" $OUT/fused_tests.log; then
  echo "FUSED TESTS NOT GREEN -> disabling the fused path for the remaining steps"; export FASTVIM_FUSED_BLOCK=0
fi
echo "== kbench"; timeout 600 python tools/kbench.py --shape t224 2>&1 | tee $OUT/kbench_t224.log
echo "== full gpu suite"; timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 | tee $OUT/gpu_tests.log
echo "== bench"; timeout 900 python bench.py 2>$OUT/bench.err | tee $OUT/bench.json
echo "== ncu launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 450 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-graph > $OUT/ncu_bench.log 2>&1
echo "== ncu full (fused kernel)"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:block_fwd -c 2 -o $OUT/block_fwd env KBENCH_EAGER=1 python tools/kbench.py --shape t224 --only block --iters 2 > $OUT/ncu_full.log 2>&1
ls -la $OUT
