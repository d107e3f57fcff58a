#!/bin/bash
OUT=gpurun_out/${1:-scan}; mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== scan tests"; timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_vs_reference_cuda.py -q -k "scan or mixer or 2048 or sharded or conv_pool" 2>&1 | tail -8 | tee $OUT/scan_tests.log
echo "== kbench t2048"; timeout 300 python tools/kbench.py --shape t2048 --only scan,conv_pool,gate 2>&1 | tee $OUT/kbench_t2048.log
echo "== bench t2048"; timeout 600 python bench.py --workload fastvim_t_2048 --no-cpu --steps 10 2>&1 | tail -1 | cut -c1-2500 | tee $OUT/bench_t2048.json
