#!/bin/bash
OUT=gpurun_out/${1:-smoke}; mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke.log
echo "== tests"; timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -8 | tee $OUT/gpu_tests.log
echo "== bench t2048"; timeout 600 python bench.py --workload fastvim_t_2048 --no-cpu --steps 10 2>&1 | tail -1 | cut -c1-1500 | tee $OUT/bench_t2048.json
echo "== bench train T"; timeout 600 python bench.py --workload fastvim_t_224_train --steps 5 2>&1 | tail -1 | cut -c1-900 | tee $OUT/bench_train_t.json
