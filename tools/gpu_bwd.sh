#!/bin/bash
OUT=gpurun_out/${1:-bwd}; mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== backward tests"; timeout 900 python -m pytest tests -q -m gpu -k "backward or bwd or grad or train" 2>&1 | tail -8 | tee $OUT/bwd_tests.log
echo "== bench train B"; timeout 400 python bench.py --workload fastvim_b_224_train --steps 5 2>&1 | tail -1 | cut -c1-400 | tee $OUT/bench_train_b.json
echo "== bench train T"; timeout 400 python bench.py --workload fastvim_t_224_train --steps 5 2>&1 | tail -1 | cut -c1-400 | tee $OUT/bench_train_t.json
