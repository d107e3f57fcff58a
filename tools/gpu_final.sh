#!/bin/bash
# Round-end check on one B200: full GPU suite, smoke, default bench line, training lines.
OUT=gpurun_out/${1:-final}; mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== full gpu suite"; timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -6 | tee $OUT/gpu_tests.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke.log
echo "== bench default"; timeout 600 python bench.py 2>$OUT/bench.err | tail -1 > $OUT/bench_t224_n1.json; cut -c1-700 $OUT/bench_t224_n1.json
echo "== bench train B"; timeout 400 python bench.py --workload fastvim_b_224_train --steps 10 2>/dev/null | tail -1 > $OUT/bench_train_b_n1.json; cut -c1-300 $OUT/bench_train_b_n1.json
echo "== bench train T"; timeout 400 python bench.py --workload fastvim_t_224_train --steps 10 2>/dev/null | tail -1 > $OUT/bench_train_t_n1.json; cut -c1-300 $OUT/bench_train_t_n1.json
