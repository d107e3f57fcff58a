#!/bin/bash
# Last check of the round on one B200 (ordered by priority, the box time left is short): full GPU suite, default bench
# line, FastVim-B training line, smoke, FastVim-T training line, then the training line with the eager patch embedding.
OUT=gpurun_out/${1:-final2}; mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== full gpu suite"; timeout 300 python -m pytest tests -q -m gpu 2>&1 | tail -8 | tee $OUT/gpu_tests.log
echo "== bench default"; timeout 200 python bench.py 2>$OUT/bench.err | tail -1 > $OUT/bench_t224_n1.json; cut -c1-600 $OUT/bench_t224_n1.json
echo "== bench train B"; timeout 120 python bench.py --workload fastvim_b_224_train --steps 10 2>/dev/null | tail -1 > $OUT/bench_train_b_n1.json; cut -c1-300 $OUT/bench_train_b_n1.json
echo "== smoke"; timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke.log
echo "== bench train T"; timeout 100 python bench.py --workload fastvim_t_224_train --steps 10 2>/dev/null | tail -1 > $OUT/bench_train_t_n1.json; cut -c1-300 $OUT/bench_train_t_n1.json
echo "== bench train B, eager patch embed"; FASTVIM_NATIVE_PATCH_TRAIN=0 timeout 120 python bench.py --workload fastvim_b_224_train --steps 10 2>/dev/null | tail -1 > $OUT/bench_train_b_n1_eager_patch.json; cut -c1-300 $OUT/bench_train_b_n1_eager_patch.json
