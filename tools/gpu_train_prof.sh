#!/bin/bash
OUT=gpurun_out/${1:-trainprof}; mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 3000 -c 3000 --csv --log-file $OUT/launches_train_b.csv python bench.py --workload fastvim_b_224_train --steps 2 --warmup 3 > $OUT/ncu_train.log 2>&1
tail -2 $OUT/ncu_train.log | cut -c1-400
timeout 300 python bench.py --workload fastvim_b_224_train --steps 5 2>&1 | tail -1 | cut -c1-600 | tee $OUT/bench_train_b.json
