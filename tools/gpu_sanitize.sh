#!/bin/bash
# compute-sanitizer passes over tools/sanitize_target.py on one B200 (SURVEY.md "race detection / sanitizers").  Every step
# has its own hard timeout; logs -> gpurun_out/<tag>/ -> profiles/r02_sanitizer_*.log.
OUT=gpurun_out/${1:-san}; mkdir -p $OUT
export PYTHONUNBUFFERED=1
CS=/usr/local/cuda/bin/compute-sanitizer
echo "== plain run (warms the page cache, proves the target itself passes)"
timeout -s KILL 70 python tools/sanitize_target.py all 2>&1 | tail -8 | tee $OUT/plain.log
echo "== memcheck, inference"
timeout -s KILL 85 $CS --tool memcheck --print-limit 20 --error-exitcode 3 python tools/sanitize_target.py infer > $OUT/memcheck_infer.log 2>&1; echo "rc=$?" >> $OUT/memcheck_infer.log; tail -6 $OUT/memcheck_infer.log
echo "== memcheck, training"
timeout -s KILL 70 $CS --tool memcheck --print-limit 20 --error-exitcode 3 python tools/sanitize_target.py train > $OUT/memcheck_train.log 2>&1; echo "rc=$?" >> $OUT/memcheck_train.log; tail -5 $OUT/memcheck_train.log
