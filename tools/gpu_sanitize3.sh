#!/bin/bash
# Third compute-sanitizer pass: initcheck (reads of device memory nobody wrote) over tools/sanitize_target.py.
OUT=gpurun_out/${1:-san3}; mkdir -p $OUT
export PYTHONUNBUFFERED=1
CS=/usr/local/cuda/bin/compute-sanitizer
echo "== initcheck, inference"
timeout -s KILL 70 $CS --tool initcheck --print-limit 40 --error-exitcode 3 python tools/sanitize_target.py infer > $OUT/initcheck_infer.log 2>&1; echo "rc=$?" >> $OUT/initcheck_infer.log; grep -c "Uninitialized" $OUT/initcheck_infer.log; grep "at .*fv::\|at fv::\|ERROR SUMMARY\|sanitize\]" $OUT/initcheck_infer.log | sort | uniq -c | cut -c1-220 | head -30
echo "== initcheck, training"
timeout -s KILL 60 $CS --tool initcheck --print-limit 40 --error-exitcode 3 python tools/sanitize_target.py train > $OUT/initcheck_train.log 2>&1; echo "rc=$?" >> $OUT/initcheck_train.log; grep -c "Uninitialized" $OUT/initcheck_train.log; grep "at .*fv::\|at fv::\|ERROR SUMMARY\|sanitize\]" $OUT/initcheck_train.log | sort | uniq -c | cut -c1-220 | head -30
