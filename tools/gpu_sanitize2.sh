#!/bin/bash
# Second compute-sanitizer pass: racecheck (shared-memory hazards) and synccheck over tools/sanitize_target.py.
OUT=gpurun_out/${1:-san2}; mkdir -p $OUT
export PYTHONUNBUFFERED=1
CS=/usr/local/cuda/bin/compute-sanitizer
echo "== racecheck, inference"
timeout -s KILL 90 $CS --tool racecheck --print-limit 30 --error-exitcode 3 python tools/sanitize_target.py infer > $OUT/racecheck_infer.log 2>&1; echo "rc=$?" >> $OUT/racecheck_infer.log; tail -12 $OUT/racecheck_infer.log | cut -c1-200
echo "== synccheck, all"
timeout -s KILL 40 $CS --tool synccheck --print-limit 30 --error-exitcode 3 python tools/sanitize_target.py all > $OUT/synccheck_all.log 2>&1; echo "rc=$?" >> $OUT/synccheck_all.log; tail -6 $OUT/synccheck_all.log | cut -c1-200
echo "== racecheck, training"
timeout -s KILL 80 $CS --tool racecheck --print-limit 30 --error-exitcode 3 python tools/sanitize_target.py train > $OUT/racecheck_train.log 2>&1; echo "rc=$?" >> $OUT/racecheck_train.log; tail -12 $OUT/racecheck_train.log | cut -c1-200
