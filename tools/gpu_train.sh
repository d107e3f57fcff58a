#!/bin/bash
OUT=gpurun_out/${1:-train}; mkdir -p $OUT
export PYTHONUNBUFFERED=1
for wl in fastvim_b_224_train fastvim_t_224_train; do
  timeout 400 python bench.py --workload $wl --steps 10 2>$OUT/$wl.err | tail -1 > $OUT/bench_${wl}_n1.json
  python - <<PY
import json
d=json.load(open("$OUT/bench_${wl}_n1.json")); print("$wl", d["value"], "img/s", d["ms_per_step"], "ms  e2e", d["e2e"]["value"], d["config"]["launch"], "loss", d["loss"], "launches", d["gpu_launches_per_step"])
PY
  tail -2 $OUT/$wl.err
done
timeout 300 python bench.py --workload fastvim_t_224_train --steps 5 --no-graph 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('eager T', d['value'], d['ms_per_step'], 'loss', d['loss'])"
