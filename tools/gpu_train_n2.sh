#!/bin/bash
OUT=gpurun_out/${1:-train2}; mkdir -p $OUT
export PYTHONUNBUFFERED=1
for wl in fastvim_b_224_train; do
  timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload $wl --steps 6 2>$OUT/$wl.err | tail -1 > $OUT/bench_${wl}_n2.json
  python - <<PY
import json
d=json.load(open("$OUT/bench_${wl}_n2.json")); print("$wl N=2", d["value"], "img/s", d["ms_per_step"], "ms  e2e", d["e2e"]["value"], d["config"]["launch"], "loss", d["loss"])
PY
  grep -v "^$" $OUT/$wl.err | tail -3
done
