#!/bin/bash
OUT=gpurun_out/${1:-r3f}; mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_gemm_general.py -q -m gpu -x 2>&1 | tail -12 | tee $OUT/tests.log
for fl in 1 0; do
FASTVIM_FLOW=$fl timeout 300 python bench.py --steps 20 --warmup 5 --no-extra --no-cpu 2>$OUT/bench_flow$fl.err | tail -1 > $OUT/bench_default_flow$fl.json
python - <<PY
import json
d=json.load(open("$OUT/bench_default_flow$fl.json")); print("flow=$fl", d["value"], "img/s", d["ms_per_step"], "ms e2e", d["e2e"]["value"], "u8", d["e2e_u8"]["value"]); print({k: v["avg_us"] for k, v in d["kernels"].items()}, d["kernels_total"]["sum_ms"], d["kernels_total"].get("ms_per_step_serialised")); print(d["roofline"])
PY
tail -2 $OUT/bench_flow$fl.err
done
