#!/bin/bash
OUT=gpurun_out/${1:-r3c}; mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_gemm_general.py -q -m gpu 2>&1 | tail -30 | tee $OUT/tests.log
timeout 200 python tools/kbench.py --shape t224 --only gemm 2>&1 | grep -v tflops | tee $OUT/kbench_t224_gemm.jsonl
timeout 300 python bench.py --steps 20 --warmup 5 2>$OUT/bench.err | tail -1 > $OUT/bench_default_n1.json
python - <<PY
import json
d=json.load(open("$OUT/bench_default_n1.json")); print(d["value"], "img/s", d["ms_per_step"], "ms e2e", d["e2e"]["value"], "u8", d["e2e_u8"]["value"]); print(json.dumps(d["kernels"])[:1500]); print(d["roofline"])
for k,v in d["extra_workloads"].items(): print(k, v["value"], v["ms_per_step"])
PY
tail -3 $OUT/bench.err
