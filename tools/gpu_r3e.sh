#!/bin/bash
OUT=gpurun_out/${1:-r3e}; mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_gemm_general.py tests/test_gpu_baseline_configs.py -q -m gpu -x 2>&1 | tail -8 | tee $OUT/tests.log
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "gemm or block_fwd or mixer or model" 2>&1 | tail -8 | tee -a $OUT/tests.log
for pdl in 1 0; do
FASTVIM_PDL=$pdl timeout 300 python bench.py --steps 20 --warmup 5 --no-extra --no-cpu 2>$OUT/bench_pdl$pdl.err | tail -1 > $OUT/bench_default_pdl$pdl.json
python - <<PY
import json
d=json.load(open("$OUT/bench_default_pdl$pdl.json")); print("pdl=$pdl", d["value"], "img/s", d["ms_per_step"], "ms e2e", d["e2e"]["value"], "u8", d["e2e_u8"]["value"]); print({k: v["avg_us"] for k, v in d["kernels"].items()}, d["kernels_total"])
PY
tail -2 $OUT/bench_pdl$pdl.err
done
