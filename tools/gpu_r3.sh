#!/bin/bash
# round-2 late session: general tcgen05 GEMM -- tests, backward-path tests, training bench with / without the small GEMMs
OUT=gpurun_out/${1:-r3}; mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_gemm_general.py tests/test_gpu_baseline_configs.py -q -m gpu -x 2>&1 | tail -15 | tee $OUT/tests.log
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "train or grad or mixer or model" 2>&1 | tail -8 | tee -a $OUT/tests.log
for sm in 1 0; do
  FASTVIM_TC_SMALL_GEMM=$sm timeout 400 python bench.py --workload fastvim_b_224_train --steps 10 2>$OUT/train_b_small$sm.err | tail -1 > $OUT/bench_train_b_small$sm.json
  python - <<PY
import json
d=json.load(open("$OUT/bench_train_b_small$sm.json")); print("small=$sm", d["value"], "img/s", d["ms_per_step"], "ms loss", d["loss"], "launches", d["gpu_launches_per_step"]); print(json.dumps(d.get("kernels_total"))[:1500])
PY
done
FASTVIM_TC_GEMM=0 timeout 400 python bench.py --workload fastvim_b_224_train --steps 10 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('cublas-all', d['value'], d['ms_per_step'], 'loss', d['loss'])"
