#!/bin/bash
OUT=gpurun_out/${1:-r3o}; mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_channel_kernels.py -q -m gpu 2>&1 | tail -25 | tee $OUT/tests.log
timeout 900 python -m pytest tests -q -m gpu -x -k "mixer or model or tiny_2048 or baseline or four_launch or conv_pool or gate" 2>&1 | tail -6 | tee -a $OUT/tests.log
for gk in 1 0; do FASTVIM_GROUP_KERNELS=$gk timeout 300 python bench.py --workload fastvim_t_2048 --no-cpu 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('group=$gk', d['value'], d['ms_per_step'], {k: v['avg_us'] for k, v in d['kernels'].items()})"; done
