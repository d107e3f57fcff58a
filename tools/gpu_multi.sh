#!/bin/bash
OUT=gpurun_out/${1:-multi}; mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== multi-gpu tests"; timeout 600 python -m pytest tests/test_gpu_multi.py -q 2>&1 | tail -4 | tee $OUT/multi_tests.log
echo "== bench n2"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 2>$OUT/bench_n2.err | tail -1 | cut -c1-1400 | tee $OUT/bench_n2.json
echo "== bench t2048 n2 gather"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --workload fastvim_t_2048 --steps 10 --no-cpu 2>$OUT/bench_t2048_n2.err | tail -1 | cut -c1-900 | tee $OUT/bench_t2048_n2_gather.json
echo "== bench t2048 n2 reduce"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 2 --workload fastvim_t_2048 --steps 10 --no-cpu --out-mode reduce 2>$OUT/bench_t2048_n2r.err | tail -1 | cut -c1-900 | tee $OUT/bench_t2048_n2_reduce.json
echo "== ref arm"; timeout 400 python bench.py --impl reference --steps 2 --warmup 1 --cpu-budget 30 2>&1 | tail -1 | cut -c1-600 | tee $OUT/bench_ref.json
tail -3 $OUT/*.err
