#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv
echo "=== pytest train step"; timeout 900 python -m pytest tests/test_train_step_gpu.py -x -q -m gpu 2>&1 | tail -5
echo "=== bench 2 GPUs (B=8)"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --batch 8 2>&1 | tail -4 | tee gpurun_out/bench_2gpu.log
echo "=== bench 1 GPU (B=8) with gemm dump"; DVLA_BENCH_DUMP=1 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --batch 8 2> gpurun_out/gemm_dump_b8.txt | tail -2 | tee gpurun_out/bench_b8.log
sort -t= -k12 -n gpurun_out/gemm_dump_b8.txt | grep "^\[gemm\]" | sort -k 16 -n -r | head -30
