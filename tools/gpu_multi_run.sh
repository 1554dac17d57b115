#!/bin/bash
# 2-GPU validation: NCCL gradient all-reduce inside the captured step, bench.py under torchrun, clean tear-down
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export DVLA_BENCH_VERBOSE=1
t0=$(date +%s)
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --batch 8 --no-cpu-baseline > gpurun_out/bench_2gpu.log 2>&1
echo "exit=$? wall=$(( $(date +%s) - t0 ))s"
grep -E "^\[bench|^\{|Error|Traceback" gpurun_out/bench_2gpu.log | cut -c1-1500
