#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== attn kernels"; CHECK_GROUPS="attn" bash tools/gpu_kernel_sweep.sh 2>&1 | grep -E "===|GROUP|FAIL|watchdog|rror|fwd [0-9]|bwd [0-9]" | cut -c1-150
echo "=== pytest all gpu"; timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
echo "=== bench graph"; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -3 | tee gpurun_out/bench.log
echo "=== bench B8"; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --batch 8 2>&1 | tail -3 | tee gpurun_out/bench_b8.log
