#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== attn kernels"; CHECK_GROUPS="attn" bash tools/gpu_kernel_sweep.sh 2>&1 | grep -E "===|GROUP|FAIL|watchdog|rror|fwd [0-9]|bwd [0-9]" | cut -c1-150
echo "=== pytest model"; timeout 900 python -m pytest tests/test_model_gpu.py -x -q -m gpu 2>&1 | tail -3
echo "=== bench graph"; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -3 | tee gpurun_out/bench.log
echo "=== bench B8"; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --batch 8 2>&1 | tail -3 | tee gpurun_out/bench_b8.log
