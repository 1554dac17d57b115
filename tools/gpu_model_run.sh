#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== kernels"; CHECK_GROUPS="gemm_basic gemm_epilogue gemm_big" bash tools/gpu_kernel_sweep.sh 2>&1 | tail -25
echo "=== pytest model"; timeout 900 python -m pytest tests/test_model_gpu.py -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/pytest_model.log
echo "=== smoke"; timeout 600 python __graft_entry__.py --smoke 2>&1 | grep smoke | tee gpurun_out/smoke.log
echo "=== bench graph"; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -4 | tee gpurun_out/bench.log
echo "=== bench B8"; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --batch 8 2>&1 | tail -3 | tee gpurun_out/bench_b8.log
