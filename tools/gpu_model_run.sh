#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== pytest model"; timeout 900 python -m pytest tests/test_model_gpu.py -x -q -m gpu 2>&1 | tail -30 | tee gpurun_out/pytest_model.log
echo "=== smoke"; timeout 600 python __graft_entry__.py --smoke 2>&1 | tail -8 | tee gpurun_out/smoke.log
echo "=== bench eager"; timeout 900 python bench.py --steps 5 --warmup 3 --no-graph --no-cpu-baseline 2>&1 | tail -12 | tee gpurun_out/bench_eager.log
echo "=== bench graph"; timeout 900 python bench.py --steps 10 --warmup 3 2>&1 | tail -12 | tee gpurun_out/bench.log
