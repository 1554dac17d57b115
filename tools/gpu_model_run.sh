#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for mode in auto 1cta 2cta; do
  echo "=== kernels DVLA_GEMM=$mode"
  if [ $mode = auto ]; then unset DVLA_GEMM; else export DVLA_GEMM=$mode; fi
  CHECK_GROUPS="gemm_basic gemm_epilogue gemm_big" bash tools/gpu_kernel_sweep.sh 2>&1 | grep -E "GROUP|FAIL|TFLOP|watchdog|rror" | cut -c1-150
done
unset DVLA_GEMM
echo "=== pytest model"; timeout 900 python -m pytest tests/test_model_gpu.py -x -q -m gpu 2>&1 | tail -3
echo "=== bench graph"; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -3 | tee gpurun_out/bench.log
echo "=== bench B8"; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --batch 8 2>&1 | tail -3 | tee gpurun_out/bench_b8.log
