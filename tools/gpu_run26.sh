#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== norm/elementwise correctness"
CHECK_GROUPS="norm gemm_epilogue" bash tools/gpu_kernel_sweep.sh 2>&1 | grep -E "GROUP|FAIL" | cut -c1-160
echo "=== torch op profile"
timeout 300 python tools/torch_op_profile.py --batch 8 2>&1 | grep -E " ms n=|layernorm_bwd|colsum|act_bwd|elementwise|Self CUDA time" | cut -c1-230 | tee gpurun_out/torch_ops_b8_v2.txt
echo "=== tests"
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
echo "=== bench B=8"
timeout 300 python bench.py --batch 8 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench26_b8.json | cut -c1-300
