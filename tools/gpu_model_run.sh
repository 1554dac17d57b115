#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== attn kernels"; CHECK_GROUPS="attn" bash tools/gpu_kernel_sweep.sh 2>&1 | grep -E "===|GROUP|FAIL|watchdog|rror|fwd [0-9]|bwd [0-9]" | cut -c1-150
echo "=== pytest train step + model"; timeout 900 python -m pytest tests/test_train_step_gpu.py tests/test_model_gpu.py -x -q -m gpu 2>&1 | tail -5
echo "=== bench B8 pipe"; DVLA_ATTN_BWD=pipe timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --batch 8 2>&1 | tail -2 | cut -c1-200
echo "=== bench B2 pipe"; DVLA_ATTN_BWD=pipe timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -2 | cut -c1-200
echo "=== bench 1 GPU (B=8) with gemm dump"; DVLA_BENCH_DUMP=1 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --batch 8 2> gpurun_out/gemm_dump_b8.txt | tail -2 | cut -c1-200 | tee gpurun_out/bench_b8.log
grep "^\[gemm\]" gpurun_out/gemm_dump_b8.txt | sort -k 18 -n -r | head -28
