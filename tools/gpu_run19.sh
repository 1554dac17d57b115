#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== attn correctness (auto = ws + mma.sync tails)"
timeout 300 python tools/gpu_kernel_check.py attn > gpurun_out/check_attn_auto.log 2>&1
grep -E "FAIL|GROUP|Error|watchdog|trap" gpurun_out/check_attn_auto.log | head -30
grep -E "us" gpurun_out/check_attn_auto.log | grep -E "fwd|dq" | cut -c1-130
echo "=== attn perf auto"
timeout 300 python tools/gpu_kernel_check.py attn_perf 2>&1 | grep -E "INFO|FAIL|Error" | tee gpurun_out/attn_perf_auto.log
echo "=== gemm sanity (new mbarrier wait)"
CHECK_GROUPS="gemm_basic gemm_big" bash tools/gpu_kernel_sweep.sh 2>&1 | grep -E "GROUP|FAIL|TFLOP" | cut -c1-140
echo "=== tests"
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
echo "=== bench B=8"
timeout 300 python bench.py --batch 8 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench19_b8.json
echo "=== bench B=2"
timeout 300 python bench.py --batch 2 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench19_b2.json
