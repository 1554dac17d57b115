#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== attn correctness"
timeout 300 python tools/gpu_kernel_check.py attn > gpurun_out/check_attn_auto.log 2>&1
grep -E "FAIL|GROUP|Error|watchdog|trap" gpurun_out/check_attn_auto.log | head -30
echo "=== attn perf"
timeout 300 python tools/gpu_kernel_check.py attn_perf 2>&1 | grep -E "INFO|FAIL|Error" | tee gpurun_out/attn_perf_ws3.log
echo "=== torch op profile"
timeout 300 python tools/torch_op_profile.py --batch 8 2>&1 | tail -60 | cut -c1-230 | tee gpurun_out/torch_ops_b8.txt
echo "=== bench B=8"
timeout 300 python bench.py --batch 8 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench25_b8.json | cut -c1-300
