#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
echo "=== attn ws correctness"
DVLA_ATTN_FWD=ws timeout 300 python tools/gpu_kernel_check.py attn > gpurun_out/check_attn_ws.log 2>&1
grep -E "FAIL|GROUP|Error|watchdog|trap" gpurun_out/check_attn_ws.log | head -30
grep -E "us" gpurun_out/check_attn_ws.log | grep -E "fwd" | cut -c1-130
echo "=== attn perf legacy"
timeout 300 python tools/gpu_kernel_check.py attn_perf 2>&1 | grep -E "INFO|FAIL|Error" | tee gpurun_out/attn_perf_legacy.log
echo "=== attn perf ws"
DVLA_ATTN_FWD=ws timeout 300 python tools/gpu_kernel_check.py attn_perf 2>&1 | grep -E "INFO|FAIL|Error" | tee gpurun_out/attn_perf_ws.log
echo "=== splitk"
CHECK_GROUPS="gemm_splitk gemm_epilogue" bash tools/gpu_kernel_sweep.sh 2>&1 | tail -30 | tee gpurun_out/sweep15.log
echo "=== splitk off"
DVLA_GEMM_SPLITK=0 timeout 200 python tools/gpu_kernel_check.py gemm_splitk 2>&1 | grep -E "PASS|FAIL|GROUP" | cut -c1-150 | tee gpurun_out/splitk_off.log
echo "=== tests"
timeout 900 python -m pytest tests/test_train_step_gpu.py tests/test_model_gpu.py -x -q -m gpu 2>&1 | tail -3
echo "=== bench B=8"
timeout 300 python bench.py --batch 8 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench15_b8.json
echo "=== bench B=8 ws"
DVLA_ATTN_FWD=ws timeout 300 python bench.py --batch 8 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench15_b8_ws.json
echo "=== bench B=8 splitk off"
DVLA_GEMM_SPLITK=0 timeout 300 python bench.py --batch 8 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench15_b8_nosplit.json
echo "=== bench B=2"
timeout 300 python bench.py --batch 2 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench15_b2.json
echo "=== ref on gpu"
timeout 500 python tools/ref_gpu_bench.py --batch 2 8 2>&1 | tail -4 | tee gpurun_out/ref_gpu_math.log
timeout 500 python tools/ref_gpu_bench.py --batch 2 8 --sdpa 2>&1 | tail -4 | tee gpurun_out/ref_gpu_sdpa.log
