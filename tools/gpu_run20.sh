#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for cfg in "X=1" "DVLA_GEMM_SPLITK=0" "DVLA_ATTN_FWD=legacy DVLA_ATTN_BWD=tc" "DVLA_GEMM_SPLITK=0 DVLA_ATTN_FWD=legacy DVLA_ATTN_BWD=tc"; do
  for i in 1 2; do
    echo "=== $cfg run $i"
    env $cfg timeout 300 python -m pytest tests/test_train_step_gpu.py -x -q -m gpu -k graph_replay 2>&1 | grep -E "passed|failed|AssertionError|assert " | head -4 | cut -c1-400
  done
done
echo "=== attn perf ws (no hybrid, try_wait hint)"
timeout 300 python tools/gpu_kernel_check.py attn_perf 2>&1 | grep -E "INFO|FAIL|Error" | tee gpurun_out/attn_perf_ws2.log
