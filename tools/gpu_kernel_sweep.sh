#!/bin/bash
# Runs every kernel-check group in its own process (with a timeout) and collects logs under gpurun_out/.
# CHECK_GROUPS="gemm_basic attn ..." selects groups; the attention group is repeated for every kernel variant.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv | tee gpurun_out/gpu_info.txt
for g in ${CHECK_GROUPS:-gemm_basic gemm_epilogue gemm_splitk gemm_big norm attn attn_perf loss}; do
  echo "=== $g"
  timeout ${GROUP_TIMEOUT:-240} python tools/gpu_kernel_check.py $g --json gpurun_out/check_$g.json > gpurun_out/check_$g.log 2>&1
  echo "exit=$?" >> gpurun_out/check_$g.log
  grep -E "FAIL|GROUP|INFO|exit=|Error|error|watchdog" gpurun_out/check_$g.log | head -40
  grep -E "TFLOP|us" gpurun_out/check_$g.log | grep PASS | head -20
done
if [[ " ${CHECK_GROUPS:-attn} " == *" attn "* ]]; then
  for cfg in "DVLA_ATTN_FWD=legacy DVLA_ATTN_BWD=legacy" "DVLA_ATTN_FWD=tc DVLA_ATTN_BWD=tc" "DVLA_ATTN_FWD=legacy DVLA_ATTN_BWD=pipe"; do
    echo "=== attn ($cfg)"
    env $cfg timeout ${GROUP_TIMEOUT:-240} python tools/gpu_kernel_check.py attn > "gpurun_out/check_attn_${cfg// /_}.log" 2>&1
    grep -E "FAIL|GROUP|Error|watchdog" "gpurun_out/check_attn_${cfg// /_}.log" | head -20
    grep -E "us" "gpurun_out/check_attn_${cfg// /_}.log" | grep -E "fwd|dq" | cut -c1-120
  done
fi
