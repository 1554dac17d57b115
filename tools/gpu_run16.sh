#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== attn bwd ws correctness (fwd legacy)"
DVLA_ATTN_BWD=ws timeout 300 python tools/gpu_kernel_check.py attn > gpurun_out/check_attn_bwd_ws.log 2>&1
grep -E "FAIL|GROUP|Error|watchdog|trap" gpurun_out/check_attn_bwd_ws.log | head -30
grep -E "us" gpurun_out/check_attn_bwd_ws.log | grep -E "dq" | cut -c1-130
echo "=== attn ws+ws correctness"
DVLA_ATTN_FWD=ws DVLA_ATTN_BWD=ws timeout 300 python tools/gpu_kernel_check.py attn > gpurun_out/check_attn_ws_ws.log 2>&1
grep -E "FAIL|GROUP|Error|watchdog|trap" gpurun_out/check_attn_ws_ws.log | head -30
echo "=== attn perf ws/ws"
DVLA_ATTN_FWD=ws DVLA_ATTN_BWD=ws timeout 300 python tools/gpu_kernel_check.py attn_perf 2>&1 | grep -E "INFO|FAIL|Error" | tee gpurun_out/attn_perf_ws_ws.log
echo "=== ncu fwd ws (decoder shape)"
DVLA_ATTN_FWD=ws timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_ws -s 1 -c 1 -o gpurun_out/prof_attn_fwd_ws python tools/prof_attn.py dec > gpurun_out/ncu_attn_fwd_ws.log 2>&1
echo "=== ncu bwd ws (decoder shape)"
DVLA_ATTN_BWD=ws timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_d.*_ws -s 0 -c 2 -o gpurun_out/prof_attn_bwd_ws python tools/prof_attn.py dec bwd > gpurun_out/ncu_attn_bwd_ws.log 2>&1
ls -la gpurun_out/*.ncu-rep
echo "=== bench B=8 ws/ws"
DVLA_ATTN_FWD=ws DVLA_ATTN_BWD=ws timeout 300 python bench.py --batch 8 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench16_b8_ws.json
