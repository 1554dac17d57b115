#!/bin/bash
# ncu evidence of the round: launch list of one eager step (B=8) + full captures of the dominant kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== ncu launch list (graph-less step, B=8)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 6600 -c 2300 --csv --log-file gpurun_out/launches_step_b8.csv \
   python bench.py --batch 8 --steps 1 --warmup 3 --no-graph --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
python tools/ncu_summarize.py gpurun_out/launches_step_b8.csv | tee gpurun_out/launches_step_b8_summary.txt | head -36
echo "=== ncu full: gemm"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 2 -c 4 -o gpurun_out/prof_gemm3 python tools/prof_gemm.py gemm > gpurun_out/ncu_gemm3.log 2>&1
echo "=== ncu full: attention (gpt2 shape, mask + dropout)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_.*_ws -s 2 -c 3 -o gpurun_out/prof_attn_gpt_ws python tools/prof_attn.py gpt bwd > gpurun_out/ncu_attn_gpt.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -4
