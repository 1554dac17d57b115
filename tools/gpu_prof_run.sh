#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== attn kernels (tcgen05 fwd)"; CHECK_GROUPS="attn" bash tools/gpu_kernel_sweep.sh 2>&1 | tail -30
echo "=== ncu launch list (1 eager step)"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -s 6500 -c 2400 --csv --log-file gpurun_out/launches_step.csv \
   python bench.py --steps 1 --warmup 3 --no-graph --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
python tools/ncu_summarize.py gpurun_out/launches_step.csv | tee gpurun_out/launches_step_summary.txt | head -45
echo "=== ncu full: gemm"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 2 -c 4 -o gpurun_out/prof_gemm python tools/prof_gemm.py gemm > gpurun_out/ncu_gemm.log 2>&1
ls -la gpurun_out/*.ncu-rep
