#!/bin/bash
# round-2 ncu evidence: launch list of ONE eager train step (B=8, exactly the timed region) + full captures of the dominant kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== launch list"
DVLA_BENCH_CUPROF=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_step_b8.csv \
   python bench.py --batch 8 --steps 1 --warmup 3 --no-graph --no-e2e --no-cpu-baseline --no-extras > gpurun_out/r2_ncu_bench.log 2>&1
python tools/ncu_summarize.py gpurun_out/r2_launches_step_b8.csv | tee gpurun_out/r2_launches_step_b8_summary.txt | head -45
echo "=== ncu full: gemm"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 2 -c 7 -o gpurun_out/r2_prof_gemm python tools/prof_gemm.py gemm > gpurun_out/r2_ncu_gemm.log 2>&1
echo "=== ncu full: attention (gpt2 shape, mask + dropout)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_.*_ws -s 2 -c 3 -o gpurun_out/r2_prof_attn_gpt python tools/prof_attn.py gpt bwd > gpurun_out/r2_ncu_attn_gpt.log 2>&1
echo "=== ncu full: HBM-bound kernels (LayerNorm fwd/bwd, act_bwd, colsum, AdamW) from one eager step"
DVLA_BENCH_CUPROF=1 timeout 900 ncu --profile-from-start off --set full --clock-control none -k regex:"layernorm_bwd|layernorm_fwd|act_bwd|colsum|adamw|sumsq|cat_broadcast" -c 12 -o gpurun_out/r2_prof_hbm \
   python bench.py --batch 8 --steps 1 --warmup 3 --no-graph --no-e2e --no-cpu-baseline --no-extras > gpurun_out/r2_ncu_hbm.log 2>&1
echo "=== ncu full: sampler megakernel"
timeout 600 ncu --set full --clock-control none -k regex:dit_ddim -c 1 -o gpurun_out/r2_prof_sampler python tools/debug_rollout.py > gpurun_out/r2_ncu_sampler.log 2>&1
ls -la gpurun_out/r2_prof_*.ncu-rep
