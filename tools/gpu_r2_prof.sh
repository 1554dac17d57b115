#!/bin/bash
# round-2 ncu evidence: launch list of ONE eager train step (B=8, exactly the timed region) + full captures of the dominant kernels.
# Reports are summarised ON THE BOX and deleted (gpurun brings back at most 64 MiB).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
sum() { python tools/ncu_kernel_summary.py gpurun_out/$1.ncu-rep > gpurun_out/$1.txt 2>&1; rm -f gpurun_out/$1.ncu-rep; wc -l gpurun_out/$1.txt; }
if [ "${LAUNCHES:-1}" = "1" ]; then
echo "=== launch list"
DVLA_BENCH_CUPROF=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_step_b8.csv \
   python bench.py --batch 8 --steps 1 --warmup 3 --no-graph --no-e2e --no-cpu-baseline --no-extras > gpurun_out/r2_ncu_bench.log 2>&1
python tools/ncu_summarize.py gpurun_out/r2_launches_step_b8.csv | tee gpurun_out/r2_launches_step_b8_summary.txt | head -12
fi
echo "=== ncu full: gemm"
timeout 600 ncu --set full --clock-control none -k regex:gemm_tcgen05 -s 2 -c 7 -o gpurun_out/r2_prof_gemm python tools/prof_gemm.py gemm > gpurun_out/r2_ncu_gemm.log 2>&1; sum r2_prof_gemm
echo "=== ncu full: attention (gpt2 shape, mask + dropout)"
timeout 600 ncu --set full --clock-control none -k regex:attn_.*_ws -s 2 -c 3 -o gpurun_out/r2_prof_attn_gpt python tools/prof_attn.py gpt bwd > gpurun_out/r2_ncu_attn_gpt.log 2>&1; sum r2_prof_attn_gpt
echo "=== ncu full: HBM-bound kernels (LayerNorm fwd/bwd, act_bwd, colsum, AdamW) from one eager step"
DVLA_BENCH_CUPROF=1 timeout 900 ncu --profile-from-start off --set full --clock-control none -k regex:"layernorm_bwd|layernorm_fwd|act_bwd|colsum|adamw|sumsq|cat_broadcast" -s 300 -c 14 -o gpurun_out/r2_prof_hbm \
   python bench.py --batch 8 --steps 1 --warmup 3 --no-graph --no-e2e --no-cpu-baseline --no-extras > gpurun_out/r2_ncu_hbm.log 2>&1; sum r2_prof_hbm
DVLA_BENCH_CUPROF=1 timeout 900 ncu --profile-from-start off --set full --clock-control none -k regex:"adamw|sumsq" -c 2 -o gpurun_out/r2_prof_adamw \
   python bench.py --batch 8 --steps 1 --warmup 3 --no-graph --no-e2e --no-cpu-baseline --no-extras > gpurun_out/r2_ncu_adamw.log 2>&1; sum r2_prof_adamw
echo "=== ncu full: sampler megakernel"
timeout 600 ncu --set full --clock-control none -k regex:dit_ddim -c 1 -o gpurun_out/r2_prof_sampler python tools/debug_rollout.py > gpurun_out/r2_ncu_sampler.log 2>&1; sum r2_prof_sampler
du -sh gpurun_out
