#!/bin/bash
# ncu --set full of the round-2 HBM-bound kernels (LayerNorm backward v2, act' + bias gradient, LayerNorm forward) from one eager step
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
DVLA_BENCH_CUPROF=1 timeout 190 ncu --profile-from-start off --set full --clock-control none -k regex:"layernorm_bwd_v2|act_bwd_colsum|layernorm_fwd" -s 150 -c 9 -o gpurun_out/r2_prof_hbm2 \
   python bench.py --batch 8 --steps 1 --warmup 3 --no-graph --no-e2e --no-cpu-baseline --no-extras > gpurun_out/r2_ncu_hbm2.log 2>&1
echo "ncu exit=$?"
python tools/ncu_kernel_summary.py gpurun_out/r2_prof_hbm2.ncu-rep > gpurun_out/r2_ncu_hbm2.txt 2>&1; rm -f gpurun_out/r2_prof_hbm2.ncu-rep
grep -E "^[a-z_:<>A-Za-z0-9 ,()*]+\(|duration|DRAM read|DRAM write" gpurun_out/r2_ncu_hbm2.txt | cut -c1-120 | head -40
