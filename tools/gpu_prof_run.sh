#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== ncu full: attention (decoder shape), tc kernels"
DVLA_ATTN_FWD=tc DVLA_ATTN_BWD=tc timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_.*tc_kernel -s 3 -c 4 -o gpurun_out/prof_attn_tc python tools/prof_gemm.py attn_dec > gpurun_out/ncu_attn.log 2>&1
echo "=== ncu full: gemm 2cta"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 2 -c 4 -o gpurun_out/prof_gemm2 python tools/prof_gemm.py gemm > gpurun_out/ncu_gemm2.log 2>&1
ls -la gpurun_out/*.ncu-rep
echo "=== ncu launch list (graph-less step, B=2)"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -s 6500 -c 2400 --csv --log-file gpurun_out/launches_step.csv \
   python bench.py --steps 1 --warmup 3 --no-graph --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
python tools/ncu_summarize.py gpurun_out/launches_step.csv | tee gpurun_out/launches_step_summary.txt | head -32
echo "=== eval latency"
timeout 600 python eval_calvin.py --phase evaluate --precision bf16 --sequence_length 10 --num_resampler_query 16 --num_obs_token_per_image 9 --action_pred_steps 3 --transformer_layers 24 --hidden_dim 1024 --transformer_heads 16 --obs_pred --depth_pred --sam_feat_pred --use_dit_head --attn_implementation sdpa --synthetic_rollout_steps 200 2>&1 | tail -3 | tee gpurun_out/eval_latency.log
