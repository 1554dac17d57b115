#!/bin/bash
# round 2, run 1: the new parity tests + baseline numbers of the unchanged kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r2_gpu_info.txt
t0=$(date +%s); timeout 1500 python -m pytest tests -q -m gpu -s -p no:cacheprovider > gpurun_out/r2_pytest1.log 2>&1; echo "pytest exit=$? wall=$(( $(date +%s) - t0 ))s"
grep -E "passed|failed|error" gpurun_out/r2_pytest1.log | tail -3
C4="--finetune_type calvin --precision bf16 --phase evaluate --num_resampler_query 16 --num_obs_token_per_image 9 --transformer_layers 24 --hidden_dim 1024 --transformer_heads 16 --action_pred_steps 3 --sequence_length 10 --obs_pred --depth_pred --sam_feat_pred --use_dit_head --attn_implementation sdpa"
t0=$(date +%s); timeout 600 python eval_calvin.py $C4 --synthetic_rollout_steps 300 > gpurun_out/r2_latency_full.json 2> gpurun_out/r2_latency_full.err; echo "latency full exit=$? wall=$(( $(date +%s) - t0 ))s"; cat gpurun_out/r2_latency_full.json
t0=$(date +%s); timeout 600 python eval_calvin.py $C4 --synthetic_rollout_steps 300 --incremental_rollout > gpurun_out/r2_latency_inc.json 2> gpurun_out/r2_latency_inc.err; echo "latency inc exit=$? wall=$(( $(date +%s) - t0 ))s"; cat gpurun_out/r2_latency_inc.json; tail -3 gpurun_out/r2_latency_inc.err
t0=$(date +%s); timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_b8_base.json 2> gpurun_out/r2_bench_b8_base.err; echo "bench exit=$? wall=$(( $(date +%s) - t0 ))s"
cut -c1-1200 gpurun_out/r2_bench_b8_base.json
