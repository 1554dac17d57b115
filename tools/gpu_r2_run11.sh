#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
DVLA_DIT_TRACE=1 timeout 200 python tools/prof_sampler.py 2>&1 | grep -E "sampler|dit trace" | head -6 | cut -c1-1200 | tee gpurun_out/r2_sampler_timing3.log
t0=$(date +%s); timeout 900 python -m pytest tests/test_rollout_gpu.py -q -s -p no:cacheprovider -k "fused or libero_wrapper" > gpurun_out/r2_pytest11.log 2>&1; echo "pytest exit=$? wall=$(( $(date +%s) - t0 ))s"
grep -E "passed|failed|error" gpurun_out/r2_pytest11.log | tail -2; grep -E "^FAILED" gpurun_out/r2_pytest11.log | head; grep -E "fused sampler:" gpurun_out/r2_pytest11.log | cut -c1-300
timeout 300 ncu --set full --clock-control none -k regex:dit_ddim -c 1 -o gpurun_out/r2_prof_sampler_v2 python tools/debug_rollout.py > gpurun_out/r2_ncu_sampler.log 2>&1
python tools/ncu_kernel_summary.py gpurun_out/r2_prof_sampler_v2.ncu-rep > gpurun_out/r2_prof_sampler_v2.txt 2>&1; rm -f gpurun_out/r2_prof_sampler_v2.ncu-rep; cat gpurun_out/r2_prof_sampler_v2.txt | cut -c1-140
C4="--finetune_type calvin --precision bf16 --phase evaluate --num_resampler_query 16 --num_obs_token_per_image 9 --transformer_layers 24 --hidden_dim 1024 --transformer_heads 16 --action_pred_steps 3 --sequence_length 10 --obs_pred --depth_pred --sam_feat_pred --use_dit_head --attn_implementation sdpa"
timeout 300 python eval_calvin.py $C4 --synthetic_rollout_steps 300 --incremental_rollout 2>/dev/null | tee gpurun_out/r2_latency_inc_v2.json
DVLA_DIT_FUSED=0 timeout 300 python eval_calvin.py $C4 --synthetic_rollout_steps 300 --incremental_rollout 2>/dev/null | tee gpurun_out/r2_latency_inc_modsampler.json
