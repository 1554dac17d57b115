#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for c in all 74 37; do
  if [ $c = all ]; then timeout 200 python tools/prof_sampler.py 2>&1 | grep sampler; else DVLA_DIT_CTAS=$c timeout 200 python tools/prof_sampler.py 2>&1 | grep "fused"; fi
done | tee gpurun_out/r2_sampler_timing.log
t0=$(date +%s); timeout 900 python -m pytest tests/test_rollout_gpu.py -q -s -p no:cacheprovider > gpurun_out/r2_pytest9.log 2>&1; echo "pytest rollout exit=$? wall=$(( $(date +%s) - t0 ))s"
grep -E "passed|failed|error" gpurun_out/r2_pytest9.log | tail -2; grep -E "^FAILED" gpurun_out/r2_pytest9.log | head; grep -E "fused sampler:|incremental\[" gpurun_out/r2_pytest9.log | cut -c1-300
bash tools/gpu_r2_prof.sh 2>&1 | tail -60
