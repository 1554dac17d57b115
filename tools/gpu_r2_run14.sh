#!/bin/bash
# round 2, run 14: global heavy-first CTA order of the ws attention kernels, act_bwd + bias-gradient fusion
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python tools/gpu_kernel_check.py act_bwd_colsum > gpurun_out/r2_check_actbwd.log 2>&1; echo "act_bwd_colsum exit=$?"; grep -E "FAIL|GROUP|fused " gpurun_out/r2_check_actbwd.log | cut -c1-170
timeout 600 python tools/gpu_kernel_check.py attn > gpurun_out/r2_check_attn14.log 2>&1; echo "attn exit=$?"; grep -E "FAIL|GROUP" gpurun_out/r2_check_attn14.log | cut -c1-170
timeout 600 python tools/gpu_kernel_check.py attn_perf > gpurun_out/r2_check_attn_perf14.log 2>&1; grep -E "FAIL|GROUP|us" gpurun_out/r2_check_attn_perf14.log | cut -c1-200 | head -30
t0=$(date +%s); timeout 1500 python -m pytest tests -q -m gpu -x -p no:cacheprovider -k "not rollout and not train_entry" > gpurun_out/r2_pytest14.log 2>&1; echo "pytest exit=$? wall=$(( $(date +%s) - t0 ))s"
grep -E "passed|failed|error" gpurun_out/r2_pytest14.log | tail -3; grep -E "^FAILED" gpurun_out/r2_pytest14.log | head -12
for i in 1 2; do
t0=$(date +%s); timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r2_bench14_$i.json 2> gpurun_out/r2_bench14_$i.err; echo "bench exit=$? wall=$(( $(date +%s) - t0 ))s"
python - <<PY
import json
d = json.load(open("gpurun_out/r2_bench14_$i.json"))
print(d["ms_per_step"], "ms", d["value"], "samples/s", "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "gemm", d["roofline"]["gemm_ms_per_step"], "frac", d["roofline"]["frac"], d["clocks"], d["e2e"]["clocks"])
PY
done
