#!/bin/bash
# round 2, run 2: parity tests again + small-sequence attention kernels + decoder algebra A/B + op profile
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
t0=$(date +%s); timeout 1500 python -m pytest tests -q -m gpu -s -p no:cacheprovider > gpurun_out/r2_pytest2.log 2>&1; echo "pytest exit=$? wall=$(( $(date +%s) - t0 ))s"
grep -E "passed|failed|error" gpurun_out/r2_pytest2.log | tail -3; grep -E "^FAILED" gpurun_out/r2_pytest2.log | head
t0=$(date +%s); timeout 600 python tools/gpu_kernel_check.py attn attn_perf > gpurun_out/r2_check_attn.log 2>&1; echo "kernel check exit=$? wall=$(( $(date +%s) - t0 ))s"
grep -E "FAIL|INFO attn_perf|us" gpurun_out/r2_check_attn.log | tail -40
for tag in "base:" "noalg:DVLA_DECODER_ALGEBRA=0" "nosmall:DVLA_ATTN_SMALL=0"; do
  name=${tag%%:*}; envs=${tag#*:}
  t0=$(date +%s); env $envs timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_bench2_$name.json 2> gpurun_out/r2_bench2_$name.err; echo "bench $name exit=$? wall=$(( $(date +%s) - t0 ))s"
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2_bench2_$name.json"))
    print("$name", d["ms_per_step"], "ms", d["value"], "samples/s", "gemm", d["roofline"]["gemm_ms_per_step"], "frac", d["roofline"]["frac"], "loss", d["config"]["final_loss"])
except Exception as e:
    print("$name failed", e)
PY
done
t0=$(date +%s); timeout 600 python tools/torch_op_profile.py --batch 8 > gpurun_out/r2_torch_ops_b8.txt 2>&1; echo "op profile exit=$? wall=$(( $(date +%s) - t0 ))s"
tail -32 gpurun_out/r2_torch_ops_b8.txt | cut -c1-260
