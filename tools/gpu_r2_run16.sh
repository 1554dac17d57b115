#!/bin/bash
# round 2, run 16: K-split tail with per-split fp32 slices, fp32 split-K weight gradients, direct-compare dropout in attention
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for g in gemm_tail gemm_splitk gemm_basic gemm_epilogue attn; do
  timeout 600 python tools/gpu_kernel_check.py $g > gpurun_out/r2_check16_$g.log 2>&1; echo "$g exit=$?"
  grep -E "FAIL|GROUP|TFLOP|Error|error" gpurun_out/r2_check16_$g.log | grep -v "^PASS gemm M" | cut -c1-200 | head -40
done
timeout 600 python tools/gpu_kernel_check.py attn_perf 2>&1 | grep -E "gpt2|FAIL" | cut -c1-200 | tee gpurun_out/r2_check16_attn_perf.log
echo "--- DVLA_GEMM_TAIL=0 / DVLA_GEMM_SPLITK=atomic"
DVLA_GEMM_TAIL=0 timeout 600 python tools/gpu_kernel_check.py gemm_tail 2>&1 | grep -E "TFLOP|FAIL" | cut -c1-200
DVLA_GEMM_SPLITK=atomic timeout 600 python tools/gpu_kernel_check.py gemm_splitk 2>&1 | grep -E "TFLOP|FAIL" | cut -c1-200
t0=$(date +%s); timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider -k "not rollout and not train_entry" > gpurun_out/r2_pytest16.log 2>&1; echo "pytest exit=$? wall=$(( $(date +%s) - t0 ))s"
grep -E "passed|failed|error" gpurun_out/r2_pytest16.log | tail -3; grep -E "^FAILED" gpurun_out/r2_pytest16.log | head -12
for v in "1 fp32" "0 fp32" "1 atomic"; do
set -- $v
t0=$(date +%s); DVLA_GEMM_TAIL=$1 DVLA_GEMM_SPLITK=$2 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r2_bench16_tail$1_$2.json 2> gpurun_out/r2_bench16_tail$1_$2.err; echo "bench tail=$1 splitk=$2 exit=$? wall=$(( $(date +%s) - t0 ))s"
python - <<PY
import json
d = json.load(open("gpurun_out/r2_bench16_tail$1_$2.json"))
print(d["ms_per_step"], "ms", d["value"], "samples/s", "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "gemm", d["roofline"]["gemm_ms_per_step"], "frac", d["roofline"]["frac"], d["clocks"]["sm_mhz"], d["e2e"]["clocks"]["sm_mhz"])
PY
done
