#!/bin/bash
# round 2, run 15: K-split tail of the GEMMs (fp32 reduction through the workspace), fp32 split-K for weight gradients
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for g in gemm_tail gemm_basic gemm_splitk gemm_epilogue gemm_big; do
  timeout 600 python tools/gpu_kernel_check.py $g > gpurun_out/r2_check15_$g.log 2>&1; echo "$g exit=$?"
  grep -E "FAIL|GROUP|TFLOP|Error|error" gpurun_out/r2_check15_$g.log | cut -c1-200 | head -40
done
echo "--- DVLA_GEMM_TAIL=0"
DVLA_GEMM_TAIL=0 timeout 600 python tools/gpu_kernel_check.py gemm_tail 2>&1 | grep -E "TFLOP|FAIL" | cut -c1-200
DVLA_GEMM_TAIL=0 timeout 600 python tools/gpu_kernel_check.py gemm_splitk 2>&1 | grep -E "TFLOP|FAIL" | cut -c1-200
t0=$(date +%s); timeout 1500 python -m pytest tests -q -m gpu -x -p no:cacheprovider -k "not rollout and not train_entry" > gpurun_out/r2_pytest15.log 2>&1; echo "pytest exit=$? wall=$(( $(date +%s) - t0 ))s"
grep -E "passed|failed|error" gpurun_out/r2_pytest15.log | tail -3; grep -E "^FAILED" gpurun_out/r2_pytest15.log | head -12
for v in 1 0; do
t0=$(date +%s); DVLA_GEMM_TAIL=$v timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r2_bench15_tail$v.json 2> gpurun_out/r2_bench15_tail$v.err; echo "bench tail=$v exit=$? wall=$(( $(date +%s) - t0 ))s"
python - <<PY
import json
d = json.load(open("gpurun_out/r2_bench15_tail$v.json"))
print(d["ms_per_step"], "ms", d["value"], "samples/s", "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "gemm", d["roofline"]["gemm_ms_per_step"], "frac", d["roofline"]["frac"], d["clocks"]["sm_mhz"], d["e2e"]["clocks"]["sm_mhz"], "loss", d.get("final_loss"))
PY
done
