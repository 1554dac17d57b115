#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
t0=$(date +%s); timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python tools/debug_rollout.py > gpurun_out/r2_sanitizer_rollout.log 2>&1; echo "sanitizer exit=$? wall=$(( $(date +%s) - t0 ))s"
grep -E "Invalid|ok|ERROR SUMMARY|at 0x|in dvla|by thread" gpurun_out/r2_sanitizer_rollout.log | head -30
t0=$(date +%s); timeout 1500 python -m pytest tests -q -m gpu -s -p no:cacheprovider > gpurun_out/r2_pytest3.log 2>&1; echo "pytest exit=$? wall=$(( $(date +%s) - t0 ))s"
grep -E "passed|failed|error" gpurun_out/r2_pytest3.log | tail -3; grep -E "^FAILED" gpurun_out/r2_pytest3.log | head -12
t0=$(date +%s); timeout 600 python tools/gpu_kernel_check.py attn_perf > gpurun_out/r2_check_attn_perf.log 2>&1; echo "attn_perf exit=$? wall=$(( $(date +%s) - t0 ))s"
grep -E "FAIL|INFO attn_perf" gpurun_out/r2_check_attn_perf.log | tail -12
t0=$(date +%s); timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench3.json 2> gpurun_out/r2_bench3.err; echo "bench exit=$? wall=$(( $(date +%s) - t0 ))s"
python - <<PY
import json
d = json.load(open("gpurun_out/r2_bench3.json"))
print(d["ms_per_step"], "ms", d["value"], "samples/s", "e2e", d["e2e"], "gemm", d["roofline"]["gemm_ms_per_step"], "frac", d["roofline"]["frac"])
print({k: (v.get("value"), v.get("ms_per_step")) if isinstance(v, dict) else v for k, v in (d.get("extras") or {}).items()})
PY
tail -3 gpurun_out/r2_bench3.err | cut -c1-300
