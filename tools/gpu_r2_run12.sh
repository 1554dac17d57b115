#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
DVLA_DIT_TRACE=1 timeout 200 python tools/prof_sampler.py 2>&1 | grep -E "sampler|dit trace" | head -4 | cut -c1-1000 | tee gpurun_out/r2_sampler_timing4.log
timeout 300 python tools/gpu_kernel_check.py norm > gpurun_out/r2_check_norm.log 2>&1; echo "norm check exit=$?"; grep -E "FAIL|GROUP|fused|4099|2600|3000" gpurun_out/r2_check_norm.log | cut -c1-150
DVLA_LN_BWD=v1 timeout 300 python tools/gpu_kernel_check.py norm 2>&1 | grep -E "fused dx" | cut -c1-150
t0=$(date +%s); timeout 1800 python -m pytest tests -q -m gpu -s -p no:cacheprovider > gpurun_out/r2_pytest12.log 2>&1; echo "pytest exit=$? wall=$(( $(date +%s) - t0 ))s"
grep -E "passed|failed|error" gpurun_out/r2_pytest12.log | tail -3; grep -E "^FAILED" gpurun_out/r2_pytest12.log | head -12; grep -E "fused sampler:" gpurun_out/r2_pytest12.log | cut -c1-300
t0=$(date +%s); timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench12.json 2> gpurun_out/r2_bench12.err; echo "bench exit=$? wall=$(( $(date +%s) - t0 ))s"
python - <<PY
import json
d = json.load(open("gpurun_out/r2_bench12.json"))
print(d["ms_per_step"], "ms", d["value"], "samples/s", "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "gemm", d["roofline"]["gemm_ms_per_step"], "frac", d["roofline"]["frac"])
ex = d.get("extras") or {}
for k, v in ex.items():
    if k == "action_latency" and isinstance(v, dict):
        print(k, {kk: (vv.get("p50"), vv.get("p99")) for kk, vv in v.items() if isinstance(vv, dict)}, v.get("error"))
    elif isinstance(v, dict):
        print(k, v.get("value"), v.get("ms_per_step"), v.get("error"))
    else:
        print(k, v)
PY
