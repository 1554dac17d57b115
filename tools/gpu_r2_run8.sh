#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
t0=$(date +%s); timeout 1800 python -m pytest tests -q -m gpu -s -p no:cacheprovider > gpurun_out/r2_pytest8.log 2>&1; echo "pytest exit=$? wall=$(( $(date +%s) - t0 ))s"
grep -E "passed|failed|error" gpurun_out/r2_pytest8.log | tail -3; grep -E "^FAILED" gpurun_out/r2_pytest8.log | head -12; grep -E "fused sampler:|incremental\[|adamw:" gpurun_out/r2_pytest8.log | cut -c1-300
t0=$(date +%s); timeout 600 python tools/gpu_kernel_check.py attn > gpurun_out/r2_check_attn8.log 2>&1; echo "attn check exit=$? wall=$(( $(date +%s) - t0 ))s"; grep -E "FAIL|GROUP|key_bias|cat_broadcast|Lq10|Lq16 " gpurun_out/r2_check_attn8.log | cut -c1-150 | tail -20
t0=$(date +%s); timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench8.json 2> gpurun_out/r2_bench8.err; echo "bench exit=$? wall=$(( $(date +%s) - t0 ))s"
python - <<PY
import json
d = json.load(open("gpurun_out/r2_bench8.json"))
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
tail -3 gpurun_out/r2_bench8.err | cut -c1-300
