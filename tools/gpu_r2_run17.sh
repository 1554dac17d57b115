#!/bin/bash
# round 2, run 17: K-split tail vs whole tiles on identical inputs; parity suite; bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
DVLA_GEMM_TAIL=0 timeout 300 python tools/gemm_tail_compare.py dump /tmp/tail0.pt && timeout 300 python tools/gemm_tail_compare.py dump /tmp/tail1.pt && python tools/gemm_tail_compare.py compare /tmp/tail0.pt /tmp/tail1.pt 2>&1 | cut -c1-260 | tee gpurun_out/r2_gemm_tail_compare.log
for v in 0 1; do
DVLA_GEMM_TAIL=$v timeout 900 python -m pytest tests/test_full_depth_gpu.py -q -s -p no:cacheprovider -k libero_full 2>&1 | grep -E "DiT loss|gripper|arm actions|comparisons|passed|failed" | cut -c1-170
done
t0=$(date +%s); timeout 1800 python -m pytest tests -q -m gpu -s -p no:cacheprovider > gpurun_out/r2_pytest17.log 2>&1; echo "pytest exit=$? wall=$(( $(date +%s) - t0 ))s"
grep -E "passed|failed|error" gpurun_out/r2_pytest17.log | tail -3; grep -E "^FAILED" gpurun_out/r2_pytest17.log | head -12
t0=$(date +%s); timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench17.json 2> gpurun_out/r2_bench17.err; echo "bench exit=$? wall=$(( $(date +%s) - t0 ))s"
python - <<PY
import json
d = json.load(open("gpurun_out/r2_bench17.json"))
print(d["ms_per_step"], "ms", d["value"], "samples/s", "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "gemm", d["roofline"]["gemm_ms_per_step"], "frac", d["roofline"]["frac"], d["clocks"]["sm_mhz"], d["e2e"]["clocks"]["sm_mhz"])
ex = d.get("extras") or {}
for k, v in ex.items():
    if k == "action_latency" and isinstance(v, dict):
        print(k, {kk: (vv.get("p50"), vv.get("p99")) for kk, vv in v.items() if isinstance(vv, dict)}, v.get("error"))
    elif isinstance(v, dict):
        print(k, v.get("value"), v.get("ms_per_step"), v.get("error"))
    else:
        print(k, v)
PY
