#!/bin/bash
# round 2 final validation: what the driver runs at round end (smoke, pytest -m gpu, default bench) + the launch list of one step
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
t0=$(date +%s); timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2; echo "smoke exit=$? wall=$(( $(date +%s) - t0 ))s"
t0=$(date +%s); timeout 1200 python -m pytest tests -q -m gpu -s -p no:cacheprovider > gpurun_out/r2_pytest_final.log 2>&1; echo "pytest exit=$? wall=$(( $(date +%s) - t0 ))s"
grep -E "passed|failed|error" gpurun_out/r2_pytest_final.log | tail -3; grep -E "^FAILED" gpurun_out/r2_pytest_final.log | head -12
t0=$(date +%s); timeout 900 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo "bench exit=$? wall=$(( $(date +%s) - t0 ))s"
python - <<PY
import json
d = json.load(open("gpurun_out/r2_bench_final.json"))
print(d["ms_per_step"], "ms", d["value"], "samples/s", "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "gemm", d["roofline"]["gemm_ms_per_step"], "frac", d["roofline"]["frac"], d["roofline"].get("traffic"), d["clocks"], d["e2e"]["clocks"]["sm_mhz"], "launches", d["gpu_launches"])
print("cpu_baseline", d.get("cpu_baseline"))
ex = d.get("extras") or {}
for k, v in ex.items():
    if k == "action_latency" and isinstance(v, dict):
        print(k, {kk: (vv.get("p50"), vv.get("p99")) for kk, vv in v.items() if isinstance(vv, dict)}, v.get("error"))
    elif isinstance(v, dict):
        print(k, v.get("value"), v.get("ms_per_step"), v.get("error"))
    else:
        print(k, v)
PY
DVLA_BENCH_CUPROF=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_final.csv python bench.py --batch 8 --steps 1 --warmup 3 --no-graph --no-e2e --no-cpu-baseline --no-extras > gpurun_out/r2_ncu_bench_final.log 2>&1
python tools/ncu_summarize.py gpurun_out/r2_launches_final.csv > gpurun_out/r2_launches_step_eager_b8_final.txt 2>&1; head -30 gpurun_out/r2_launches_step_eager_b8_final.txt; rm -f gpurun_out/r2_launches_final.csv
