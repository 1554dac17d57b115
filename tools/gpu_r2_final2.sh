#!/bin/bash
# round 2, last validation: smoke, the whole GPU suite (incl. device-side augmentation and the flow-matching head), short bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
t0=$(date +%s); timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2; echo "smoke exit=$? wall=$(( $(date +%s) - t0 ))s"
t0=$(date +%s); timeout 1500 python -m pytest tests -q -m gpu -s -p no:cacheprovider > gpurun_out/r2_pytest_final2.log 2>&1; echo "pytest exit=$? wall=$(( $(date +%s) - t0 ))s"
grep -E "passed|failed|error" gpurun_out/r2_pytest_final2.log | tail -3; grep -E "^FAILED|^ERROR" gpurun_out/r2_pytest_final2.log | head -12; grep -E "shift_crop \[" gpurun_out/r2_pytest_final2.log | head -2
grep -E "^E  " gpurun_out/r2_pytest_final2.log | head -20 | cut -c1-250
t0=$(date +%s); timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2_bench_final2.json 2> gpurun_out/r2_bench_final2.err; echo "bench exit=$? wall=$(( $(date +%s) - t0 ))s"
python - <<PY
import json
d = json.load(open("gpurun_out/r2_bench_final2.json"))
print(d["ms_per_step"], "ms", d["value"], "samples/s", "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "gemm", d["roofline"]["gemm_ms_per_step"], "frac", d["roofline"]["frac"])
PY
