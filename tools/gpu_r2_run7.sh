#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python tools/debug_rollout3.py > gpurun_out/r2_dbg3.log 2>&1; echo "dbg3 exit=$?"; grep "^step" gpurun_out/r2_dbg3.log | cut -c1-700; grep -E "Error|error" gpurun_out/r2_dbg3.log | head -3
t0=$(date +%s); timeout 1500 python -m pytest tests -q -m gpu -s -p no:cacheprovider --deselect "tests/test_rollout_gpu.py::test_incremental_rollout_matches_full_window" --deselect "tests/test_rollout_gpu.py::test_libero_wrapper_gripper_width" > gpurun_out/r2_pytest7.log 2>&1; echo "pytest exit=$? wall=$(( $(date +%s) - t0 ))s"
grep -E "passed|failed|error" gpurun_out/r2_pytest7.log | tail -3; grep -E "^FAILED" gpurun_out/r2_pytest7.log | head -12
t0=$(date +%s); timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2_bench7.json 2> gpurun_out/r2_bench7.err; echo "bench exit=$? wall=$(( $(date +%s) - t0 ))s"
python - <<PY
import json
d = json.load(open("gpurun_out/r2_bench7.json"))
print(d["ms_per_step"], "ms", d["value"], "samples/s", "e2e", d["e2e"], "gemm", d["roofline"]["gemm_ms_per_step"], "frac", d["roofline"]["frac"])
PY
