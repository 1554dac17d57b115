#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
t0=$(date +%s); timeout 600 compute-sanitizer --tool memcheck --print-limit 8 python -m pytest tests/test_rollout_gpu.py -x -q -p no:cacheprovider -k "incremental and libero_dit-False" > gpurun_out/r2_sanitizer_test.log 2>&1; echo "sanitizer exit=$? wall=$(( $(date +%s) - t0 ))s"
grep -E "Invalid|ERROR SUMMARY|at 0x|in dvla|in void|by thread|Address|passed|failed" gpurun_out/r2_sanitizer_test.log | head -30
t0=$(date +%s); timeout 1500 python -m pytest tests -q -m gpu -s -p no:cacheprovider --deselect "tests/test_rollout_gpu.py::test_incremental_rollout_matches_full_window" --deselect "tests/test_rollout_gpu.py::test_libero_wrapper_gripper_width" > gpurun_out/r2_pytest4.log 2>&1; echo "pytest exit=$? wall=$(( $(date +%s) - t0 ))s"
grep -E "passed|failed|error" gpurun_out/r2_pytest4.log | tail -3; grep -E "^FAILED" gpurun_out/r2_pytest4.log | head -12
t0=$(date +%s); timeout 600 python tools/gpu_kernel_check.py gemm_big > gpurun_out/r2_check_gemm_big.log 2>&1; echo "gemm_big exit=$? wall=$(( $(date +%s) - t0 ))s"
grep -E "FAIL|PASS|INFO" gpurun_out/r2_check_gemm_big.log | tail -16 | cut -c1-140
t0=$(date +%s); DVLA_BENCH_DUMP=1 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2_bench4.json 2> gpurun_out/r2_bench4.err; echo "bench exit=$? wall=$(( $(date +%s) - t0 ))s"
python - <<PY
import json
d = json.load(open("gpurun_out/r2_bench4.json"))
print(d["ms_per_step"], "ms", d["value"], "samples/s", "e2e", d["e2e"], "gemm", d["roofline"]["gemm_ms_per_step"], "frac", d["roofline"]["frac"])
PY
grep "\[gemm\]" gpurun_out/r2_bench4.err | sort -t' ' -k1 | awk '{print}' > gpurun_out/r2_gemm_dump_b8.txt; wc -l gpurun_out/r2_gemm_dump_b8.txt
