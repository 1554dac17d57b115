#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 120 python tools/debug_rollout2.py > gpurun_out/r2_dbg_$name.log 2>&1; echo "$name exit=$? last: $(grep -E 'ok|OK' gpurun_out/r2_dbg_$name.log | tail -1) | err: $(grep -E 'Error|error' gpurun_out/r2_dbg_$name.log | head -1 | cut -c1-160)"; }
run default DBG_X=1
run default2 DBG_X=2
run blocking CUDA_LAUNCH_BLOCKING=1
run nosmall DVLA_ATTN_SMALL=0
run nograph DBG_GRAPH=0
run nofull DBG_FULL=0
run prune DBG_PRUNE=1
run legacyattn DVLA_ATTN_FWD=legacy
t0=$(date +%s); timeout 300 compute-sanitizer --tool racecheck --print-limit 6 python tools/debug_rollout.py > gpurun_out/r2_racecheck.log 2>&1; echo "racecheck exit=$? wall=$(( $(date +%s) - t0 ))s"; grep -E "RACECHECK SUMMARY|hazard|in dvla|in void" gpurun_out/r2_racecheck.log | head -12
t0=$(date +%s); timeout 300 compute-sanitizer --tool synccheck --print-limit 6 python tools/debug_rollout.py > gpurun_out/r2_synccheck.log 2>&1; echo "synccheck exit=$? wall=$(( $(date +%s) - t0 ))s"; grep -E "ERROR SUMMARY|Barrier|in dvla|in void" gpurun_out/r2_synccheck.log | head -12
t0=$(date +%s); timeout 900 python -m pytest tests/test_train_step_gpu.py tests/test_kernels_gpu.py -q -m gpu -s -p no:cacheprovider > gpurun_out/r2_pytest5.log 2>&1; echo "pytest exit=$? wall=$(( $(date +%s) - t0 ))s"; grep -E "passed|failed" gpurun_out/r2_pytest5.log | tail -2; grep -E "^FAILED" gpurun_out/r2_pytest5.log | head
t0=$(date +%s); timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2_bench5.json 2> gpurun_out/r2_bench5.err; echo "bench exit=$? wall=$(( $(date +%s) - t0 ))s"
python - <<PY
import json
d = json.load(open("gpurun_out/r2_bench5.json"))
print(d["ms_per_step"], "ms", d["value"], "samples/s", "e2e", d["e2e"], "gemm", d["roofline"]["gemm_ms_per_step"], "frac", d["roofline"]["frac"])
PY
