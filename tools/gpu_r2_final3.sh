#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
t0=$(date +%s); timeout 1500 python -m pytest tests -q -m gpu -s -p no:cacheprovider > gpurun_out/r2_pytest_final3.log 2>&1; echo "pytest exit=$? wall=$(( $(date +%s) - t0 ))s"
grep -E "passed|failed|error" gpurun_out/r2_pytest_final3.log | tail -3; grep -E "^FAILED|^ERROR" gpurun_out/r2_pytest_final3.log | head -12; grep -E "shift_crop \[" gpurun_out/r2_pytest_final3.log | head -2
grep -E "^E  " gpurun_out/r2_pytest_final3.log | head -20 | cut -c1-250
