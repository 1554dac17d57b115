#!/bin/bash
# what the driver does at round end: smoke, pytest -m gpu, default bench (both arms), each with its wall time
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
t0=$(date +%s); timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2; echo "smoke exit=$? wall=$(( $(date +%s) - t0 ))s"
t0=$(date +%s); timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -2; echo "pytest wall=$(( $(date +%s) - t0 ))s"
t0=$(date +%s); timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench exit=$? wall=$(( $(date +%s) - t0 ))s"
tail -c 3000 gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err | cut -c1-300
t0=$(date +%s); timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference exit=$? wall=$(( $(date +%s) - t0 ))s"
tail -c 1500 gpurun_out/bench_reference.json; tail -3 gpurun_out/bench_reference.err | cut -c1-300
