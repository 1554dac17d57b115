#!/bin/bash
# model-level GPU run: parity tests, bench at both measured batch sizes, per-shape GEMM table of one step
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== pytest train step + model"; timeout 900 python -m pytest tests/test_train_step_gpu.py tests/test_model_gpu.py -x -q -m gpu 2>&1 | tail -5
echo "=== bench B=2"; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --batch 2 2>&1 | tail -1 | cut -c1-300
echo "=== bench B=8 with per-shape GEMM dump"; DVLA_BENCH_DUMP=1 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --batch 8 2> gpurun_out/gemm_dump_b8.txt | tail -1 | cut -c1-300 | tee gpurun_out/bench_b8.log
grep "^\[gemm\]" gpurun_out/gemm_dump_b8.txt | sort -k 18 -n -r | head -28
