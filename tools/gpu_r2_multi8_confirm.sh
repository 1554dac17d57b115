#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29581 bench.py --gpus $N --steps 15 --warmup 3 --batch 8 --no-cpu-baseline > gpurun_out/r2_bench_8gpu_final.log 2>&1
echo "exit=$?"; grep -E "Error|Traceback" gpurun_out/r2_bench_8gpu_final.log | head -3
grep -E '^\{' gpurun_out/r2_bench_8gpu_final.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['ms_per_step'], 'ms', d['value'], 'samples/s  e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['clocks'], d['e2e']['clocks'])
"
