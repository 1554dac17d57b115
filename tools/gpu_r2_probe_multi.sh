#!/bin/bash
# e2e pipeline probe under data parallelism: which ingredient of the input pipeline interacts with the overlapped all-reduce
N=${1:-4}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
DVLA_E2E_PROBE=1 timeout 400 $TR --master-port 29571 bench.py --gpus $N --steps 12 --warmup 3 --batch 8 --no-cpu-baseline > gpurun_out/r2_e2e_probe_${N}gpu.log 2>&1
echo "exit=$?"; grep -E "e2e probe|Error|Traceback" gpurun_out/r2_e2e_probe_${N}gpu.log | cut -c1-200
grep -E '^\{' gpurun_out/r2_e2e_probe_${N}gpu.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['ms_per_step'], 'ms', d['value'], 'samples/s  e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
"
