#!/bin/bash
# multi-GPU runs of round 2: N=$1 GPUs.  Overlap check (N=2 only), bench.py under torchrun with the exchange variants, the
# reference-on-GPU arm with DDP.
N=${1:-2}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "$N" = "2" ]; then
  CUDA_VISIBLE_DEVICES=0 DVLA_E2E_PROBE=1 timeout 300 python bench.py --steps 15 --warmup 3 --no-extras --no-cpu-baseline 2>&1 | grep -E "e2e probe|Error|Traceback" | tee gpurun_out/r2_e2e_probe.log
  DVLA_GEMM_SPLITK=0 timeout 300 $TR --master-port 29521 tools/ddp_overlap_check.py 2>&1 | grep -E "segment|reduced gradient|ranks agree|losses|max\||DDP_OVERLAP|Error|error" | cut -c1-300 | tee gpurun_out/r2_ddp_overlap_check.log
fi
port=29530
run() { name=$1; shift; port=$((port+1)); t0=$(date +%s)
  env "$@" timeout 240 $TR --master-port $port bench.py --gpus $N --steps ${STEPS:-15} --warmup 3 --batch 8 --no-cpu-baseline > gpurun_out/r2_bench_${N}gpu_$name.log 2>&1
  echo "$name exit=$? wall=$(( $(date +%s) - t0 ))s $(grep -E '^\{' gpurun_out/r2_bench_${N}gpu_$name.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['ms_per_step'], 'ms', d['value'], 'samples/s  e2e', d['e2e']['value'] if d.get('e2e') else None, d['e2e']['ms_per_step'] if d.get('e2e') else None)
")"; grep -E "Error|Traceback" gpurun_out/r2_bench_${N}gpu_$name.log | head -2 | cut -c1-200; }
run default DVLA_X=1
run nobudget DVLA_SM_BUDGET=0
if [ "$N" != "2" ]; then
  run nooverlap DVLA_AR_OVERLAP=0
  [ -n "$MORE" ] && run ctas8 DVLA_NCCL_CTAS=8
  [ -n "$MORE" ] && run ctas32 DVLA_NCCL_CTAS=32
  timeout 300 $TR --master-port 29560 tools/ddp_timeline.py 2>&1 | grep -E "^\[timeline\]|Error|Traceback" | cut -c1-220 | tee gpurun_out/r2_ddp_timeline_${N}gpu.log
fi
port=$((port+1)); t0=$(date +%s)
timeout 300 $TR --master-port $port bench.py --impl reference_gpu --gpus $N --steps 5 --warmup 3 --batch 8 > gpurun_out/r2_bench_${N}gpu_reference_gpu.log 2>&1
echo "reference_gpu exit=$? wall=$(( $(date +%s) - t0 ))s"; grep -E '^\{' gpurun_out/r2_bench_${N}gpu_reference_gpu.log | cut -c1-400; grep -E "Error|Traceback" gpurun_out/r2_bench_${N}gpu_reference_gpu.log | head -3 | cut -c1-200
