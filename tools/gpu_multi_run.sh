#!/bin/bash
# 2-GPU validation: backward-overlapped gradient all-reduce == single all-reduce (eager), then bench.py under torchrun
# (captured step with the overlapped all-reduce, clean tear-down).  STEPS / CHECK=0 shorten the run.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
if [ "${CHECK:-1}" = "1" ]; then
  DVLA_GEMM_SPLITK=0 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tools/ddp_overlap_check.py 2>&1 | grep -E "segments|reduced gradient|ranks agree|losses|max\||DDP_OVERLAP|Error|error" | cut -c1-300 | tee gpurun_out/ddp_overlap_check.log
fi
t0=$(date +%s)
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps ${STEPS:-10} --warmup 3 --batch 8 --no-cpu-baseline --no-e2e > gpurun_out/bench_2gpu.log 2>&1
echo "exit=$? wall=$(( $(date +%s) - t0 ))s"
grep -E "^\{|Error|Traceback" gpurun_out/bench_2gpu.log | cut -c1-400
