#!/bin/bash
# 2-GPU validation of the backward-overlapped gradient all-reduce (eager equivalence check, then captured in bench.py)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
DVLA_GEMM_SPLITK=0 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tools/ddp_overlap_check.py 2>&1 | grep -E "segments|reduced gradient|ranks agree|losses|max\||DDP_OVERLAP|Error|error" | cut -c1-300 | tee gpurun_out/ddp_overlap_check.log
for ov in 1 0; do
  echo "=== bench 2 GPUs B=8 overlap=$ov"
  t0=$(date +%s)
  DVLA_AR_OVERLAP=$ov timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$ov bench.py --gpus 2 --steps 10 --warmup 3 --batch 8 --no-cpu-baseline --no-e2e > gpurun_out/bench_2gpu_ov$ov.log 2>&1
  echo "exit=$? wall=$(( $(date +%s) - t0 ))s"
  grep -E "^\{|Error|Traceback" gpurun_out/bench_2gpu_ov$ov.log | cut -c1-260
done
