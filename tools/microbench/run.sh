#!/bin/bash
# Build (on the GPU box: nvcc is in the image) and run the micro-benchmarks; output goes to gpurun_out/microbench.log.
#   gpurun --timeout 600 -- 'bash tools/microbench/run.sh'
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for b in tmem_mufu handoff_latency mma_a_tmem; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Idreamvla_b200/csrc -o /tmp/$b tools/microbench/$b.cu || exit 1
done
{ for b in tmem_mufu handoff_latency mma_a_tmem; do echo "=== $b"; timeout 120 /tmp/$b; done; } 2>&1 | tee gpurun_out/microbench.log
