#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
DVLA_DIT_TRACE=1 timeout 200 python tools/prof_sampler.py 2>&1 | grep -E "sampler|dit trace" | head -8 | cut -c1-1500 | tee gpurun_out/r2_sampler_timing2.log
LAUNCHES=0 bash tools/gpu_r2_prof.sh 2>&1 | tail -20
