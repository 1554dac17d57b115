#!/bin/bash
# gpurun with retries while the pod has no free slot (exit 3 / "transient"): tools/gpurun_retry.sh <timeout_s> [--gpus N] -- '<cmd>'
T=$1; shift
for i in $(seq 1 20); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$T" "$@" 2>&1); rc=$?
  echo "$out" | tail -60
  if echo "$out" | grep -q "status=transient\|no box\|retry in a few minutes"; then echo "[retry $i] sleeping 120 s"; sleep 120; continue; fi
  exit $rc
done
exit 3
