#!/bin/bash
# Retry a gpurun command until the pod gives a slot (exit code 3 / transient = nothing charged).
# Usage: tools/gpu_try.sh <timeout_s> '<command>'
T=$1; shift
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --timeout $T -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient\|status=busy\|retry in a few minutes"; then sleep 120; continue; fi
  echo "$out" | tail -60
  exit 0
done
echo "gave up"
