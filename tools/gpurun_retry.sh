#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout-seconds> '<command>' [gpus]   -- retries while the pod answers "busy" (nothing is charged then)
T=$1; CMD=$2; G=${3:-1}
for i in $(seq 1 30); do
  if [ "$G" = "1" ]; then OUT=$(/usr/local/graft/bin/gpurun --timeout "$T" -- "$CMD" 2>&1); else OUT=$(/usr/local/graft/bin/gpurun --gpus "$G" --timeout "$T" -- "$CMD" 2>&1); fi
  echo "$OUT" | tail -60
  if echo "$OUT" | grep -q "status=transient"; then echo "[retry $i] busy, sleeping 90 s"; sleep 90; else exit 0; fi
done
