#!/bin/bash
# tools/gpu.sh with retries while the pod answers "transient" (no slot / draining; nothing is charged for those).
# usage: tools/gpu_retry.sh <timeout_s> <logname> '<command>' [max_tries]
cd "$(dirname "$0")/.."
for i in $(seq 1 ${4:-20}); do
  tools/gpu.sh "$1" "$2" "$3" > /dev/null 2>&1
  if ! grep -q "status=transient\|status=busy" "gpurun_out/$2.log"; then break; fi
  sleep 90
done
tail -5 "gpurun_out/$2.log"
