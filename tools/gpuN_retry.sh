#!/bin/bash
# gpurun --gpus N with retries on transient answers.  usage: tools/gpuN_retry.sh <N> <timeout_s> <logname> '<command>' [tries]
cd "$(dirname "$0")/.."
python -m dfq_b200._build >/dev/null 2>&1
mkdir -p gpurun_out
for i in $(seq 1 ${5:-15}); do
  /usr/local/graft/bin/gpurun --gpus "$1" --timeout "$2" -- "$4" > "gpurun_out/$3.log" 2>&1
  if ! grep -q "status=transient\|status=busy" "gpurun_out/$3.log"; then break; fi
  sleep 120
done
tail -6 "gpurun_out/$3.log"
