#!/bin/bash
# Build the CUDA library here (nvcc cross-compiles), then run a command on the B200 box.
# usage: tools/gpu.sh <timeout_s> <logname> '<command>'   (inside <command> use `timeout -k 5 N ...`: a Python
# process blocked in a CUDA call ignores SIGTERM)
set -e
cd "$(dirname "$0")/.."
python -m dfq_b200._build >/dev/null 2>&1 || { python -m dfq_b200._build 2>&1 | grep -E "error" ; exit 1; }
mkdir -p gpurun_out
/usr/local/graft/bin/gpurun --timeout "$1" -- "$3" > "gpurun_out/$2.log" 2>&1
tail -5 "gpurun_out/$2.log"
