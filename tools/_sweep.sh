#!/bin/bash
for lib in libdfq_sm100 lib_novec lib_nocm lib_both; do
  echo "== $lib"
  DFQ_LIB=/root/repo/dfq_b200/$lib.so DFQ_TRACE=1 timeout -k 5 200 python bench.py --steps 3 --layers ${1:-2048} --no-cpu-baseline --no-e2e --no-mbv2 2>&1 | grep -E "phase ms" | tail -1 | cut -c1-300
done
