#!/bin/bash
# MobileNetV2 e2e inside bench.py's process environment (NUMA binding, clock sampler) against the host-copy thread count
# and the number of upload parts
for cfg in "1 1" "1 4" "0 1" "0 4" "0 2"; do
  set -- $cfg
  echo "== DFQ_HOST_COPY_THREADS=$1 DFQ_UPLOAD_PARTS=$2"
  DFQ_HOST_COPY_THREADS=$1 DFQ_UPLOAD_PARTS=$2 timeout -k 5 200 python bench.py --layers 256 --no-e2e --no-cpu-baseline --steps 3 --warmup 3 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); m=d['mobilenetv2']; print(m['e2e_ms'], m['device_ms'], m['e2e_parts_ms'])"
done
