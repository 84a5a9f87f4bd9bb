#!/bin/bash
for cfg in "4 1024" "4 512" "2 1024" "2 512" "6 1024" "8 1024"; do
  set -- $cfg
  CDNET_STRIP_ROWS=$1 CDNET_STRIP_THREADS=$2 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); k=d['roofline']['kernels_ms_per_step']; print('rows=$1 threads=$2', 'strip', k.get('k_ccl_strip'), 'merge4', k.get('k_ccl_merge4'), 'flatten_fill4', k.get('k_flatten_fill4'), 'step_ms', round(d['ms_per_step'],4))"
done
