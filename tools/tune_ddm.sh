#!/bin/bash
# rebuild the library with a few DDM kernel shapes ON THE GPU BOX and time the kernel (tuning aid)
for cfg in "4 2 5" "8 1 4" "8 1 5" "8 2 4" "4 2 6" "4 1 6"; do
  set -- $cfg
  python -m cdnet_b200.build --force -DCDNET_DDM_ROWS_DEFAULT=$1 -DCDNET_DDM_PB=$2 -DCDNET_DDM_MINB=$3 > /dev/null 2>&1
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('rows=$1 pb=$2 minb=$3', 'ddm_ms', d['roofline']['kernels_ms_per_step'].get('k_ddm_codes_simd'), 'step_ms', round(d['ms_per_step'],4))"
done
python -m cdnet_b200.build --force > /dev/null 2>&1
