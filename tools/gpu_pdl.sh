#!/bin/bash
# A/B of programmatic dependent launch on the chain of short kernels (boost -> run-based tail), graph and eager
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python -m pytest tests -m gpu -x -q > $OUT/pdl_tests.log 2>&1; echo "tests rc=$?"; tail -2 $OUT/pdl_tests.log
for V in pdl nopdl pdl nopdl; do
  if [ $V = nopdl ]; then export CDNET_NO_PDL=1; else unset CDNET_NO_PDL; fi
  timeout 120 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --skip-extra all --device-only > $OUT/pdl_$V.json 2>$OUT/pdl.err || tail -3 $OUT/pdl.err
  python - <<P
import json
d=json.load(open("$OUT/pdl_$V.json"))
print("$V", round(d["ms_per_step"],4), d["value"])
P
done
