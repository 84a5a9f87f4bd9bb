#!/bin/bash
# A/B of the direction-difference kernels on the GPU box: parity tests, bench with the bit-sliced kernel at several
# strip heights / occupancy targets and with the byte-SIMD kernel, one full ncu capture of a 14-tile k_ddm_bits launch.
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
timeout 200 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -2 $OUT/${TAG}_tests.log
for MB in 3 2; do for R in 16 8 32; do
  CDNET_DDM_BITS_MB=$MB CDNET_DDM_BITS_ROWS=$R timeout 120 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --skip-extra all > $OUT/${TAG}_bench_mb${MB}_r$R.json 2>$OUT/${TAG}_bench.err
  python - <<P
import json
d=json.load(open("$OUT/${TAG}_bench_mb${MB}_r$R.json"))
print("mb $MB rows $R", round(d["ms_per_step"],4), {k:v for k,v in d["roofline"]["kernels_ms_per_step"].items() if "ddm" in k})
P
done; done
timeout 200 ncu --set full --clock-control none --import-source on -k "regex:k_ddm_bits" -s 9 -c 1 -f -o $OUT/${TAG}_ddm_bits \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --skip-extra all > $OUT/${TAG}_ddm_bits_ncu.log 2>&1; echo "ncu rc=$?"
