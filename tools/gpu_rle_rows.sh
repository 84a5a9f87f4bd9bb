#!/bin/bash
# A/B of the rows per block of k_rle_pack_link (shared-memory union-find) on the GPU box
OUT=gpurun_out
mkdir -p $OUT
for R in 8 16 32 8 16 32; do
  CDNET_RLE_PACK_ROWS=$R timeout 120 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --skip-extra all --device-only > $OUT/rle_rows_$R.json 2>$OUT/rle_rows.err
  python - <<P
import json
d=json.load(open("$OUT/rle_rows_$R.json"))
print("pack rows $R", round(d["ms_per_step"],4), {k:round(v,4) for k,v in d["roofline"]["kernels_ms_per_step"].items() if "pack" in k or "rle_link" in k or "holes" in k or "touch" in k})
P
done
