#!/bin/bash
# bench.py --device-only under a list of environment settings: bash tools/gpu_ab.sh "A=1" "B=2 C=3" ...
OUT=gpurun_out
mkdir -p $OUT
i=0
for V in "" "$@" ""; do
  i=$((i+1))
  env $V timeout 120 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --skip-extra all --device-only > $OUT/ab_$i.json 2>$OUT/ab.err || tail -3 $OUT/ab.err
  python - <<P
import json
d=json.load(open("$OUT/ab_$i.json"))
k=d["roofline"]["kernels_ms_per_step"]
print("[$V]", round(d["ms_per_step"],4), " ".join("%s=%.1f"%(a.replace("k_rle_","").replace("k_",""),1e3*b) for a,b in sorted(k.items(), key=lambda kv:-kv[1])))
P
done
