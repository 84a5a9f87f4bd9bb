#!/bin/bash
# the watershed tile path (postproc = 1, 14 x 1000^2): GPU tests, one timed call, ncu launch list with per-kernel shares
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python tools/ws_prof.py
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/ws_launches.csv python tools/ws_prof.py > /dev/null 2>&1
python - <<EOF
import csv, collections
rows=[r for r in csv.reader(open("$OUT/ws_launches.csv")) if len(r)>10 and r[0].isdigit()]
n=len(rows)//4
agg=collections.OrderedDict()
for r in rows[-n:]:
    k=r[4].split("(")[0][:60]
    agg[k]=agg.get(k,0)+float(r[-1].replace(",",""))
tot=sum(agg.values())
for k,v in sorted(agg.items(), key=lambda kv:-kv[1])[:22]: print("%-60s %10.1f us %5.1f%%"%(k,v/1e3,100*v/tot))
print("total us", tot/1e3, "launches", n)
EOF
