#!/bin/bash
# ncu evidence for the post-processing step (14-tile launches only): launch list with DRAM bytes, then one full
# capture of every kernel of one step.   gpurun --timeout 900 -- 'bash tools/gpu_profile.sh r02'
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --skip-extra all --device-only"
# one step = L launches (12 since the frame-flag pass went away); skip the population launch + 3 warm-ups, list two steps
L=${2:-12}
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -s $((4 * L)) -c $((2 * L)) --csv --log-file $OUT/${TAG}_launches.csv $B > $OUT/${TAG}_launches.log 2>&1; echo "launch list rc=$?"
timeout 500 ncu --set full --clock-control none --import-source on -s $((4 * L)) -c $L -f -o $OUT/${TAG}_step $B > $OUT/${TAG}_step.log 2>&1
echo "ncu full rc=$?"
