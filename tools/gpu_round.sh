#!/bin/bash
# One measurement pass on the GPU box, everything a round needs in a single gpurun call (about 3 GPU-minutes):
#   gpurun --timeout 420 -- 'bash tools/gpu_round.sh r02'
# writes into gpurun_out/: <tag>_tests.log, <tag>_bench.json, <tag>_widening_times.json, <tag>_launches.csv,
# <tag>_top.ncu-rep (ncu --set full of the kernels named in $KERNELS).  Summarise afterwards, where ncu is
# installed, with tools/ncu_extract.py and copy what is to be judged into profiles/.
TAG=${1:-rXX}
KERNELS=${KERNELS:-"k_ddm_codes_simd|k_ccl_strip|k_boost_inside4|k_tta_merge4"}
OUT=gpurun_out
mkdir -p $OUT
timeout 120 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_tests.log 2>&1; echo "tests rc=$?" | tee -a $OUT/${TAG}_tests.log
timeout 240 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
timeout 60 python tools/time_widening.py > $OUT/${TAG}_widening_times.json 2> $OUT/${TAG}_widening.err; echo "widening rc=$?"
# launch list of the bench command: shares per kernel (cold-cache, serialised: compare shares, not absolutes)
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -s 120 -c 80 --csv --log-file $OUT/${TAG}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --skip-extra all \
    > $OUT/${TAG}_launches.log 2>&1; echo "launch list rc=$?"
# one full capture per dominant kernel (post-processing step, then the hand-off kernel)
timeout 200 ncu --set full --clock-control none --import-source on -k "regex:$KERNELS" -s 8 -c 4 -f -o $OUT/${TAG}_top \
    python bench.py --steps 2 --warmup 2 --no-cpu-baseline --skip-extra all > $OUT/${TAG}_top.log 2>&1; echo "ncu full rc=$?"
timeout 120 ncu --set full --clock-control none --import-source on -k "regex:k_tta_merge" -s 1 -c 1 -f -o $OUT/${TAG}_tta \
    python tools/ncu_tta.py > $OUT/${TAG}_tta.log 2>&1; echo "ncu tta rc=$?"
tail -3 $OUT/${TAG}_tests.log; head -c 600 $OUT/${TAG}_bench.json
# target-transform kernels: full captures with stall reasons (one launch each)
timeout 200 ncu --set full --clock-control none --import-source on \
    -k "regex:k_t_direction_lab|k_t_centerness|k_flood|k_t_gauss|k_edt_cols|k_edt_rows" -s 12 -c 8 -f -o $OUT/${TAG}_tpath \
    python tools/bench_extra.py --what targets --steps 1 > $OUT/${TAG}_tpath.log 2>&1; echo "ncu tpath rc=$?"
