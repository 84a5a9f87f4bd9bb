#!/bin/bash
# compute-sanitizer pass over the small-tile GPU parity tests (racecheck / memcheck / initcheck on
# the kernels).  Run on the GPU box:  gpurun --timeout 900 -- 'bash tools/gpu_sanitize.sh r02'
# Writes gpurun_out/<tag>_sanitizer_<tool>.log; summarise into profiles/<tag>_sanitizer.md.
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
SEL=${SEL:-"96x128 or 200x150 or thin_shards or process_golden or t_64_single or t_128 or primitives or circshift or ddm_golden or voting or metrics or sharded_equals_oracle"}
SEL="($SEL) and not p_1000 and not 333x517"
for TOOL in ${TOOLS:-memcheck racecheck initcheck}; do
    EXTRA=""
    [ "$TOOL" = "racecheck" ] && EXTRA="--racecheck-report all"
    [ "$TOOL" = "initcheck" ] && EXTRA=""
    timeout ${TMO:-420} compute-sanitizer --tool $TOOL $EXTRA --print-limit 40 --error-exitcode 0 \
        --log-file $OUT/${TAG}_sanitizer_${TOOL}.log \
        python -m pytest tests -m gpu -q -x -k "$SEL" -p no:cacheprovider > $OUT/${TAG}_sanitizer_${TOOL}_pytest.log 2>&1
    echo "$TOOL rc=$? : $(tail -1 $OUT/${TAG}_sanitizer_${TOOL}_pytest.log)"
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY" $OUT/${TAG}_sanitizer_${TOOL}.log | tail -2
done
