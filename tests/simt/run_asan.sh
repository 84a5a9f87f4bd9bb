#!/bin/bash
# Runs the [simt] kernel tests against an AddressSanitizer build of the emulated library (test infrastructure):
# out-of-bounds accesses of any kernel on its inputs, outputs or workspace slices abort with a report.
#   tests/simt/run_asan.sh [pytest args, default: all tests/test_gpu_*.py]
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REPO="$(dirname "$(dirname "$HERE")")"
cd "$REPO"
LIB="$(python tests/simt/build.py --asan | tail -1)"
ASAN_RT="$(/usr/bin/gcc -print-file-name=libasan.so)"
[ -f "$LIB" ] && [ -f "$ASAN_RT" ] || { echo "no AddressSanitizer build / runtime"; exit 1; }
[ $# -eq 0 ] && set -- tests/test_gpu_postproc.py tests/test_gpu_targets.py tests/test_gpu_training.py tests/test_gpu_metrics.py tests/test_gpu_sharded.py tests/test_gpu_widening.py
rm -f tests/simt/_build/asan/report.*
CDNET_SIMT_LIB="$LIB" LD_PRELOAD="$ASAN_RT" \
ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0:halt_on_error=1:log_path=tests/simt/_build/asan/report \
    python -m pytest "$@" -q -m "not gpu" -x 2>&1 | grep -v "WARNING: ASan doesn't fully support" | tail -8
for r in tests/simt/_build/asan/report.*; do
    [ -f "$r" ] && grep -v WARNING "$r" | grep -E "ERROR|^    #[0-9]+ .*(cdnet|simt)|is located" | head -20
done
exit 0
