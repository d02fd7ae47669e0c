#!/bin/bash
# compute-sanitizer over one small run of every kernel family.  ONE tool per gpurun call (B200_PROFILING.md):
#   scripts/gpu_sanitize.sh memcheck   |   scripts/gpu_sanitize.sh racecheck
set -u
TOOL=${1:-memcheck}
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python scripts/sanitize_target.py > $OUT/sanitize_plain.log 2>&1 || { echo "plain run failed"; tail -20 $OUT/sanitize_plain.log; exit 1; }
timeout 1500 compute-sanitizer --tool $TOOL --print-limit 50 --launch-timeout 600 python scripts/sanitize_target.py > $OUT/sanitize_$TOOL.log 2>&1
echo "compute-sanitizer $TOOL rc=$?"
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Error|Hazard|hazard" $OUT/sanitize_$TOOL.log | sort | uniq -c | head -30
tail -4 $OUT/sanitize_$TOOL.log
