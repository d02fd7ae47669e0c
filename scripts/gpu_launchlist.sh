#!/bin/bash
# usage: scripts/gpu_launchlist.sh <tag> [bench args]   -- warm-cache (--cache-control none) per-launch device times
set -u
TAG=$1; shift
OUT=gpurun_out; mkdir -p $OUT
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-roofline $*"
$CMD > $OUT/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -20 $OUT/${TAG}_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 3000 --csv --log-file $OUT/${TAG}_launches.csv $CMD > $OUT/${TAG}_ncu_list.log 2>&1
echo "ncu list rc=$?"
python scripts/ncu_summary.py launches $OUT/${TAG}_launches.csv > $OUT/${TAG}_launches.txt 2>&1; head -45 $OUT/${TAG}_launches.txt
