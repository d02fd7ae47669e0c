#!/bin/bash
# usage: scripts/gpu_ncu_full.sh <tag> <kernel-regex> <skip> <count> [bench args]
set -u
TAG=$1; KRE=$2; SKIP=$3; COUNT=$4; shift 4
OUT=gpurun_out; mkdir -p $OUT
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-roofline $*"
$CMD > $OUT/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -20 $OUT/${TAG}_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k "regex:$KRE" -s $SKIP -c $COUNT -o $OUT/${TAG}_full $CMD > $OUT/${TAG}_ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -3 $OUT/${TAG}_ncu_full.log
