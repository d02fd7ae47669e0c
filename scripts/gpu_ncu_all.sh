#!/bin/bash
# usage: scripts/gpu_ncu_all.sh <tag> [precision]
# ncu --set full of every kernel of one training step, one MLP step, one augmentation launch and one inference pass.
set -u
TAG=$1; PREC=${2:-fp32}
OUT=gpurun_out; mkdir -p $OUT
CMD="python scripts/prof_all.py $PREC"
$CMD > $OUT/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -20 $OUT/${TAG}_plain.log; exit 1; }
timeout 800 ncu --set full --clock-control none --profile-from-start off -o $OUT/${TAG}_all $CMD > $OUT/${TAG}_ncu_all.log 2>&1
echo "ncu all rc=$?"; tail -2 $OUT/${TAG}_ncu_all.log
ncu -i $OUT/${TAG}_all.ncu-rep --page raw --csv > $OUT/${TAG}_all_raw.csv 2>/dev/null
du -sh $OUT/*
tot=$(du -sm $OUT | cut -f1)
if [ "$tot" -gt 58 ]; then rm -f $OUT/${TAG}_all.ncu-rep; echo "report dropped (too large), csv kept"; fi
