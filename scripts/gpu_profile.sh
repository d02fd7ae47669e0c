#!/bin/bash
# Run on the GPU box (through gpurun): ncu launch list of a short bench run + one full capture.
# usage: scripts/gpu_profile.sh <tag> [kernel-regex] [extra bench args]
set -u
TAG=${1:-prof}; KRE=${2:-}; shift; shift || true
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline $*"
$CMD > $OUT/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -20 $OUT/${TAG}_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $OUT/${TAG}_launches.csv $CMD > $OUT/${TAG}_ncu_list.log 2>&1
echo "ncu list rc=$?"
if [ -n "$KRE" ]; then
  ncu --set full --clock-control none --import-source on -k "regex:$KRE" -s ${NCU_SKIP:-63} -c ${NCU_COUNT:-21} -o $OUT/${TAG}_full $CMD > $OUT/${TAG}_ncu_full.log 2>&1
  echo "ncu full rc=$?"
fi
ls -la $OUT | tail -20
