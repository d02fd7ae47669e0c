#!/bin/bash
# ncu --set full of the thin kernels at one batch size (one launch each after two skipped)
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-thin}; B=${2:-4096}; export THIN_ONLY=${3:-}; export THIN_ITERS=1
python scripts/thin_bench.py $B > $OUT/${TAG}_plain.log 2>&1 || { tail -20 $OUT/${TAG}_plain.log; exit 1; }
cat $OUT/${TAG}_plain.log
ncu --set full --clock-control none --import-source on -k regex:"^k_thin$" -c ${4:-12} -o $OUT/${TAG}_thin python scripts/thin_bench.py $B > $OUT/${TAG}_ncu.log 2>&1
echo "ncu rc=$?"; ls -la $OUT/${TAG}_thin.ncu-rep
