#!/bin/bash
# usage: scripts/gpu_ncu_mix.sh <tag> [precision]
# ncu --set full of the thin kernels (2 training steps at batch 256 + 2 inference passes at 4096) and of the inference row GEMMs;
# text summaries are produced on the box, the reports are kept only while they fit the 64 MiB return limit.
set -u
TAG=$1; PREC=${2:-fp32}
OUT=gpurun_out; mkdir -p $OUT
CMD="python scripts/prof_mix.py $PREC"
$CMD > $OUT/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -20 $OUT/${TAG}_plain.log; exit 1; }
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_thin" -c 14 -o $OUT/${TAG}_thin $CMD > $OUT/${TAG}_ncu_thin.log 2>&1
echo "ncu thin rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_tma_rowgemm" -s 24 -c 6 -o $OUT/${TAG}_gemm $CMD > $OUT/${TAG}_ncu_gemm.log 2>&1
echo "ncu gemm rc=$?"
for r in thin gemm; do
  ncu -i $OUT/${TAG}_$r.ncu-rep --page raw --csv > $OUT/${TAG}_${r}_raw.csv 2>/dev/null
  ncu -i $OUT/${TAG}_$r.ncu-rep --page details > $OUT/${TAG}_${r}_details.txt 2>/dev/null
done
du -sh $OUT/*
tot=$(du -sm $OUT | cut -f1)
if [ "$tot" -gt 55 ]; then rm -f $OUT/${TAG}_gemm.ncu-rep; fi
tot=$(du -sm $OUT | cut -f1)
if [ "$tot" -gt 55 ]; then rm -f $OUT/${TAG}_thin.ncu-rep; fi
