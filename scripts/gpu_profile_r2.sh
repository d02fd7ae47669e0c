#!/bin/bash
# round-2 evidence: warm launch list of the step + `ncu --set full` of the roofline kernel (all ncu runs of ONE call)
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r2p}
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-roofline --no-extras"
$CMD > $OUT/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -20 $OUT/${TAG}_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 3000 --csv --log-file $OUT/${TAG}_launches.csv $CMD > $OUT/${TAG}_ncu_list.log 2>&1
echo "ncu list rc=$?"
python scripts/ncu_summary.py launches $OUT/${TAG}_launches.csv > $OUT/${TAG}_launches.txt 2>&1; head -40 $OUT/${TAG}_launches.txt
CMD2="python bench.py --roofline-only"
$CMD2 > $OUT/${TAG}_roof_plain.log 2>&1 || { echo "roofline plain run failed"; tail -20 $OUT/${TAG}_roof_plain.log; exit 1; }
cat $OUT/${TAG}_roof_plain.log | tail -1
ncu --set full --clock-control none --import-source on -k regex:k_rowgemm2 -s 2 -c 3 -o $OUT/${TAG}_roof $CMD2 > $OUT/${TAG}_ncu_full.log 2>&1
echo "ncu full rc=$?"; ls -la $OUT/${TAG}_roof.ncu-rep
python scripts/ncu_summary.py full $OUT/${TAG}_roof.ncu-rep > $OUT/${TAG}_roof_full.txt 2>&1; cat $OUT/${TAG}_roof_full.txt
