#!/bin/bash
# per-kernel device times of the inference pass (ncu launch list, warm caches) -> profiles/
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-inflist}; B=${2:-4096}
python scripts/infer_kernels.py $B fp32 > $OUT/${TAG}_plain.log 2>&1 || { tail -20 $OUT/${TAG}_plain.log; exit 1; }
python scripts/infer_kernels.py $B bf16 >> $OUT/${TAG}_plain.log 2>&1
cat $OUT/${TAG}_plain.log
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none --cache-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv python scripts/infer_kernels.py $B fp32 > $OUT/${TAG}_ncu.log 2>&1
echo "ncu rc=$?"
