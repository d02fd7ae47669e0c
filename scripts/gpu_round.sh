#!/bin/bash
# One GPU call: parity tests, smoke, bench (fp32 + bf16), warm launch list of the step.
# usage: scripts/gpu_round.sh <tag> [pytest -k expression]
set -u
TAG=${1:-run}; KEXPR=${2:-}
OUT=gpurun_out
mkdir -p $OUT
if [ -n "$KEXPR" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q -k "$KEXPR" > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest.log
else
  timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest.log
fi
tail -15 $OUT/${TAG}_pytest.log
timeout 300 python __graft_entry__.py smoke > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py --steps 100 --warmup 10 > $OUT/${TAG}_bench_fp32.json 2> $OUT/${TAG}_bench_fp32.err; echo "bench fp32 rc=$?"
cat $OUT/${TAG}_bench_fp32.json; tail -3 $OUT/${TAG}_bench_fp32.err
timeout 600 python bench.py --steps 100 --warmup 10 --precision bf16 --no-cpu-baseline > $OUT/${TAG}_bench_bf16.json 2> $OUT/${TAG}_bench_bf16.err; echo "bench bf16 rc=$?"
cat $OUT/${TAG}_bench_bf16.json
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-roofline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/${TAG}_launches.csv $CMD > $OUT/${TAG}_ncu_list.log 2>&1
echo "ncu list rc=$?"
