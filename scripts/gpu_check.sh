#!/bin/bash
# Run on the GPU box (through gpurun): GPU parity tests, smoke, bench, ncu launch list.
# usage: scripts/gpu_check.sh <tag> [pytest-args...]
set -u
TAG=${1:-run}; shift || true
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q "$@" > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest.log
timeout 300 python __graft_entry__.py smoke > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/${TAG}_smoke.log
timeout 600 python bench.py --steps 100 --warmup 10 > $OUT/${TAG}_bench_fp32.json 2> $OUT/${TAG}_bench_fp32.err; echo "bench fp32 rc=$?"
timeout 600 python bench.py --steps 100 --warmup 10 --precision bf16 --no-cpu-baseline > $OUT/${TAG}_bench_bf16.json 2> $OUT/${TAG}_bench_bf16.err; echo "bench bf16 rc=$?"
tail -3 $OUT/${TAG}_pytest.log; cat $OUT/${TAG}_smoke.log | tail -2; cat $OUT/${TAG}_bench_fp32.json $OUT/${TAG}_bench_bf16.json
