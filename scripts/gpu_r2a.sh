#!/bin/bash
# round-2 call A: does the fused tail (grid barrier + BatchNorm job + operand conversion inside the row GEMM) work on hardware?
set -u
OUT=gpurun_out; mkdir -p $OUT
AE_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -x -q -k fused_tail > $OUT/s2a_tail_pytest.log 2>&1; echo "tail pytest rc=$?"
tail -5 $OUT/s2a_tail_pytest.log
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $OUT/s2a_bench_default.json 2> $OUT/s2a_bench_default.err; echo "bench default rc=$?"
cat $OUT/s2a_bench_default.json
AE_B200_FUSED_TAIL=1 timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $OUT/s2a_bench_tail.json 2> $OUT/s2a_bench_tail.err; echo "bench tail rc=$?"
cat $OUT/s2a_bench_tail.json; tail -3 $OUT/s2a_bench_tail.err
