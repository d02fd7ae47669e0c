#!/bin/bash
# quick check of a row-GEMM change: op + model parity, gen1-vs-gen2 kernel table, one short bench
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-q}
timeout 1200 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/${TAG}_pytest.log
timeout 600 python scripts/rowgemm_bench.py 256 > $OUT/${TAG}_rg256.log 2>&1; tail -14 $OUT/${TAG}_rg256.log
timeout 600 python bench.py --steps 100 --warmup 10 --no-extras > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("$OUT/${TAG}_bench.json"))
print({k:d[k] for k in ("value","ms_per_step","kernels_per_step")}, "roof", d["roofline"]["avg_launch_us"], d["roofline"]["frac"])
PY
