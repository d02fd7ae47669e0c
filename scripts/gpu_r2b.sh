#!/bin/bash
# round-2: parity of the second-generation row GEMM + its timeline + the step
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-s2d}
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -x -q > $OUT/${TAG}_ops.log 2>&1; echo "ops pytest rc=$?"; tail -4 $OUT/${TAG}_ops.log
timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -x -q > $OUT/${TAG}_model.log 2>&1; echo "model pytest rc=$?"; tail -4 $OUT/${TAG}_model.log
AE_B200_LIB=$PWD/hybrid-autoencoder-mlp-pipeline-for-satellite-image-classification_b200/libae_b200_trace.so timeout 300 python scripts/trace_rowgemm.py 256 all > $OUT/${TAG}_trace.log 2>&1; echo "trace rc=$?"
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("$OUT/${TAG}_bench.json"))
print({k:d[k] for k in ("value","ms_per_step","kernels_per_step")}, d["e2e"]["value"], d["roofline"]["avg_launch_us"], d["roofline"]["frac"], d["inference"]["value"])
PY
