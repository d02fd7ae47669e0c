#!/bin/bash
# round-2 check: full GPU parity suite, smoke, bench (all extras), gen1-vs-gen2 row GEMM at two batch sizes
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-s2m}
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/${TAG}_pytest.log
timeout 300 python __graft_entry__.py smoke > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/${TAG}_smoke.log
timeout 900 python bench.py --steps 100 --warmup 10 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"; tail -3 $OUT/${TAG}_bench.err
python - <<PY
import json
d=json.load(open("$OUT/${TAG}_bench.json"))
print({k:d[k] for k in ("value","ms_per_step","kernels_per_step")}, "e2e", d["e2e"]["value"], "roof", d["roofline"]["avg_launch_us"], d["roofline"]["frac"], "infer", d["inference"]["value"])
for k in ("torch_gpu_baseline","dropin_loop","mlp_train","dp_global_4096","cpu_baseline"):
    print(k, d.get(k))
PY
timeout 600 python scripts/rowgemm_bench.py 256 > $OUT/${TAG}_rg256.log 2>&1; tail -14 $OUT/${TAG}_rg256.log
timeout 600 python scripts/rowgemm_bench.py 2048 > $OUT/${TAG}_rg2048.log 2>&1; tail -14 $OUT/${TAG}_rg2048.log
