#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=$1; N=$2
for R in 1 2; do
AE_B200_DP_ROUNDS=$R timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 200 --warmup 20 --no-cpu-baseline --no-roofline --no-e2e --no-extras > $OUT/${TAG}_r$R.json 2> $OUT/${TAG}_r$R.err
python - <<PY
import json
d=json.loads(open("$OUT/${TAG}_r$R.json").read().strip().splitlines()[-1])
print("rounds $R", {k:d[k] for k in ("value","ms_per_step","kernels_per_step","n_gpus")})
PY
done
