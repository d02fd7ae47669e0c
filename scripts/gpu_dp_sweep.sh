#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=$1; N=$2
for F in 0 1; do
AE_B200_DP_FUSED=$F timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --warmup 10 --no-cpu-baseline --no-roofline --no-e2e --no-extras > $OUT/${TAG}_f$F.json 2> $OUT/${TAG}_f$F.err
echo "rc=$?"; tail -4 $OUT/${TAG}_f$F.err
python - <<PY
import json
d=json.loads(open("$OUT/${TAG}_f$F.json").read().strip().splitlines()[-1])
print("fused $F", {k:d[k] for k in ("value","ms_per_step","kernels_per_step","n_gpus")}, "e2e", d["e2e"]["value"])
print(d.get("dp_global_4096")); print(d.get("dp_check"))
PY
done
