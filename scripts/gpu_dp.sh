#!/bin/bash
# data-parallel bench on N GPUs of one box: usage scripts/gpu_dp.sh <tag> <N> [steps]
set -u
TAG=$1; N=$2; STEPS=${3:-100}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps $STEPS --warmup 10 --no-cpu-baseline > $OUT/${TAG}_dp$N.json 2> $OUT/${TAG}_dp$N.err
echo "dp$N rc=$?"; tail -5 $OUT/${TAG}_dp$N.err
python - <<PY
import json
d=json.loads(open("$OUT/${TAG}_dp$N.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","kernels_per_step","n_gpus")}, "e2e", d["e2e"]["value"])
for k in ("dp_global_4096","dp_check"):
    print(k, d.get(k))
PY
