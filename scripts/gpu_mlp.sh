#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-s3c}
timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_data.py -m gpu -x -q -k "mlp or pipeline or fit" 2>&1 | tail -4
for c in 1 2 4 8; do
AE_B200_MLP_CLUSTER=$c timeout 120 python - <<PY
import torch, sys, bench
print("cluster $c", bench.mlp_train_rate(torch.device("cuda", 0), steps=1280))
PY
done
timeout 120 python - <<PY
import torch, bench
print("auto", bench.mlp_train_rate(torch.device("cuda", 0), steps=1280))
print("auto b256", bench.mlp_train_rate(torch.device("cuda", 0), steps=1280, batch=256))
PY
timeout 300 python scripts/full_pipeline.py --precision bf16 2>&1 | tail -2
