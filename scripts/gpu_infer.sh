#!/bin/bash
# eval-path change: model parity tests, then the inference leg of the bench with and without the tensor-core dense layer
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-inf}
timeout 1200 python -m pytest tests/test_gpu_model.py tests/test_gpu_data.py -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/${TAG}_pytest.log
for v in 1 0; do
AE_B200_DENSE_TC=$v timeout 600 python - <<PY
import json, sys, torch
sys.argv=["bench.py"]
import bench
dev=torch.device("cuda",0)
r=bench.inference_rate(dev,"fp32","tc")
print("DENSE_TC=$v", json.dumps(r))
PY
done
