#!/usr/bin/env python
"""Encoder + MLP inference (clf(enc(x)).argmax, eval mode) at one batch size, a few passes: the target of the per-kernel ncu
launch list under profiles/ (BASELINE configs[4]).   python scripts/infer_kernels.py [batch] [precision]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import ae_b200

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
prec = sys.argv[2] if len(sys.argv) > 2 else "fp32"
dev = torch.device("cuda", 0)
torch.manual_seed(1)
ae = ae_b200.SupervisedAutoencoder(64, 10, precision=prec, backend="tc").to(dev).eval()
clf = ae_b200.MLP(64, 10).to(dev).eval()
xs = [torch.rand(batch, 3, 64, 64, device=dev) for _ in range(3)]
for i in range(6):
    pred = ae_b200.encode_predict(ae.enc, clf, xs[i % 3])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(10):
    pred = ae_b200.encode_predict(ae.enc, clf, xs[i % 3])
e1.record()
torch.cuda.synchronize()
print(f"batch {batch} {prec}: {e0.elapsed_time(e1) / 10 * 1e3:.1f} us per pass, {batch * 10 / e0.elapsed_time(e1) / 1e3:.3f} M images/s")
