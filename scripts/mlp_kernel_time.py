#!/usr/bin/env python
"""Device time of one captured MLP training step (k_mlp or k_mlp_small + k_adam), replayed back to back."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ae_b200
dev = torch.device("cuda", 0)
for batch in ([int(a) for a in sys.argv[1:]] or [64]):
    clf = ae_b200.MLP(64, 10).to(dev).train()
    clf._state.prepare(dev, batch)
    opt = ae_b200.Adam(clf.parameters(), lr=1e-4, weight_decay=1e-4)
    step = ae_b200.MLPTrainStep(clf, opt, batch)
    step.x.normal_(); step.y.random_(0, 10)
    for _ in range(20): step.run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(500): step.run()
    e1.record(); torch.cuda.synchronize()
    print(f"batch {batch}: {e0.elapsed_time(e1) / 500 * 1e3:.1f} us per replayed step (AE_B200_MLP_CLUSTER={os.environ.get('AE_B200_MLP_CLUSTER', 'auto')})")

# eval: k_mlp_eval on 4096 latents (part of the inference pass)
clf = ae_b200.MLP(64, 10).to(dev).eval()
X = torch.randn(4096, 64, device=dev)
with torch.no_grad():
    for _ in range(10): clf.predict(X)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200): clf.predict(X)
    e1.record(); torch.cuda.synchronize()
print(f"MLP eval, 4096 rows: {e0.elapsed_time(e1) / 200 * 1e3:.1f} us per call (host loop included)")
