"""Profiling target: 2 eager training steps at batch 256 (same kernels as the captured step) + 2 inference passes at 4096."""
import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ae_b200
dev = torch.device('cuda', 0)
prec = sys.argv[1] if len(sys.argv) > 1 else 'fp32'
torch.manual_seed(0)
m = ae_b200.SupervisedAutoencoder(64, 10, precision=prec).to(dev).train()
clf = ae_b200.MLP(64, 10).to(dev).eval()
B = 256
x = torch.rand(B, 3, 64, 64, device=dev); y = torch.randint(0, 10, (B,), device=dev)
m.engine().prepare(dev, B)
opt = ae_b200.Adam(m.parameters(), lr=5e-3)
for i in range(2):
    opt.zero_grad(); loss = m.train_step_grads(x, y, 35.0); opt.step(); m.engine().prepare(dev, B)
torch.cuda.synchronize()
m.eval()
xi = torch.rand(4096, 3, 64, 64, device=dev)
for i in range(2):
    z, lg, am = ae_b200.encode_predict(m.enc, clf, xi)
torch.cuda.synchronize()
print('ok', loss.tolist(), int(am.sum()))
