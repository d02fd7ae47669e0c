"""Profiling target for the per-kernel evidence table: after a warm-up, ONE eager training step (batch 256), one MLP
step, one augmentation launch, one inference pass (batch 4096) between cudaProfilerStart/Stop (ncu --profile-from-start off)."""
import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ae_b200
dev = torch.device('cuda', 0)
prec = sys.argv[1] if len(sys.argv) > 1 else 'fp32'
torch.manual_seed(0)
m = ae_b200.SupervisedAutoencoder(64, 10, precision=prec).to(dev).train()
clf = ae_b200.MLP(64, 10).to(dev)
B = 256
x = torch.rand(B, 3, 64, 64, device=dev); y = torch.randint(0, 10, (B,), device=dev)
u8 = torch.randint(0, 256, (B, 64, 64, 3), dtype=torch.uint8, device=dev)
xi = torch.rand(4096, 3, 64, 64, device=dev)
z = torch.randn(B, 64, device=dev)
m.engine().prepare(dev, 4096)
opt = ae_b200.Adam(m.parameters(), lr=5e-3)
clf._state.prepare(dev, B)
copt = ae_b200.Adam(clf.parameters(), lr=1e-3, weight_decay=1e-4)
tf = ae_b200.TrainTransformAE(seed=1)

def one():
    m.train()
    opt.zero_grad(); loss = m.train_step_grads(x, y, 35.0); opt.step(); m.engine().prepare(dev, B)
    clf.train(); copt.zero_grad(); clf.fused_step_grads(z, y); copt.step()
    tf(u8)
    m.eval(); clf.eval()
    ae_b200.encode_predict(m.enc, clf, xi)
    return loss

for _ in range(2):
    one()
torch.cuda.synchronize()
torch.cuda.profiler.start()
loss = one()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print('ok', loss.tolist())
