import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ae_b200
dev = torch.device('cuda', 0)
prec = sys.argv[1] if len(sys.argv) > 1 else 'fp32'
backend = sys.argv[2] if len(sys.argv) > 2 else 'tc'
B = int(sys.argv[3]) if len(sys.argv) > 3 else 256
torch.manual_seed(0)
m = ae_b200.SupervisedAutoencoder(64, 10, precision=prec, backend=backend).to(dev).train()
x = torch.rand(B, 3, 64, 64, device=dev); y = torch.randint(0, 10, (B,), device=dev)
m.engine().prepare(dev, B)
opt = ae_b200.Adam(m.parameters(), lr=5e-3)
for i in range(3):
    opt.zero_grad(); loss = m.train_step_grads(x, y, 35.0); opt.step(); m.engine().prepare(dev, B)
torch.cuda.synchronize()
print('loss', loss.tolist())
