import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ae_b200, numpy as np
from oracle import seeded, torch_port as tp
dev = torch.device('cuda', 0)
def rel(a, b): return float((a.detach().double().cpu() - b.double()).abs().max() / b.double().abs().max())
for B in (8, 33, 256):
    st = seeded.seeded_state(seeded.ae_state_shapes(64, 10), 11)
    x, y = seeded.seeded_images(B, 11), seeded.seeded_labels(B, 11)
    s32 = {k: v.clone() for k, v in st.items()}
    loss, _, _, g32, _ = tp.ae_train_step(s32, {}, x, y, 35.0, 5e-3)
    s64 = {k: (v.double() if v.dtype == torch.float32 else v.clone()) for k, v in st.items()}
    loss64, _, _, g64, _ = tp.ae_train_step(s64, {}, x.double(), y, 35.0, 5e-3)
    m = ae_b200.SupervisedAutoencoder(64, 10, backend='simt')
    m.load_state_dict(st); m = m.to(dev).train()
    out = m.train_step_grads(x.to(dev), y.to(dev), 35.0)
    torch.cuda.synchronize()
    print(f'B={B} loss gpu {float(out[0]):.7f} cpu32 {float(loss):.7f} cpu64 {float(loss64):.7f}')
    for k, p in m.named_parameters():
        print(f'  {k:28s} gpu-vs-64 {rel(p.grad, g64[k]):.2e}  cpu32-vs-64 {rel(g32[k], g64[k]):.2e}  gpu-vs-32 {rel(p.grad, g32[k]):.2e}  max|g| {float(g64[k].abs().max()):.2e}')
