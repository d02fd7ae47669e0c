import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ae_b200, numpy as np
from oracle import seeded, torch_port as tp
dev = torch.device('cuda', 0)
def rel(a, b): return float((a.double().cpu() - b.double()).abs().max() / b.double().abs().max())
for B in (6, 8):
    st = seeded.seeded_state(seeded.ae_state_shapes(64, 10), 3)
    x, y = seeded.seeded_images(B, 3), seeded.seeded_labels(B, 3)
    nb = {}
    with torch.no_grad():
        xh, lg, z = tp.ae_forward(st, x, True, nb)
    m = ae_b200.SupervisedAutoencoder(64, 10, backend='simt')
    m.load_state_dict(st); m = m.to(dev).train()
    xh2, lg2, z2 = m(x.to(dev))
    print('dropin train B', B, 'z', rel(z2, z), 'logits', rel(lg2, lg), 'xhat', rel(xh2, xh))
    # warm-up then reload
    m2 = ae_b200.SupervisedAutoencoder(64, 10, backend='simt')
    m2.load_state_dict(st); m2 = m2.to(dev).train()
    m2(seeded.seeded_images(2, 99).to(dev))
    m2.load_state_dict(st)
    xh3, lg3, z3 = m2(x.to(dev))
    print('after warmup+reload B', B, 'z', rel(z3, z), 'logits', rel(lg3, lg), 'xhat', rel(xh3, xh))
    for k in ['enc.encoder.1.running_mean', 'dec.decoder.8.running_var']:
        print(' ', k, rel(m2.state_dict()[k], nb[k]))
