import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ae_b200, numpy as np
from oracle import seeded, torch_port as tp
dev = torch.device('cuda', 0)
seed, alpha, lr, batch, steps = 21, 35.0, 5e-3, 16, 4
def l2(a, b): return float((a.detach().double().cpu() - b.double()).norm() / b.double().norm())
st = seeded.seeded_state(seeded.ae_state_shapes(64, 10), seed)
for mode in ('graph', 'eager'):
    ref_state = {k: v.clone() for k, v in st.items()}
    ref64 = {k: (v.double() if v.dtype == torch.float32 else v.clone()) for k, v in st.items()}
    opt, opt64 = {}, {}
    model = ae_b200.SupervisedAutoencoder(64, 10, backend='simt')
    model.load_state_dict(st); model = model.to(dev).train(); model.engine().prepare(dev, batch)
    optimizer = ae_b200.Adam(model.parameters(), lr=lr)
    stepper = ae_b200.TrainStep(model, optimizer, alpha, batch) if mode == 'graph' else None
    for s in range(steps):
        x, y = seeded.seeded_images(batch, seed + s), seeded.seeded_labels(batch, seed + s)
        loss, _, _, g32, _ = tp.ae_train_step(ref_state, opt, x, y, alpha, lr)
        loss64, _, _, g64, _ = tp.ae_train_step(ref64, opt64, x.double(), y, alpha, lr)
        if mode == 'graph':
            got = stepper(x, y); torch.cuda.synchronize()
        else:
            optimizer.zero_grad()
            got = model.train_step_grads(x.to(dev), y.to(dev), alpha)
            grads = {k: p.grad.clone() for k, p in model.named_parameters()}
            optimizer.step(); torch.cuda.synchronize()
        print(mode, 'step', s, 'loss', float(got[0]), float(loss), float(loss64))
        for k in ['enc.encoder.0.weight', 'enc.encoder.3.weight', 'dec.decoder.10.weight', 'classifier.2.weight']:
            p = dict(model.named_parameters())[k]
            d = (p.detach().cpu() - ref_state[k]).abs(); d64 = (ref_state[k].double() - ref64[k]).abs()
            extra = f" gradL2 {l2(grads[k], g64[k]):.2e} cpu32gradL2 {l2(g32[k], g64[k]):.2e}" if mode == 'eager' else ''
            print(f'   {k:24s} max|dp| {float(d.max()):.2e} frac>0.1lr {float((d > 0.1*lr).float().mean()):.3f} | cpu32-vs-64 max {float(d64.max()):.2e} frac {float((d64 > 0.1*lr).float().mean()):.3f}{extra}')
