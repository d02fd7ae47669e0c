"""Where does an epoch of fit_autoencoder go?  Times the training and the validation phase separately."""
import sys, os, time; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ae_b200
dev = torch.device('cuda', 0)
bs = int(sys.argv[1]) if len(sys.argv) > 1 else 64
torch.manual_seed(0)
mk = lambda n: ae_b200.DeviceDataset(torch.randint(0, 256, (n, 64, 64, 3), dtype=torch.uint8, device=dev), torch.randint(0, 10, (n,), device=dev))
tr, va = mk(18900), mk(4050)
g = torch.Generator(device=dev).manual_seed(1)
train = ae_b200.DeviceLoader(tr, bs, shuffle=True, transform=ae_b200.TrainTransformAE(generator=g, seed=2), generator=g)
val = ae_b200.DeviceLoader(va, bs, shuffle=False)
model = ae_b200.SupervisedAutoencoder(64, 10).to(dev)
model.engine().prepare(dev, max(bs, 1024))
opt = ae_b200.Adam(model.parameters(), lr=5e-3)
st = ae_b200.TrainStep(model, opt, 35.0, bs)
for ep in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    tl = ae_b200.fit.train_epoch_ae(st, train)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    vl = ae_b200.fit.eval_epoch_ae(model, val, 35.0)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    vl2 = ae_b200.fit.eval_epoch_ae(model, val, 35.0, 1024)
    torch.cuda.synchronize(); t3 = time.perf_counter()
    print(f"epoch {ep} bs={bs}: train {1e3*(t1-t0):.1f} ms ({18900/(t1-t0):,.0f} img/s)  val(bs) {1e3*(t2-t1):.1f} ms  val(1024) {1e3*(t3-t2):.1f} ms  "
          f"loss {tl:.4f} val {vl:.5f} {vl2:.5f}", flush=True)
