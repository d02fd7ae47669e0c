"""Times whole training epochs through the on-device loop (TrainStep.run_loader): uint8 split resident in HBM, augmentation
kernel + captured step per batch, one host sync per epoch.  BASELINE configs[2] shape: 18,900 training images."""
import sys, os, time; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ae_b200
dev = torch.device('cuda', 0)
n, bs = 18900, int(sys.argv[1]) if len(sys.argv) > 1 else 256
torch.manual_seed(0)
imgs = torch.randint(0, 256, (n, 64, 64, 3), dtype=torch.uint8, device=dev)
labels = torch.randint(0, 10, (n,), device=dev)
ds = ae_b200.DeviceDataset(imgs, labels)
g = torch.Generator(device=dev).manual_seed(1)
loader = ae_b200.DeviceLoader(ds, bs, shuffle=True, transform=ae_b200.TrainTransformAE(generator=g, seed=2), generator=g)
model = ae_b200.SupervisedAutoencoder(64, 10).to(dev).train()
model.engine().prepare(dev, bs)
opt = ae_b200.Adam(model.parameters(), lr=5e-3)
st = ae_b200.TrainStep(model, opt, 35.0, bs)
for ep in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    losses, sizes = st.run_loader(loader)
    dt = time.perf_counter() - t0
    print(f"epoch {ep}: {len(sizes)} steps, {n / dt:,.0f} images/s, {dt * 1e3:.1f} ms, loss {ae_b200.fit.weighted_mean(losses, sizes):.4f}", flush=True)
