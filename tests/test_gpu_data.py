"""-m gpu: the device-side data path against the reference's transforms and loop.

ae_augment_u8 is byte / fp32-exact work: it must reproduce the reference's torchvision pipeline (NB:386-395) bit for
bit when given the same random draws (golden fixture recorded from the reference's own Compose objects), and the
oracle restatement on arbitrary draws, including the extreme crop offsets.
"""
import ctypes as C
import os

import numpy as np
import pytest
import torch

import ae_b200
from ae_b200 import _lib
from oracle import augment_port as ap, seeded, torch_port as tp
from tests import gpu_util as gu

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "augment_B6.npz")


def test_augment_matches_reference_golden_bit_exact():
    g = np.load(GOLD)
    d = gu.dev()
    imgs = torch.from_numpy(g["images"]).to(d)
    out = ae_b200.augment_u8(imgs, None, torch.from_numpy(g["flip"]), torch.from_numpy(g["off_y"]), torch.from_numpy(g["off_x"]),
                             4, torch.from_numpy(g["noise"]), noise_std=0.03, noise_mean=0.0)
    torch.cuda.synchronize()
    assert torch.equal(out.cpu(), torch.from_numpy(g["train_out"]))
    ev = ae_b200.EvalTransform()(imgs)
    assert torch.equal(ev.cpu(), torch.from_numpy(g["eval_out"]))


@pytest.mark.parametrize("batch", [1, 7, 64])
def test_augment_gather_flip_crop_extremes_vs_oracle(batch):
    rs = np.random.RandomState(batch)
    n = 19
    imgs = ap.synthetic_images(n, seed=3 + batch)
    index = rs.randint(0, n, size=batch)
    flip = rs.randint(0, 2, size=batch).astype(np.uint8)
    oy, ox = rs.randint(0, 9, size=batch).astype(np.int32), rs.randint(0, 9, size=batch).astype(np.int32)
    oy[0], ox[0] = 0, 8                                         # extreme offsets: 4 rows / columns of zero fill
    if batch > 1:
        oy[1], ox[1], flip[1] = 8, 0, 1
    noise = torch.from_numpy(rs.standard_normal((batch, 3, 64, 64)).astype(np.float32))
    d = gu.dev()
    out = ae_b200.augment_u8(torch.from_numpy(imgs).to(d), torch.from_numpy(index), torch.from_numpy(flip), torch.from_numpy(oy),
                             torch.from_numpy(ox), 4, noise, noise_std=0.03, noise_mean=0.01)
    torch.cuda.synchronize()
    for b in range(batch):
        ref = ap.train_transform(imgs[index[b]], flip[b], oy[b], ox[b], noise[b], std=0.03, mean=0.01)
        assert torch.equal(out[b].cpu(), ref), b


def test_augment_device_noise_statistics_and_streams():
    d = gu.dev()
    imgs = torch.zeros(64, 64, 64, 3, dtype=torch.uint8, device=d)
    a = ae_b200.augment_u8(imgs, noise_std=0.03, seed=1234)
    b = ae_b200.augment_u8(imgs, noise_std=0.03, seed=1234)
    c = ae_b200.augment_u8(imgs, noise_std=0.03, seed=1235)
    torch.cuda.synchronize()
    assert torch.equal(a, b) and not torch.equal(a, c)
    v = a.double().flatten()
    n = v.numel()
    assert abs(float(v.mean())) <= 5 * 0.03 / n ** 0.5
    assert abs(float(v.std()) - 0.03) <= 0.03 * 5 / (2 * n) ** 0.5
    z = v / 0.03
    assert abs(float((z ** 3).mean())) <= 0.02 and abs(float((z ** 4).mean()) - 3.0) <= 0.05      # skewness, kurtosis of N(0,1)
    assert 0.6 < float((z.abs() < 1).double().mean()) / 0.6827 < 1.4 and float(z.abs().max()) < 7.0
    # neighbouring values are independent draws
    assert abs(float((z[:-1] * z[1:]).mean())) <= 5 / n ** 0.5


def test_device_loader_covers_every_image_once_and_keeps_the_last_batch():
    d = gu.dev()
    n, bs = 150, 64
    imgs = torch.from_numpy(ap.synthetic_images(n, seed=5)).to(d)
    labels = torch.arange(n) % 10
    ds = ae_b200.DeviceDataset(imgs, labels)
    g = torch.Generator(device=d).manual_seed(7)
    loader = ae_b200.DeviceLoader(ds, bs, shuffle=True, transform=ae_b200.TrainTransformAE(generator=g, seed=3), generator=g)
    assert len(loader) == 3
    seen, sizes = [], []
    for idx in loader.batches():
        seen.append(idx.cpu())
        sizes.append(int(idx.numel()))
    assert sizes == [64, 64, 22] and sorted(torch.cat(seen).tolist()) == list(range(n))
    assert not torch.equal(torch.cat(seen), torch.arange(n))
    for x, y in loader:
        assert x.shape[1:] == (3, 64, 64) and x.dtype == torch.float32 and x.is_cuda and y.dtype == torch.int64
    # eval pipeline, no shuffle: batches are ToTensor of consecutive images with their labels
    ev = ae_b200.DeviceLoader(ds, bs, shuffle=False)
    x0, y0 = next(iter(ev))
    assert torch.equal(x0.cpu(), torch.stack([ap.to_tensor(imgs[i].cpu().numpy()) for i in range(bs)]))
    assert torch.equal(y0.cpu(), labels[:bs])


def test_on_device_epoch_matches_oracle_loop():
    """TrainStep.run_loader (NB:2672-2688 with the epoch on the device) against the oracle's loop on the same batches:
    eval transform (deterministic), no shuffle, 40 images in batches of 16 -> 16, 16, 8 (the last batch goes through a
    second captured graph)."""
    seed, alpha, lr, bs, n = 33, 35.0, 1e-4, 16, 40
    d = gu.dev()
    imgs_u8 = ap.synthetic_images(n, seed=9)
    labels = torch.from_numpy(np.random.RandomState(1).randint(0, 10, size=n))
    st = seeded.seeded_state(seeded.ae_state_shapes(64, 10), seed)
    ref_state = {k: v.clone() for k, v in st.items()}
    model = ae_b200.SupervisedAutoencoder(64, 10, precision="fp32")
    model.load_state_dict(st)
    model = model.to(d).train()
    model.engine().prepare(d, bs)
    opt = ae_b200.Adam(model.parameters(), lr=lr)
    stepper = ae_b200.TrainStep(model, opt, alpha, bs)
    loader = ae_b200.DeviceLoader(ae_b200.DeviceDataset(torch.from_numpy(imgs_u8), labels, device=d), bs, shuffle=False)
    losses, sizes = stepper.run_loader(loader)
    assert sizes == [16, 16, 8] and losses.shape == (3, 3)
    ref_opt, ref_losses = {}, []
    for i in range(0, n, bs):
        x = torch.stack([ap.to_tensor(imgs_u8[j]) for j in range(i, min(n, i + bs))])
        loss, *_ = tp.ae_train_step(ref_state, ref_opt, x, labels[i:i + bs], alpha, lr)
        ref_losses.append(float(loss))
    for k in range(3):
        assert abs(float(losses[k, 0]) - ref_losses[k]) <= 5e-3 * abs(ref_losses[k]), (k, float(losses[k, 0]), ref_losses[k])
    assert abs(float(losses[0, 0]) - ref_losses[0]) <= 2e-4 * abs(ref_losses[0])
    ep = ae_b200.fit.weighted_mean(losses, sizes)
    assert abs(ep - sum(l * b for l, b in zip(ref_losses, sizes)) / n) <= 5e-3 * abs(ep)
    # the validation phase of the epoch (eval mode, one host read)
    vl = ae_b200.fit.eval_epoch_ae(model, loader, alpha)
    with torch.no_grad():
        tot = 0.0
        for i in range(0, n, bs):
            x = torch.stack([ap.to_tensor(imgs_u8[j]) for j in range(i, min(n, i + bs))])
            x_hat, logits, _ = tp.ae_forward(ref_state, x, False)
            l = alpha * torch.nn.functional.mse_loss(x_hat, x) + torch.nn.functional.cross_entropy(logits, labels[i:i + bs])
            tot += float(l) * x.shape[0]
    assert abs(vl - tot / n) <= 5e-3 * abs(tot / n), (vl, tot / n)
    stepper.close()


def test_fit_autoencoder_early_stopping_runs_on_device():
    d = gu.dev()
    n = 96
    rs = np.random.RandomState(0)
    labels = torch.from_numpy(rs.randint(0, 10, size=n))
    imgs = torch.from_numpy(ap.synthetic_images(n, seed=21))
    ds = ae_b200.DeviceDataset(imgs, labels, device=d)
    g = torch.Generator(device=d).manual_seed(1)
    train = ae_b200.DeviceLoader(ds, 32, shuffle=True, transform=ae_b200.TrainTransformAE(generator=g, seed=5), generator=g)
    val = ae_b200.DeviceLoader(ds, 32, shuffle=False)
    torch.manual_seed(0)
    model = ae_b200.SupervisedAutoencoder(64, 10).to(d)
    model.engine().prepare(d, 32)
    opt = ae_b200.Adam(model.parameters(), lr=1e-3)
    res = ae_b200.fit.fit_autoencoder(model, opt, train, val, alpha=35.0, num_epochs=4, patience=2)
    assert 1 <= res["epochs"] <= 4 and len(res["val_curve"]) == res["epochs"]
    assert all(np.isfinite(v) for v in res["train_curve"] + res["val_curve"])
    assert res["train_curve"][-1] < res["train_curve"][0]          # it learns


def _structured_u8(n_per_class, seed):
    labels = torch.arange(10).repeat_interleave(n_per_class)
    x = seeded.structured_images(labels, seed)                       # [N,3,64,64] fp32 in [0,1]
    u8 = (x * 255.0).round().clamp(0, 255).to(torch.uint8).permute(0, 2, 3, 1).contiguous()
    perm = torch.from_numpy(np.random.RandomState(seed).permutation(labels.numel()))
    return u8[perm], labels[perm]


def test_mlp_eval_epoch_matches_oracle():
    d = gu.dev()
    clf = ae_b200.MLP(64, 10).to(d)
    st = gu.load_mlp(clf, 5)
    clf = clf.to(d).eval()
    rs = np.random.RandomState(2)
    X = torch.from_numpy(rs.standard_normal((5000, 64)).astype(np.float32))
    y = torch.from_numpy(rs.randint(0, 10, size=5000))
    loss, acc = ae_b200.fit.eval_epoch_mlp(clf, X.to(d), y.to(d), batch_size=2048)      # 2048, 2048, 904
    with torch.no_grad():
        logits = tp.mlp_forward(st, X, False)
    ref_loss = float(torch.nn.functional.cross_entropy(logits.double(), y))
    ref_acc = float((logits.argmax(1) == y).double().mean())
    assert abs(loss - ref_loss) <= 1e-5 * abs(ref_loss) and abs(acc - ref_acc) <= 2 / 5000


def test_full_pipeline_on_class_structured_data():
    """BASELINE configs[2] at test scale: AE training with early stopping -> frozen encoder -> latents on the device ->
    MLP, all through the device-side loaders.  The classes are separable, so the pipeline must learn them."""
    d = gu.dev()
    tr = ae_b200.DeviceDataset(*_structured_u8(60, 1), device=d)
    va = ae_b200.DeviceDataset(*_structured_u8(15, 2), device=d)
    te = ae_b200.DeviceDataset(*_structured_u8(15, 3), device=d)
    torch.manual_seed(0)
    g = torch.Generator(device=d).manual_seed(0)
    res = ae_b200.pipeline.run_pipeline(tr, va, te, alpha=35.0, ae_lr=2e-3, mlp_lr=5e-3, ae_epochs=6, ae_patience=3,
                                        mlp_epochs=12, generator=g, seed=4)
    assert res["ae"]["epochs"] >= 1 and res["ae"]["train_curve"][-1] < res["ae"]["train_curve"][0]
    assert res["mlp"]["best_val_acc"] > 0.6 and res["test_acc"] > 0.6, (res["mlp"]["best_val_acc"], res["test_acc"])
    assert not any(p.requires_grad for p in res["model"].enc.parameters())
    # the latents the MLP saw are the encoder's: predict through the public inference path and compare accuracies
    x = ae_b200.EvalTransform()(te.images)
    _, _, am = ae_b200.encode_predict(res["model"].enc, res["clf"], x)
    assert abs(float((am == te.labels).float().mean()) - res["test_acc"]) <= 1e-6


def test_grid_search_single_gpu_runs_the_reference_grid_order():
    """NB:2629-2741 / NB:3447-3540 through search.fan_out on one device (the multi-rank exchange is covered on CPU with
    gloo in tests/test_search_cpu.py): every configuration runs, the winner follows the strict-improvement rule."""
    d = gu.dev()
    tr = ae_b200.DeviceDataset(*_structured_u8(24, 11), device=d)
    va = ae_b200.DeviceDataset(*_structured_u8(8, 12), device=d)
    g = torch.Generator(device=d).manual_seed(3)
    train = ae_b200.DeviceLoader(tr, 48, shuffle=True, transform=ae_b200.TrainTransformAE(generator=g, seed=1), generator=g)
    val = ae_b200.DeviceLoader(va, 48)
    torch.manual_seed(0)
    res = ae_b200.search.grid_search_autoencoder([10.0, 35.0], [1e-3, 5e-3], train, val, num_epochs=2, patience=2)
    assert len(res["results"]) == 4 and all(r is not None and r["epochs"] == 2 for r in res["results"])
    losses = [r["best_val_loss"] for r in res["results"]]
    assert res["best_index"] == losses.index(min(losses)) and res["best_config"] == ae_b200.search.ae_grid([10.0, 35.0], [1e-3, 5e-3])[res["best_index"]]
    model = ae_b200.SupervisedAutoencoder(64, 10)
    model.load_state_dict(res["best_state"])                           # the reference's checkpoint keys (NB:2735, NB:3431)
    model = model.to(d).eval()
    X, y = ae_b200.extract_features(val, model.enc)
    Xt, yt = ae_b200.extract_features(ae_b200.DeviceLoader(tr, 48), model.enc)
    mres = ae_b200.search.grid_search_mlp([1e-3, 1e-2], (Xt, yt), (X, y), num_epochs=3)
    assert len(mres["results"]) == 2 and mres["best_index"] in (0, 1)
    accs = [r["best_val_acc"] for r in mres["results"]]
    assert mres["best"]["best_val_acc"] == max(accs)
    clf = ae_b200.MLP(64, 10)
    clf.load_state_dict(mres["best_state"])
