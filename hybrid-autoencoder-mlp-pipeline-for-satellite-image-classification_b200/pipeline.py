"""Callers on either side of the hot path, kept on the device.

extract_features  NB:2902-2914: encoder-only inference over batches, latents concatenated.  The reference
                  moves every batch to the CPU (`z.cpu()`); here latents stay on the GPU unless asked.
encode_predict    BASELINE config 1/5: clf(enc(x)).argmax(1) in eval mode (NB:2908 composed with NB:3702).
"""
from __future__ import annotations

import torch


@torch.no_grad()
def extract_features(loader, encoder, device=None, to_cpu: bool = False):
    X_list, y_list = [], []
    encoder.eval()
    dev = device
    for imgs, labels in loader:
        if dev is None:
            dev = imgs.device if imgs.is_cuda else torch.device("cuda", torch.cuda.current_device())
        imgs = imgs.to(dev, non_blocking=True)
        z = encoder(imgs)
        X_list.append(z.cpu() if to_cpu else z)
        y_list.append(labels.cpu() if to_cpu else labels.to(dev, non_blocking=True))
    return torch.cat(X_list, dim=0), torch.cat(y_list, dim=0)


@torch.no_grad()
def encode_predict(encoder, clf, imgs):
    """Returns (z, logits, argmax) for a batch of images, all on the device."""
    encoder.eval()
    clf.eval()
    z = encoder(imgs)
    logits, am = clf.predict(z)
    return z, logits, am


def run_pipeline(train_ds, val_ds, test_ds, *, alpha: float = 35.0, ae_lr: float = 5e-3, mlp_lr: float = 1e-4,
                 ae_epochs: int = 80, ae_patience: int = 15, mlp_epochs: int = 30, ae_batch: int = 64, mlp_batch: int = 64,
                 latent_dim: int = 64, num_classes: int = 10, precision=None, generator=None, seed=None, log=None):
    """The reference's two-stage pipeline for ONE hyper-parameter configuration (BASELINE configs[2]), entirely on the
    device: supervised autoencoder training with early stopping (NB:2650-2727), encoder frozen (NB:3434-3436), latents
    of the three splits extracted through the same loaders as the reference -- the training split WITH its
    augmentation, NB:3439 -- and kept on the device, then the MLP (NB:3460-3533).  Splits are ``data.DeviceDataset``s.
    Returns the curves, the best validation accuracy, the test accuracy of that state and per-stage wall times."""
    import time
    from . import fit
    from .data import DeviceLoader, TrainTransformAE
    from .modules import MLP, SupervisedAutoencoder
    from .optim import Adam

    dev = train_ds.images.device
    tf = TrainTransformAE(generator=generator, seed=seed)
    train_loader = DeviceLoader(train_ds, ae_batch, shuffle=True, transform=tf, generator=generator)
    val_loader = DeviceLoader(val_ds, ae_batch, shuffle=False)
    test_loader = DeviceLoader(test_ds, ae_batch, shuffle=False)
    t0 = time.perf_counter()
    ae = SupervisedAutoencoder(latent_dim, num_classes, precision=precision).to(dev)
    ae.engine().prepare(dev, ae_batch)
    opt = Adam(ae.parameters(), lr=ae_lr)
    ae_res = fit.fit_autoencoder(ae, opt, train_loader, val_loader, alpha, ae_epochs, ae_patience, log=log)
    torch.cuda.synchronize(dev)
    t1 = time.perf_counter()
    for p in ae.enc.parameters():
        p.requires_grad = False
    ae.enc.eval()
    big = lambda ld: DeviceLoader(ld.dataset, 4096, ld.shuffle, ld.transform, ld.generator)   # same batches' content, fewer launches
    Xt, yt = extract_features(big(train_loader), ae.enc)
    Xv, yv = extract_features(big(val_loader), ae.enc)
    Xs, ys = extract_features(big(test_loader), ae.enc)
    torch.cuda.synchronize(dev)
    t2 = time.perf_counter()
    clf = MLP(latent_dim, num_classes).to(dev)
    clf._state.prepare(dev, mlp_batch)                          # flat parameter storage exists before the optimizer binds to it
    copt = Adam(clf.parameters(), lr=mlp_lr, weight_decay=1e-4)
    mlp_res = fit.fit_mlp(clf, copt, (Xt, yt), (Xv, yv), mlp_epochs, mlp_batch, generator, log)
    if mlp_res["best_state"] is not None:
        clf.load_state_dict(mlp_res["best_state"])
    _, test_acc = fit.eval_epoch_mlp(clf, Xs, ys)
    torch.cuda.synchronize(dev)
    t3 = time.perf_counter()
    return {"ae": ae_res, "mlp": {k: v for k, v in mlp_res.items() if k != "best_state"}, "test_acc": test_acc,
            "seconds": {"autoencoder": t1 - t0, "extract": t2 - t1, "mlp": t3 - t2, "total": t3 - t0},
            "model": ae, "clf": clf}
