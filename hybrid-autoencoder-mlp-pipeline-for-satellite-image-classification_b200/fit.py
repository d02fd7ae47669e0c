"""The reference's autoencoder training loop (NB:2656-2727) with the epoch kept on the device.

Per epoch the reference makes two host round trips per batch (H2D of the fp32 batch, ``loss.item()``).  Here the split
lives in HBM as uint8 (``data.DeviceDataset``), batches are produced by the augmentation kernel, the training step is the
captured CUDA graph (``TrainStep``), per-step losses stay in a device history, and the host reads them once per epoch
phase to form the reference's epoch statistics (sample-weighted mean loss, early stopping).
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from .data import DeviceLoader
from .train import TrainStep


def weighted_mean(losses: torch.Tensor, sizes: List[int]) -> float:
    """NB:2686-2690: ``sum(loss.item() * bs) / sum(bs)`` in Python floats, as the reference accumulates it."""
    tot, n = 0.0, 0
    for v, b in zip(losses[:, 0].tolist(), sizes):
        tot += v * b
        n += b
    return tot / max(n, 1)


def train_epoch_ae(stepper: TrainStep, loader: DeviceLoader) -> float:
    """NB:2667-2691."""
    stepper.model.train()
    losses, sizes = stepper.run_loader(loader)
    return weighted_mean(losses, sizes)


@torch.no_grad()
def eval_epoch_ae(model, loader: DeviceLoader, alpha: float, batch_size: Optional[int] = None) -> float:
    """NB:2693-2714: eval-mode forward + alpha*MSE + CE over the validation loader, one host read at the end.
    `batch_size` (default: the loader's) only regroups the images: the sample-weighted mean of batch means is the mean."""
    model.eval()
    ds, dev = loader.dataset, loader.dataset.images.device
    bs = int(batch_size or loader.batch_size)
    n = len(ds)
    steps = (n + bs - 1) // bs
    hist = torch.zeros(steps, 4, dtype=torch.float32, device=dev)
    sizes = []
    order = torch.arange(n, device=dev)
    for k in range(steps):
        idx = order[k * bs:(k + 1) * bs]
        imgs = loader.transform(ds.images, idx)
        loss, _, _, _ = model.eval_step(imgs, ds.labels.index_select(0, idx), alpha)
        hist[k, :3].copy_(loss, non_blocking=True)
        sizes.append(int(idx.numel()))
    return weighted_mean(hist.cpu(), sizes)


def fit_autoencoder(model, optimizer, train_loader: DeviceLoader, val_loader: DeviceLoader, alpha: float,
                    num_epochs: int = 80, patience: int = 15, stepper: Optional[TrainStep] = None, log=None,
                    eval_batch: Optional[int] = 1024) -> Dict:
    """NB:2656-2727 for one (alpha, lr) configuration.  Returns the curves and the best validation loss; like the
    reference (NB:2735) the model is left in its LAST state.  `eval_batch`: images per validation launch (None: the
    validation loader's own batch size; a stepper passed in must have been captured on an engine prepared for it)."""
    own = stepper is None
    if own:
        # one engine (workspace) for both phases: the captured training graph addresses it
        model.engine().prepare(train_loader.dataset.images.device, max(train_loader.batch_size, eval_batch or val_loader.batch_size))
        stepper = TrainStep(model, optimizer, alpha, train_loader.batch_size, device=train_loader.dataset.images.device)
    best, counter = float("inf"), 0
    train_curve, val_curve = [], []
    try:
        for epoch in range(num_epochs):
            tl = train_epoch_ae(stepper, train_loader)
            vl = eval_epoch_ae(model, val_loader, alpha, eval_batch)
            train_curve.append(tl)
            val_curve.append(vl)
            if log:
                log(f"[AE alpha={alpha}] Epoch {epoch + 1} | TrainLoss={tl:.4f} | ValLoss={vl:.4f}")
            if vl < best:
                best, counter = vl, 0
            else:
                counter += 1
                if counter >= patience:
                    break
    finally:
        if own:
            stepper.close()
    return {"train_curve": train_curve, "val_curve": val_curve, "best_val_loss": best, "epochs": len(train_curve)}


# --------------------------------------------------------------------------------------------------
# Stage D: the latent classifier (NB:3443-3533) on device-resident latents
# --------------------------------------------------------------------------------------------------
def train_epoch_mlp(clf, optimizer, X: torch.Tensor, y: torch.Tensor, batch_size: int = 64, shuffle: bool = True,
                    generator: Optional[torch.Generator] = None):
    """NB:3471-3489 over latents X [N,D] / labels y [N] that already live on the device (the reference wraps the CPU
    copies in a TensorDataset, NB:3443).  Per batch: ONE replay of a captured two-launch graph (forward + CE + backward in
    one kernel -- one CTA with everything in shared memory up to 64 rows, a cluster above that -- and the fused flat Adam)
    that reads its rows through the epoch's permutation and records loss / correct count in a device history itself
    (`ae_mlp_train_step_indexed`); the history is read once per epoch.  Returns (train_loss, train_acc)."""
    from .train import MLPTrainStep
    clf.train()
    n, dev = int(X.shape[0]), X.device
    order = torch.randperm(n, device=dev, generator=generator) if shuffle else torch.arange(n, device=dev)
    steps = (n + batch_size - 1) // batch_size
    cache = clf.__dict__.setdefault("_train_steps", {})       # captured steps per (optimizer, batch size[, data])
    if clf._dropout_keep_override is not None:                 # test hook: the explicit-mask path is not graph-captured
        hist = torch.zeros(steps, 2, dtype=torch.float32, device=dev)
        sizes = []
        for k in range(steps):
            idx = order[k * batch_size:(k + 1) * batch_size]
            xb, yb = X.index_select(0, idx), y.index_select(0, idx)
            optimizer.zero_grad()
            loss, correct, _ = clf.fused_step_grads(xb, yb)
            optimizer.step()
            hist[k, 0:1].copy_(loss, non_blocking=True)
            hist[k, 1:2].copy_(correct, non_blocking=True)    # int32 -> float32 (exact below 2^24)
            sizes.append(int(idx.numel()))
    else:
        # Epoch mode: the captured step gathers its own batch through `order` and records (loss, correct) itself; the host
        # issues one graph replay per batch.  The buffers the graphs address live in the per-data cache entry.
        if not (X.dtype == torch.float32 and y.dtype == torch.int64 and X.is_contiguous() and y.is_contiguous()):
            raise TypeError("train_epoch_mlp: X must be contiguous float32 [N,D] and y contiguous int64 [N]")
        key = (id(optimizer), X.data_ptr(), y.data_ptr(), n, batch_size)
        ent = cache.get(key)
        if ent is None or ent["X"] is not X or ent["y"] is not y or not all(s_.valid() for s_ in ent["steps"].values()):
            for old_key in [k_ for k_ in cache if k_[0] == id(optimizer)]:
                del cache[old_key]                             # one data set per optimizer at a time: drop stale graphs
            ent = cache[key] = {"X": X, "y": y, "order": torch.zeros(n, dtype=torch.int64, device=dev),
                                "cursor": torch.zeros(2, dtype=torch.int64, device=dev),
                                "hist": torch.zeros(steps, 2, dtype=torch.float32, device=dev), "steps": {}}
        ent["order"].copy_(order)
        ent["cursor"].zero_()
        sizes = [min(batch_size, n - k * batch_size) for k in range(steps)]
        for b in sizes:
            step = ent["steps"].get(b)
            if step is None:
                step = ent["steps"][b] = MLPTrainStep(clf, optimizer, b, dev,
                                                      epoch=(X, y, ent["order"], ent["cursor"], ent["hist"]))
            step.run()
        hist = ent["hist"]
    h = hist.cpu()
    tot = sum(sizes)
    return sum(v * b for v, b in zip(h[:, 0].tolist(), sizes)) / tot, float(h[:, 1].sum()) / tot


@torch.no_grad()
def eval_epoch_mlp(clf, X: torch.Tensor, y: torch.Tensor, batch_size: int = 4096):
    """NB:3492-3507 (and the test pass NB:3523-3533): eval-mode logits, mean cross-entropy and accuracy.  The reference
    walks batches of 64 and forms the sample-weighted mean of the batch means -- the same quantity."""
    from . import _lib
    from ._lib import check, ptr, stream_ptr
    clf.eval()
    n, dev = int(X.shape[0]), X.device
    steps = (n + batch_size - 1) // batch_size
    loss = torch.zeros(steps, 4, dtype=torch.float32, device=dev)
    correct = torch.zeros(steps, dtype=torch.int32, device=dev)
    sizes = []
    for k in range(steps):
        xb, yb = X[k * batch_size:(k + 1) * batch_size], y[k * batch_size:(k + 1) * batch_size].contiguous()
        logits, _ = clf.predict(xb)
        check(_lib.load().ae_softmax_ce_fwd_bwd(ptr(logits), ptr(yb), int(xb.shape[0]), int(logits.shape[1]), 1.0, ptr(loss[k]),
                                                None, ptr(correct[k:k + 1]), stream_ptr()))
        sizes.append(int(xb.shape[0]))
    lh, ch = loss[:, 0].cpu().tolist(), correct.cpu().tolist()
    tot = sum(sizes)
    return sum(v * b for v, b in zip(lh, sizes)) / tot, sum(ch) / tot


def fit_mlp(clf, optimizer, train, val, num_epochs: int = 30, batch_size: int = 64, generator=None, log=None) -> Dict:
    """NB:3465-3519 for one learning rate: returns the curves, the best validation accuracy and the state dict of the
    best epoch (the reference's ``best_state_lr``; cloned here -- ``state_dict().copy()`` in the reference aliases the
    live tensors)."""
    (Xt, yt), (Xv, yv) = train, val
    curves = {"train_acc": [], "val_acc": [], "train_loss": [], "val_loss": []}
    best, best_state = 0.0, None
    for e in range(num_epochs):
        tl, ta = train_epoch_mlp(clf, optimizer, Xt, yt, batch_size, True, generator)
        vl, va = eval_epoch_mlp(clf, Xv, yv)
        for k, v in zip(("train_acc", "val_acc", "train_loss", "val_loss"), (ta, va, tl, vl)):
            curves[k].append(v)
        if log:
            log(f"Epoch {e + 1}/{num_epochs} | TrainAcc={ta:.3f} ValAcc={va:.3f}")
        if va > best:
            best = va
            best_state = {k: v.detach().clone() for k, v in clf.state_dict().items()}
    return {**curves, "best_val_acc": best, "best_state": best_state}
