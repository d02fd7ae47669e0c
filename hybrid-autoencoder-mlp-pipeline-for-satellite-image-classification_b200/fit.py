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
def eval_epoch_ae(model, loader: DeviceLoader, alpha: float) -> float:
    """NB:2693-2714: eval-mode forward + alpha*MSE + CE over the validation loader, one host read at the end."""
    model.eval()
    dev = loader.dataset.images.device
    hist = torch.zeros(len(loader), 4, dtype=torch.float32, device=dev)
    sizes = []
    for k, (imgs, labels) in enumerate(loader):
        loss, _, _, _ = model.eval_step(imgs, labels, alpha)
        hist[k, :3].copy_(loss, non_blocking=True)
        sizes.append(int(imgs.shape[0]))
    return weighted_mean(hist[:len(sizes)].cpu(), sizes)


def fit_autoencoder(model, optimizer, train_loader: DeviceLoader, val_loader: DeviceLoader, alpha: float,
                    num_epochs: int = 80, patience: int = 15, stepper: Optional[TrainStep] = None, log=None) -> Dict:
    """NB:2656-2727 for one (alpha, lr) configuration.  Returns the curves, the best validation loss and the state dict
    of the best epoch's model is left to the caller (the reference keeps the LAST state, NB:2735)."""
    own = stepper is None
    if own:
        stepper = TrainStep(model, optimizer, alpha, train_loader.batch_size, device=train_loader.dataset.images.device)
    best, counter = float("inf"), 0
    train_curve, val_curve = [], []
    try:
        for epoch in range(num_epochs):
            tl = train_epoch_ae(stepper, train_loader)
            vl = eval_epoch_ae(model, val_loader, alpha)
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
