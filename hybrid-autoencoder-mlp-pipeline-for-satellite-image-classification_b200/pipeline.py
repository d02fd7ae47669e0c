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
