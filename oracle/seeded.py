"""Deterministic, torch-version-independent parameter / input generation.

Test infrastructure (see oracle/__init__.py).  The reference initialises its modules with
torch's unseeded default init (U(-1/sqrt(fan_in), 1/sqrt(fan_in))).
Golden vectors need values that can be regenerated bit-for-bit on any box, so parameters
are drawn with numpy's MT19937 ``RandomState`` (stable across numpy versions) using the
same distribution family, and BatchNorm affine/running tensors are perturbed away from
(1, 0, 0, 1) so that every term of the BN arithmetic is exercised.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch


def ae_state_shapes(latent_dim: int = 64, num_classes: int = 10) -> "OrderedDict[str, tuple]":
    """state_dict keys and shapes of the reference SupervisedAutoencoder (NB:499-525, 607-635, 685-702)."""
    s: "OrderedDict[str, tuple]" = OrderedDict()
    chans = [3, 32, 64, 128, 256]
    for i in range(4):
        ci, co = chans[i], chans[i + 1]
        s[f"enc.encoder.{3 * i}.weight"] = (co, ci, 3, 3)
        s[f"enc.encoder.{3 * i}.bias"] = (co,)
        b = f"enc.encoder.{3 * i + 1}"
        s[b + ".weight"] = (co,)
        s[b + ".bias"] = (co,)
        s[b + ".running_mean"] = (co,)
        s[b + ".running_var"] = (co,)
        s[b + ".num_batches_tracked"] = ()
    s["enc.encoder.13.weight"] = (latent_dim, 4096)
    s["enc.encoder.13.bias"] = (latent_dim,)
    s["dec.decoder_input.weight"] = (4096, latent_dim)
    s["dec.decoder_input.bias"] = (4096,)
    dch = [256, 128, 64, 32, 3]
    for i in range(4):
        ci, co = dch[i], dch[i + 1]
        s[f"dec.decoder.{3 * i + 1}.weight"] = (ci, co, 3, 3)
        s[f"dec.decoder.{3 * i + 1}.bias"] = (co,)
        if i < 3:
            b = f"dec.decoder.{3 * i + 2}"
            s[b + ".weight"] = (co,)
            s[b + ".bias"] = (co,)
            s[b + ".running_mean"] = (co,)
            s[b + ".running_var"] = (co,)
            s[b + ".num_batches_tracked"] = ()
    s["classifier.0.weight"] = (128, latent_dim)
    s["classifier.0.bias"] = (128,)
    s["classifier.2.weight"] = (num_classes, 128)
    s["classifier.2.bias"] = (num_classes,)
    return s


def mlp_state_shapes(input_dim: int = 64, num_classes: int = 10) -> "OrderedDict[str, tuple]":
    """state_dict keys and shapes of the reference MLP (NB:2970-2987)."""
    s: "OrderedDict[str, tuple]" = OrderedDict()
    s["net.0.weight"] = (128, input_dim)
    s["net.0.bias"] = (128,)
    for b, c in (("net.1", 128),):
        s[b + ".weight"] = (c,)
        s[b + ".bias"] = (c,)
        s[b + ".running_mean"] = (c,)
        s[b + ".running_var"] = (c,)
        s[b + ".num_batches_tracked"] = ()
    s["net.4.weight"] = (64, 128)
    s["net.4.bias"] = (64,)
    for b, c in (("net.5", 64),):
        s[b + ".weight"] = (c,)
        s[b + ".bias"] = (c,)
        s[b + ".running_mean"] = (c,)
        s[b + ".running_var"] = (c,)
        s[b + ".num_batches_tracked"] = ()
    s["net.7.weight"] = (num_classes, 64)
    s["net.7.bias"] = (num_classes,)
    return s


def _is_bn(key: str, shapes) -> bool:
    return (key.rsplit(".", 1)[0] + ".running_mean") in shapes


def seeded_state(shapes, seed: int) -> "OrderedDict[str, torch.Tensor]":
    """Fill every tensor of ``shapes`` from RandomState(seed*1000 + index)."""
    out: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for idx, (k, shp) in enumerate(shapes.items()):
        rs = np.random.RandomState(seed * 1000 + idx)
        leaf = k.rsplit(".", 1)[1]
        if leaf == "num_batches_tracked":
            out[k] = torch.tensor(0, dtype=torch.int64)
            continue
        if _is_bn(k, shapes):
            n = rs.standard_normal(shp).astype(np.float32)
            if leaf == "weight":
                v = 1.0 + 0.1 * n
            elif leaf == "running_var":
                v = 1.0 + 0.1 * np.abs(n)
            else:  # bias, running_mean
                v = 0.1 * n
        else:
            base = k.rsplit(".", 1)[0] + ".weight"
            wshape = shapes[base]
            if len(wshape) == 4:
                # torch quirk: fan_in = weight.size(1) * k*k for both Conv2d and ConvTranspose2d
                fan_in = wshape[1] * wshape[2] * wshape[3]
            else:
                fan_in = wshape[1]
            bound = 1.0 / math.sqrt(fan_in)
            v = rs.uniform(-bound, bound, size=shp)
        out[k] = torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32))
    return out


def seeded_images(batch: int, seed: int) -> torch.Tensor:
    """Class-free uniform images in [0,1), shape [B,3,64,64] fp32 (BASELINE config 1/2 inputs)."""
    rs = np.random.RandomState(7000 + seed)
    return torch.from_numpy(rs.random_sample((batch, 3, 64, 64)).astype(np.float32))


def seeded_labels(batch: int, seed: int, num_classes: int = 10) -> torch.Tensor:
    rs = np.random.RandomState(9000 + seed)
    return torch.from_numpy(rs.randint(0, num_classes, size=(batch,)).astype(np.int64))


def structured_images(labels: torch.Tensor, seed: int, noise: float = 0.03) -> torch.Tensor:
    """Class-structured synthetic EuroSAT-shaped images: a per-class
    colour mean plus a low-frequency pattern plus N(0, noise^2), clipped to [0,1]."""
    rs = np.random.RandomState(11000 + seed)
    n = int(labels.numel())
    ncls = 10
    means = np.random.RandomState(424242).uniform(0.2, 0.8, size=(ncls, 3)).astype(np.float32)
    freq = np.random.RandomState(434343).uniform(0.5, 3.0, size=(ncls, 2)).astype(np.float32)
    yy, xx = np.meshgrid(np.linspace(0, 1, 64, dtype=np.float32), np.linspace(0, 1, 64, dtype=np.float32), indexing="ij")
    lab = labels.numpy()
    img = np.empty((n, 3, 64, 64), dtype=np.float32)
    phase = rs.uniform(0, 2 * np.pi, size=(n, 2)).astype(np.float32)
    for c in range(3):
        pat = 0.15 * np.sin(2 * np.pi * freq[lab, 0, None, None] * yy[None] + phase[:, 0, None, None] + c) * \
            np.cos(2 * np.pi * freq[lab, 1, None, None] * xx[None] + phase[:, 1, None, None])
        img[:, c] = means[lab, c, None, None] + pat
    img += noise * rs.standard_normal(img.shape).astype(np.float32)
    return torch.from_numpy(np.clip(img, 0.0, 1.0).astype(np.float32))
