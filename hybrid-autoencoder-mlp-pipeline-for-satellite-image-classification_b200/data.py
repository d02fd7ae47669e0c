"""Device-side data path: the reference's transforms and DataLoader batches without leaving the GPU.

The reference decodes PIL images one by one on the host (``TransformDataset.__getitem__`` NB:340-343, ``DataLoader(...,
num_workers=0)`` NB:420) and ships an fp32 batch per step.  Here the split's images stay resident in HBM as uint8 HWC
(12 KB per image: the whole 27k-image set is 332 MB) and one kernel (``ae_augment_u8``) produces the fp32 NCHW batch:
gather by shuffled index, RandomHorizontalFlip, RandomCrop(64, padding=4), ToTensor, AddGaussianNoise(0, 0.03)
(NB:386-391) -- or plain ToTensor for the validation / test pipelines (NB:393-395).

PyTorch draws the per-image random decisions (flip bit, crop offsets, permutation: a few bytes per image) on the
device; the kernel does every per-pixel operation, including the Gaussian noise (Philox4x32-10).
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib
from ._lib import check, ptr, stream_ptr


def _u8_images(images: torch.Tensor) -> torch.Tensor:
    if images.dtype != torch.uint8 or images.dim() != 4 or tuple(images.shape[1:]) != (64, 64, 3):
        raise RuntimeError(f"ae_b200: expected uint8 images of shape [N,64,64,3] (HWC), got {images.dtype} {tuple(images.shape)}")
    if not images.is_cuda:
        raise RuntimeError("ae_b200: the image store must live on a CUDA device (no CPU fallback)")
    return images.contiguous()


def augment_u8(images: torch.Tensor, index: Optional[torch.Tensor] = None, flip: Optional[torch.Tensor] = None,
               off_y: Optional[torch.Tensor] = None, off_x: Optional[torch.Tensor] = None, pad: int = 4,
               noise: Optional[torch.Tensor] = None, noise_std: float = 0.0, noise_mean: float = 0.0, seed: int = 0,
               out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Thin wrapper of ``ae_augment_u8`` (include/ae_b200.h): every random draw is an argument."""
    images = _u8_images(images)
    dev = images.device
    b = int(index.numel()) if index is not None else int(images.shape[0])
    if out is None:
        out = torch.empty(b, 3, 64, 64, dtype=torch.float32, device=dev)
    if tuple(out.shape) != (b, 3, 64, 64) or out.dtype != torch.float32 or not out.is_contiguous() or out.device != dev:
        raise RuntimeError("ae_b200: augment output must be a contiguous float32 [B,3,64,64] tensor on the images' device")

    def arg(t, dtype, what):
        if t is None:
            return None
        t = t.to(device=dev, dtype=dtype).contiguous()
        if t.numel() != b * (12288 if what == "noise" else 1):
            raise RuntimeError(f"ae_b200: augment argument `{what}` has {t.numel()} elements for a batch of {b}")
        return t

    index = arg(index, torch.int64, "index")      # (its bounds are the caller's contract: checking would cost a sync per batch)
    flip, off_y, off_x = arg(flip, torch.uint8, "flip"), arg(off_y, torch.int32, "off_y"), arg(off_x, torch.int32, "off_x")
    noise = arg(noise, torch.float32, "noise")
    check(_lib.load().ae_augment_u8(ptr(images), int(images.shape[0]), ptr(index), ptr(flip), ptr(off_y), ptr(off_x), int(pad),
                                    ptr(noise), int(seed) & (2 ** 64 - 1), float(noise_mean), float(noise_std), ptr(out), b,
                                    stream_ptr()))
    return out


class TrainTransformAE:
    """``train_transform_ae`` of NB:386-391 for a whole batch on the device."""

    def __init__(self, p_flip: float = 0.5, padding: int = 4, noise_mean: float = 0.0, noise_std: float = 0.03,
                 generator: Optional[torch.Generator] = None, seed: Optional[int] = None):
        """generator: device generator of the flip / crop draws (None: torch's default CUDA generator);
        seed: base of the noise streams (None: drawn once from torch's CPU generator)."""
        self.p_flip, self.padding, self.noise_mean, self.noise_std, self.generator = p_flip, padding, noise_mean, noise_std, generator
        self.seed = int(seed) if seed is not None else int(torch.randint(0, 2 ** 62, (1,)).item())
        self._calls = 0

    def draws(self, b: int, device):
        g = self.generator
        flip = (torch.rand(b, device=device, generator=g) < self.p_flip).to(torch.uint8)       # RandomHorizontalFlip
        off = torch.randint(0, 2 * self.padding + 1, (2, b), device=device, generator=g, dtype=torch.int32)   # RandomCrop (i, j)
        seed = (self.seed + self._calls * 0x9E3779B97F4A7C15) & (2 ** 64 - 1)     # a fresh noise stream per batch, no host sync
        self._calls += 1
        return flip, off[0], off[1], seed

    def __call__(self, images: torch.Tensor, index: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None):
        b = int(index.numel()) if index is not None else int(images.shape[0])
        flip, oy, ox, seed = self.draws(b, images.device)
        return augment_u8(images, index, flip, oy, ox, self.padding, None, self.noise_std, self.noise_mean, seed, out)


class EvalTransform:
    """``test_val_transform`` of NB:393-395 (ToTensor) for a whole batch on the device."""

    def __call__(self, images: torch.Tensor, index: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None):
        return augment_u8(images, index, out=out)


class DeviceDataset:
    """A split held on the device: uint8 HWC images + int64 labels (what ``TransformDataset(subset, transform)``
    NB:333-343 holds on the host)."""

    def __init__(self, images_u8: torch.Tensor, labels: torch.Tensor, device=None):
        dev = torch.device(device) if device is not None else (images_u8.device if images_u8.is_cuda else
                                                                torch.device("cuda", torch.cuda.current_device()))
        self.images = _u8_images(images_u8.to(dev))
        self.labels = labels.to(device=dev, dtype=torch.int64).contiguous()
        if self.labels.numel() != self.images.shape[0]:
            raise RuntimeError("ae_b200: one label per image expected")

    def __len__(self):
        return int(self.images.shape[0])


class DeviceLoader:
    """``DataLoader(TransformDataset(subset, transform), batch_size, shuffle)`` (NB:420-422) whose batches are produced on
    the device.  Like the reference's loaders it keeps the last, smaller batch."""

    def __init__(self, dataset: DeviceDataset, batch_size: int = 64, shuffle: bool = False, transform=None,
                 generator: Optional[torch.Generator] = None):
        self.dataset, self.batch_size, self.shuffle, self.generator = dataset, int(batch_size), shuffle, generator
        self.transform = transform if transform is not None else EvalTransform()

    def __len__(self):
        return (len(self.dataset) + self.batch_size - 1) // self.batch_size

    def batches(self):
        """Yields the index tensor of every batch of one epoch (device int64)."""
        n, dev = len(self.dataset), self.dataset.images.device
        order = torch.randperm(n, device=dev, generator=self.generator) if self.shuffle else torch.arange(n, device=dev)
        for i in range(0, n, self.batch_size):
            yield order[i:i + self.batch_size]

    def __iter__(self):
        for idx in self.batches():
            yield self.transform(self.dataset.images, idx), self.dataset.labels.index_select(0, idx)
