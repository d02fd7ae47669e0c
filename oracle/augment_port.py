"""CPU restatement of the reference's input transforms.  TEST INFRASTRUCTURE ONLY.

Restates, on uint8 HWC arrays, what the reference's torchvision pipeline does to one PIL image:

  train_transform_ae (NB:386-391): RandomHorizontalFlip() -> RandomCrop(64, padding=4) -> ToTensor()
                                   -> AddGaussianNoise(0., 0.03)   (class at NB:361-368)
  test_val_transform (NB:393-395): ToTensor()

The arithmetic lives in third-party torchvision (0.26 in this image; the reference pins no version):
  RandomHorizontalFlip.forward : ``if torch.rand(1) < p: img = hflip(img)``
  RandomCrop.forward           : zero-pad `padding` on every side, then ``i = randint(0, h - th + 1)``,
                                 ``j = randint(0, w - tw + 1)`` and crop rows i.., columns j..
  ToTensor                     : HWC uint8 -> CHW float32, ``.div(255)``
  AddGaussianNoise.__call__    : ``tensor + (torch.randn_like(tensor) * std + mean)``
The random draws are arguments here (flip bit, crop offsets, noise tensor) so that the result is a
pure function; ``draws_like_reference`` reproduces the reference's draw ORDER from a seeded torch RNG.
Pinned against the reference's own Compose objects by ``oracle/make_golden_augment.py`` +
``tests/test_oracle_pin.py``.
"""
from __future__ import annotations

import numpy as np
import torch

PAD = 4            # NB:388
NOISE_STD = 0.03   # NB:390
NOISE_MEAN = 0.0


def to_tensor(img_u8: np.ndarray) -> torch.Tensor:
    """ToTensor (NB:389, NB:394): [H,W,3] uint8 -> [3,H,W] float32 in [0,1]."""
    t = torch.from_numpy(np.ascontiguousarray(img_u8)).permute(2, 0, 1).contiguous()
    return t.to(torch.float32).div(255)


def flip_crop(img_u8: np.ndarray, flip: bool, off_y: int, off_x: int, pad: int = PAD) -> np.ndarray:
    """RandomHorizontalFlip + RandomCrop(64, padding=pad) with the draws given (NB:387-388)."""
    h, w, _ = img_u8.shape
    a = img_u8[:, ::-1, :] if flip else img_u8
    p = np.zeros((h + 2 * pad, w + 2 * pad, 3), dtype=np.uint8)
    p[pad:pad + h, pad:pad + w] = a
    return p[off_y:off_y + h, off_x:off_x + w]


def train_transform(img_u8, flip, off_y, off_x, noise, std: float = NOISE_STD, mean: float = NOISE_MEAN, pad: int = PAD):
    """NB:386-391 on one image; `noise` is the standard-normal draw of AddGaussianNoise ([3,H,W] fp32) or None."""
    t = to_tensor(flip_crop(img_u8, bool(flip), int(off_y), int(off_x), pad))
    if noise is None:
        return t
    return t + (noise * std + mean)


def draws_like_reference(gen_seed: int, shape=(3, 64, 64), pad: int = PAD):
    """The draws one call of train_transform_ae makes after ``torch.manual_seed(gen_seed)``, in the reference's order:
    torch.rand(1) (flip), torch.randint (row offset), torch.randint (column offset), torch.randn_like (noise)."""
    torch.manual_seed(gen_seed)
    flip = bool(torch.rand(1) < 0.5)
    off_y = int(torch.randint(0, 2 * pad + 1, size=(1,)).item())
    off_x = int(torch.randint(0, 2 * pad + 1, size=(1,)).item())
    noise = torch.randn(shape)
    return flip, off_y, off_x, noise


def synthetic_images(n: int, seed: int) -> np.ndarray:
    """EuroSAT-shaped uint8 RGB patches [n,64,64,3] (smooth pattern + texture, not constant)."""
    rs = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:64, 0:64].astype(np.float32)
    out = np.empty((n, 64, 64, 3), dtype=np.uint8)
    for i in range(n):
        base = rs.uniform(40, 200, size=3)
        fx, fy, ph = rs.uniform(0.02, 0.3, size=3), rs.uniform(0.02, 0.3, size=3), rs.uniform(0, 6.28, size=3)
        img = base[None, None, :] + 40 * np.sin(xx[..., None] * fx + yy[..., None] * fy + ph) + rs.normal(0, 6, size=(64, 64, 3))
        out[i] = np.clip(img, 0, 255).astype(np.uint8)
    return out
