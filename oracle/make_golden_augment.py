"""Generate tests/golden/augment_*.npz by running the REFERENCE's own transform cells.  TEST INFRASTRUCTURE.

Run in the build container (needs /root/reference, torchvision, PIL):  python -m oracle.make_golden_augment
The cells defining AddGaussianNoise (NB:361-368) and the two Compose pipelines (NB:386-395) are executed from the
notebook (not copied); each synthetic uint8 image goes through ``train_transform_ae`` as a PIL image after
``torch.manual_seed(seed_i)``, exactly as ``TransformDataset.__getitem__`` (NB:340-343) would call it.
"""
from __future__ import annotations

import glob
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import augment_port as ap  # noqa: E402
from oracle.load_reference import REFERENCE_ROOT  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def load_reference_transforms():
    import torchvision.transforms as transforms
    path = glob.glob(os.path.join(REFERENCE_ROOT, "Code", "*.ipynb"))[0]
    nb = json.load(open(path, encoding="utf-8"))
    ns = {"torch": torch, "transforms": transforms}
    for cell in nb["cells"]:
        if cell.get("cell_type") != "code":
            continue
        src = "".join(cell["source"])
        if src.lstrip().startswith("class AddGaussianNoise"):
            exec(compile(src, path, "exec"), ns)  # noqa: S102 - reference code, executed not copied
        elif "train_transform_ae = transforms.Compose" in src:
            # only the two Compose definitions; the TransformDataset lines below them need the dataset
            head = src.split("trainset_ae")[0]
            exec(compile(head, path, "exec"), ns)  # noqa: S102
    return ns["train_transform_ae"], ns["test_val_transform"]


def main():
    from PIL import Image
    train_tf, eval_tf = load_reference_transforms()
    n = 6
    imgs = ap.synthetic_images(n, seed=11)
    seeds = np.arange(100, 100 + n)
    outs, evals, flips, oys, oxs, noises = [], [], [], [], [], []
    for i in range(n):
        pil = Image.fromarray(imgs[i], mode="RGB")
        torch.manual_seed(int(seeds[i]))
        outs.append(train_tf(pil).numpy())
        evals.append(eval_tf(pil).numpy())
        f, oy, ox, nz = ap.draws_like_reference(int(seeds[i]))
        flips.append(f); oys.append(oy); oxs.append(ox); noises.append(nz.numpy())
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, f"augment_B{n}.npz"), images=imgs, seeds=seeds, train_out=np.stack(outs),
                        eval_out=np.stack(evals), flip=np.array(flips, dtype=np.uint8), off_y=np.array(oys, dtype=np.int32),
                        off_x=np.array(oxs, dtype=np.int32), noise=np.stack(noises))
    print("wrote augment fixture:", n, "images; flips", flips, "offsets", list(zip(oys, oxs)))


if __name__ == "__main__":
    main()
