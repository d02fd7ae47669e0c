"""Load the reference's own model classes by executing its notebook cells.  TEST INFRASTRUCTURE.

Works only where ``/root/reference`` is mounted (this build container, not the GPU box).
Nothing is copied: the four class cells (NB:499-525 Encoder, NB:607-635 Decoder,
NB:685-702 SupervisedAutoencoder, NB:2970-2987 MLP) are read from the .ipynb at run time and
``exec``-ed into a namespace that holds only ``torch`` and ``nn``.
"""
from __future__ import annotations

import glob
import json
import os

REFERENCE_ROOT = os.environ.get("AE_REFERENCE_ROOT", "/root/reference")
_WANTED = ("class Encoder(", "class Decoder(", "class SupervisedAutoencoder(", "class MLP(")


def reference_available() -> bool:
    return bool(glob.glob(os.path.join(REFERENCE_ROOT, "Code", "*.ipynb")))


def load_reference_classes():
    """Returns a dict with the reference's Encoder, Decoder, SupervisedAutoencoder, MLP classes."""
    import torch
    import torch.nn as nn

    paths = glob.glob(os.path.join(REFERENCE_ROOT, "Code", "*.ipynb"))
    if not paths:
        raise FileNotFoundError(f"reference notebook not found under {REFERENCE_ROOT}")
    with open(paths[0], "r", encoding="utf-8") as f:
        nb = json.load(f)
    ns = {"torch": torch, "nn": nn}
    found = set()
    for cell in nb["cells"]:
        if cell.get("cell_type") != "code":
            continue
        src = "".join(cell["source"])
        hits = [w for w in _WANTED if w in src]
        if hits and src.lstrip().startswith("class "):
            exec(compile(src, paths[0], "exec"), ns)  # noqa: S102 - reference code, executed not copied
            found.update(hits)
    missing = set(_WANTED) - found
    if missing:
        raise RuntimeError(f"reference cells not found: {sorted(missing)}")
    return {k: ns[k] for k in ("Encoder", "Decoder", "SupervisedAutoencoder", "MLP")}
