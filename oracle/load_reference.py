"""Load the reference's own model classes by executing its notebook cells.  TEST INFRASTRUCTURE.

Looks for the notebook under ``$AE_REFERENCE_ROOT``, ``/root/reference`` (this build container) and ``baseline/_ref``
(the git-ignored copy ``__graft_entry__.build()`` places there so that it travels to the GPU box).
Nothing is copied into the product: the four class cells (NB:499-525 Encoder, NB:607-635 Decoder,
NB:685-702 SupervisedAutoencoder, NB:2970-2987 MLP) are read from the .ipynb at run time and
``exec``-ed into a namespace that holds only ``torch`` and ``nn``.
"""
from __future__ import annotations

import glob
import json
import os

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE_ROOTS = [r for r in (os.environ.get("AE_REFERENCE_ROOT"), "/root/reference", os.path.join(_REPO, "baseline", "_ref")) if r]
REFERENCE_ROOT = REFERENCE_ROOTS[0]
_WANTED = ("class Encoder(", "class Decoder(", "class SupervisedAutoencoder(", "class MLP(")


def notebook_path():
    for root in REFERENCE_ROOTS:
        paths = sorted(glob.glob(os.path.join(root, "Code", "*.ipynb")))
        if paths:
            return paths[0]
    return None


def reference_available() -> bool:
    return notebook_path() is not None


def load_reference_classes():
    """Returns a dict with the reference's Encoder, Decoder, SupervisedAutoencoder, MLP classes."""
    import torch
    import torch.nn as nn

    path = notebook_path()
    if path is None:
        raise FileNotFoundError(f"reference notebook not found under any of {REFERENCE_ROOTS}")
    paths = [path]
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
