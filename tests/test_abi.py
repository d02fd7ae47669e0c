"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol declared in
include/ae_b200.h, the host shells expose the reference's surface, and nothing silently falls back."""
import os
import re

import pytest
import torch

import ae_b200
from ae_b200 import _lib
from oracle import seeded

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "ae_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ae_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _header_functions()
    assert len(names) >= 50
    for n in names:
        assert hasattr(lib, n), f"libae_b200.so does not export {n}"
    assert set(names) == set(_lib.declared_symbols()), set(names) ^ set(_lib.declared_symbols())
    assert lib.ae_abi_version() == 1


def test_layout_queries_match_reference_shapes():
    import ctypes as C
    lib = _lib.load()
    cfg = _lib.EngineConfig(64, 10, 8, _lib.PREC_FP32, _lib.BACKEND_SIMT)
    h = C.c_void_p()
    _lib.check(lib.ae_engine_create(C.byref(cfg), C.byref(h)))
    shapes = seeded.ae_state_shapes(64, 10)
    pkeys = [k for k in shapes if k.rsplit(".", 1)[1] in ("weight", "bias")]
    sizes_ref = [int(torch.Size(shapes[k]).numel()) for k in pkeys]
    got = []
    total = 0
    for part in range(3):
        offs, sizes, flen = (C.c_int64 * 64)(), (C.c_int64 * 64)(), C.c_int64()
        n = lib.ae_engine_param_layout(h, part, offs, sizes, C.byref(flen))
        got += [sizes[i] for i in range(n)]
        assert all(offs[i] % 4 == 0 for i in range(n))
        total += flen.value
    assert got == sizes_ref
    assert sum(got) == 1316045
    assert total >= 1316045 and total % 4 == 0
    assert lib.ae_engine_workspace_bytes(h) > 0
    lib.ae_engine_destroy(h)
    offs, sizes = (C.c_int64 * 10)(), (C.c_int64 * 10)()
    lib.ae_mlp_param_layout(64, 10, offs, sizes)
    assert sum(sizes) == 17610


def test_bad_arguments_report_errors_without_a_gpu():
    import ctypes as C
    lib = _lib.load()
    cfg = _lib.EngineConfig(60, 10, 8, 0, 0)
    h = C.c_void_p()
    assert lib.ae_engine_create(C.byref(cfg), C.byref(h)) != 0
    assert b"latent_dim" in lib.ae_last_error()


def test_module_surface_matches_reference():
    m = ae_b200.SupervisedAutoencoder(latent_dim=64, num_classes=10)
    assert list(m.state_dict().keys()) == list(seeded.ae_state_shapes(64, 10).keys())
    for k, v in m.state_dict().items():
        assert tuple(v.shape) == tuple(seeded.ae_state_shapes(64, 10)[k]), k
    assert hasattr(m, "enc") and hasattr(m, "dec") and hasattr(m, "classifier")
    assert hasattr(m.enc, "encoder") and hasattr(m.dec, "decoder_input") and hasattr(m.dec, "decoder")
    m128 = ae_b200.SupervisedAutoencoder(latent_dim=128, num_classes=10)      # NB:807
    assert m128.enc.encoder[13].weight.shape == (128, 4096)
    clf = ae_b200.MLP(input_dim=64, num_classes=10)
    assert list(clf.state_dict().keys()) == list(seeded.mlp_state_shapes(64, 10).keys())
    m.load_state_dict(seeded.seeded_state(seeded.ae_state_shapes(64, 10), 1))
    m.train(); m.eval()
    for p in m.enc.parameters():
        p.requires_grad = False
    assert sum(p.numel() for p in m.parameters()) == 1316045


def test_forward_on_cpu_raises_instead_of_falling_back():
    m = ae_b200.SupervisedAutoencoder(64)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 3, 64, 64))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ae_b200.MLP(64)(torch.zeros(2, 64))


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "hybrid-autoencoder-mlp-pipeline-for-satellite-image-classification_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(root, f), errors="ignore").read()
                assert "oracle" not in txt.replace("no oracle", ""), f


def test_data_path_on_cpu_raises_instead_of_falling_back():
    u8 = torch.zeros(4, 64, 64, 3, dtype=torch.uint8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ae_b200.augment_u8(u8)
    with pytest.raises(RuntimeError, match="uint8 images of shape"):
        ae_b200.augment_u8(torch.zeros(4, 3, 64, 64))
    # host-side pieces of the epoch statistics (NB:2686-2690) need no device
    assert ae_b200.fit.weighted_mean(torch.tensor([[2.0, 0, 0], [4.0, 0, 0]]), [3, 1]) == pytest.approx(2.5)
