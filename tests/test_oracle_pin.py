"""Pin the oracle (oracle/torch_port.py) to the reference's own outputs (tests/golden/*.npz).

The fixtures were produced by executing the reference notebook's classes and loop body
(oracle/make_golden.py).  CPU-only; runs everywhere.
"""
import numpy as np
import pytest
import torch

from oracle import seeded, torch_port as tp
from tests import golden_util as gu

TOL = 2e-6   # same ATen kernels as the reference; fixtures were generated single-threaded


@pytest.fixture(autouse=True)
def _single_thread():
    """The fixtures were generated with torch.set_num_threads(1); ATen's CPU reductions are
    deterministic for a fixed thread count.  This matters for the pre-BatchNorm conv/linear
    biases: their true gradient is exactly 0, torch's is ~1e-7 rounding noise, and Adam turns
    that noise into +-lr updates ('noise-driven parameters')."""
    n = torch.get_num_threads()
    torch.set_num_threads(1)
    yield
    torch.set_num_threads(n)


@pytest.mark.parametrize("tag", ["ae_eval_L64_B4", "ae_eval_L128_B3"])
def test_ae_eval_forward(tag):
    g = gu.load(tag)
    latent, batch, seed = [int(v) for v in g["meta"]]
    st = seeded.seeded_state(seeded.ae_state_shapes(latent, 10), seed)
    x = seeded.seeded_images(batch, seed)
    with torch.no_grad():
        x_hat, logits, z = tp.ae_forward(st, x, False)
    gu.check(g, "x_hat", x_hat.numpy(), TOL)
    gu.check(g, "logits", logits.numpy(), TOL)
    gu.check(g, "z", z.numpy(), TOL)


@pytest.mark.parametrize("tag", ["ae_train_L64_B6", "ae_train_L64_B33"])
def test_ae_train_steps(tag):
    g = gu.load(tag)
    latent, batch, seed, steps = [int(v) for v in g["meta"]]
    alpha, lr = [float(v) for v in g["hyper"]]
    st = seeded.seeded_state(seeded.ae_state_shapes(latent, 10), seed)
    opt = {}
    for s in range(steps):
        x = seeded.seeded_images(batch, seed + 10 * s)
        y = seeded.seeded_labels(batch, seed + 10 * s)
        loss, lrec, lcls, grads, (x_hat, logits, z) = tp.ae_train_step(st, opt, x, y, alpha, lr)
        np.testing.assert_allclose([loss.item(), lrec.item(), lcls.item()], g[f"s{s}/loss"], rtol=1e-6)
        gu.check(g, f"s{s}/x_hat", x_hat.numpy(), TOL)
        gu.check(g, f"s{s}/logits", logits.numpy(), TOL)
        gu.check(g, f"s{s}/z", z.numpy(), TOL)
        for k, gr in grads.items():
            rt, at = gu.grad_tolerances(k, 1e-5, 1.0)
            gu.check(g, f"s{s}/grad/{k}", gr.numpy(), rt, atol=at)
        for k, v in st.items():
            rt, at = gu.state_tolerances(k, 1e-5, lr, s)
            gu.check(g, f"s{s}/state/{k}", v.numpy(), rt, atol=at)


def test_ae_probe_train_mode_forward():
    g = gu.load("ae_probe_L128_B5")
    latent, batch, seed = [int(v) for v in g["meta"]]
    st = seeded.seeded_state(seeded.ae_state_shapes(latent, 10), seed)
    x = seeded.seeded_images(batch, seed)
    y = seeded.seeded_labels(batch, seed)
    nb = {}
    with torch.no_grad():
        x_hat, logits, _ = tp.ae_forward(st, x, True, nb)
        _, lrec, lcls = tp.ae_loss(x_hat, logits, x, y, 1.0)
    np.testing.assert_allclose([lrec.item(), lcls.item()], g["loss"], rtol=1e-6)
    gu.check(g, "logits", logits.numpy(), TOL)
    for k, v in nb.items():
        gu.check(g, "state/" + k, v.numpy(), TOL)


def test_mlp_train_and_eval():
    g = gu.load("mlp_train_B16")
    batch, seed, steps = [int(v) for v in g["meta"]]
    lr, wd = [float(v) for v in g["hyper"]]
    st = seeded.seeded_state(seeded.mlp_state_shapes(64, 10), seed)
    opt = {}
    for s in range(steps):
        x = torch.from_numpy(g[f"s{s}/x"])
        keep = torch.from_numpy(g[f"s{s}/keep"].astype(np.float32))
        y = seeded.seeded_labels(batch, seed + s)
        loss, grads, logits = tp.mlp_train_step(st, opt, x, y, lr, wd, keep)
        np.testing.assert_allclose(loss.item(), g[f"s{s}/loss"][0], rtol=1e-6)
        gu.check(g, f"s{s}/logits", logits.numpy(), TOL)
        for k, gr in grads.items():
            rt, at = gu.grad_tolerances(k, 1e-5, 1.0)
            gu.check(g, f"s{s}/grad/{k}", gr.numpy(), rt, atol=at)
        for k, v in st.items():
            rt, at = gu.state_tolerances(k, 1e-5, lr, s)
            gu.check(g, f"s{s}/state/{k}", v.numpy(), rt, atol=at)
    with torch.no_grad():
        le = tp.mlp_forward(st, torch.from_numpy(g["eval/x"]), False)
    gu.check(g, "eval/logits", le.numpy(), 1e-5)
    assert np.array_equal(le.argmax(1).numpy(), g["eval/argmax"])


def test_encode_predict():
    g = gu.load("encode_predict_B8")
    batch, seed = [int(v) for v in g["meta"]]
    ae = seeded.seeded_state(seeded.ae_state_shapes(64, 10), seed)
    mlp = seeded.seeded_state(seeded.mlp_state_shapes(64, 10), seed + 1)
    z, logits = tp.encode_predict(ae, mlp, seeded.seeded_images(batch, seed))
    gu.check(g, "z", z.numpy(), TOL)
    gu.check(g, "logits", logits.numpy(), TOL)
    assert np.array_equal(logits.argmax(1).numpy(), g["argmax"])


def test_augment_oracle_matches_reference_transforms():
    """oracle/augment_port.py against the outputs of the reference's own torchvision Compose objects (NB:386-395),
    recorded by oracle/make_golden_augment.py: bit-exact, including the order of the random draws."""
    import os
    from oracle import augment_port as ap
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "augment_B6.npz"))
    assert len(set(g["flip"].tolist())) == 2, "fixture must cover flipped and unflipped images"
    for i in range(g["images"].shape[0]):
        flip, oy, ox, noise = ap.draws_like_reference(int(g["seeds"][i]))
        assert (flip, oy, ox) == (bool(g["flip"][i]), int(g["off_y"][i]), int(g["off_x"][i]))
        assert np.array_equal(noise.numpy(), g["noise"][i])
        out = ap.train_transform(g["images"][i], flip, oy, ox, noise)
        assert np.array_equal(out.numpy(), g["train_out"][i])
        assert np.array_equal(ap.to_tensor(g["images"][i]).numpy(), g["eval_out"][i])
