"""Generate tests/golden/*.npz by running the REFERENCE's own classes.  TEST INFRASTRUCTURE.

Run in the build container (needs /root/reference):  python -m oracle.make_golden
The reference modules are built from the notebook cells (oracle/load_reference.py), their
parameters are overwritten with oracle.seeded values, and the literal loop body of
NB:2676-2684 / NB:3477-3482 is executed with torch.optim.Adam exactly as NB:2654 / NB:3461
construct it.  Large tensors are stored as a strided sample (every STRIDE-th element of
the flattened tensor) plus fp64 sum / abs-sum, so the fixtures stay small.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import seeded  # noqa: E402
from oracle.load_reference import load_reference_classes  # noqa: E402

STRIDE = 61
FULL_LIMIT = 8192
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def pack(name, t, out):
    a = t.detach().cpu().numpy().copy()   # copy: parameters are updated in place by later steps
    flat = a.reshape(-1)
    if flat.size <= FULL_LIMIT:
        out[name] = a
    else:
        out[name + "@sample"] = flat[::STRIDE].copy()
        out[name + "@sums"] = np.array([flat.astype(np.float64).sum(), np.abs(flat.astype(np.float64)).sum()])


def load_seeded(model, shapes, seed):
    st = seeded.seeded_state(shapes, seed)
    missing, unexpected = model.load_state_dict(st, strict=True)
    assert not missing and not unexpected
    return st


def golden_ae_eval(cls, latent, batch, seed, tag):
    torch.manual_seed(0)
    model = cls["SupervisedAutoencoder"](latent_dim=latent, num_classes=10)
    load_seeded(model, seeded.ae_state_shapes(latent, 10), seed)
    model.eval()
    x = seeded.seeded_images(batch, seed)
    with torch.no_grad():
        x_hat, logits, z = model(x)
        z_enc = model.enc(x)
    out = {"meta": np.array([latent, batch, seed])}
    pack("x_hat", x_hat, out)
    pack("logits", logits, out)
    pack("z", z, out)
    assert torch.equal(z, z_enc)
    np.savez_compressed(os.path.join(OUT, tag + ".npz"), **out)


def golden_ae_train(cls, latent, batch, seed, tag, steps=2, alpha=35.0, lr=5e-3):
    torch.manual_seed(0)
    model = cls["SupervisedAutoencoder"](latent_dim=latent, num_classes=10)
    load_seeded(model, seeded.ae_state_shapes(latent, 10), seed)
    criterion_recon = nn.MSELoss()                 # NB:2652
    criterion_class = nn.CrossEntropyLoss()        # NB:2653
    optimizer = torch.optim.Adam(model.parameters(), lr=lr)   # NB:2654
    model.train()                                  # NB:2668
    out = {"meta": np.array([latent, batch, seed, steps]), "hyper": np.array([alpha, lr])}
    for s in range(steps):
        imgs = seeded.seeded_images(batch, seed + 10 * s)
        labels = seeded.seeded_labels(batch, seed + 10 * s)
        optimizer.zero_grad()                      # NB:2676
        x_hat, logits, z = model(imgs)             # NB:2677
        loss_recon = criterion_recon(x_hat, imgs)  # NB:2679
        loss_class = criterion_class(logits, labels)   # NB:2680
        loss = alpha * loss_recon + loss_class     # NB:2681
        loss.backward()                            # NB:2683
        grads = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
        optimizer.step()                           # NB:2684
        out[f"s{s}/loss"] = np.array([loss.item(), loss_recon.item(), loss_class.item()])
        pack(f"s{s}/x_hat", x_hat, out)
        pack(f"s{s}/logits", logits, out)
        pack(f"s{s}/z", z, out)
        for k, g in grads.items():
            pack(f"s{s}/grad/{k}", g, out)
        for k, v in model.state_dict().items():
            pack(f"s{s}/state/{k}", v, out)
    np.savez_compressed(os.path.join(OUT, tag + ".npz"), **out)


def golden_ae_probe(cls, latent, batch, seed, tag):
    """Init-time loss-ratio probe, NB:804-821: one no-grad forward in train() mode at latent_dim=128."""
    torch.manual_seed(0)
    model = cls["SupervisedAutoencoder"](latent_dim=latent, num_classes=10)
    load_seeded(model, seeded.ae_state_shapes(latent, 10), seed)
    imgs = seeded.seeded_images(batch, seed)
    labels = seeded.seeded_labels(batch, seed)
    with torch.no_grad():
        x_hat, logits, _ = model(imgs)
        lr_ = nn.MSELoss()(x_hat, imgs)
        lc_ = nn.CrossEntropyLoss()(logits, labels)
    out = {"meta": np.array([latent, batch, seed]), "loss": np.array([lr_.item(), lc_.item()])}
    pack("logits", logits, out)
    pack("x_hat", x_hat, out)
    for k, v in model.state_dict().items():
        if "running" in k or "num_batches" in k:
            pack("state/" + k, v, out)
    np.savez_compressed(os.path.join(OUT, tag + ".npz"), **out)


def golden_mlp(cls, batch, seed, tag, steps=2, lr=1e-3):
    torch.manual_seed(1234)
    clf = cls["MLP"](input_dim=64, num_classes=10)
    load_seeded(clf, seeded.mlp_state_shapes(64, 10), seed)
    optimizer = torch.optim.Adam(clf.parameters(), lr=lr, weight_decay=1e-4)   # NB:3461
    criterion = torch.nn.CrossEntropyLoss()                                    # NB:3463
    masks = []
    drop = clf.net[3]
    hook = drop.register_forward_hook(lambda m, i, o: masks.append((o != 0) | (i[0] == 0)))
    out = {"meta": np.array([batch, seed, steps]), "hyper": np.array([lr, 1e-4])}
    clf.train()
    for s in range(steps):
        rs = np.random.RandomState(500 + seed + s)
        xb = torch.from_numpy(rs.standard_normal((batch, 64)).astype(np.float32))
        yb = seeded.seeded_labels(batch, seed + s)
        optimizer.zero_grad()
        logits = clf(xb)
        loss = criterion(logits, yb)
        loss.backward()
        grads = {k: p.grad.detach().clone() for k, p in clf.named_parameters()}
        optimizer.step()
        out[f"s{s}/x"] = xb.numpy()
        out[f"s{s}/keep"] = masks[-1].numpy().astype(np.uint8)
        out[f"s{s}/loss"] = np.array([loss.item()])
        pack(f"s{s}/logits", logits, out)
        for k, g in grads.items():
            pack(f"s{s}/grad/{k}", g, out)
        for k, v in clf.state_dict().items():
            pack(f"s{s}/state/{k}", v, out)
    hook.remove()
    clf.eval()
    rs = np.random.RandomState(777 + seed)
    xe = torch.from_numpy(rs.standard_normal((batch + 3, 64)).astype(np.float32))
    with torch.no_grad():
        le = clf(xe)
    out["eval/x"] = xe.numpy()
    pack("eval/logits", le, out)
    out["eval/argmax"] = le.argmax(1).numpy()
    np.savez_compressed(os.path.join(OUT, tag + ".npz"), **out)


def golden_encode_predict(cls, batch, seed, tag):
    """BASELINE config 1: clf(enc(x)) in eval mode with seeded weights."""
    ae = cls["SupervisedAutoencoder"](latent_dim=64, num_classes=10)
    load_seeded(ae, seeded.ae_state_shapes(64, 10), seed)
    clf = cls["MLP"](input_dim=64, num_classes=10)
    load_seeded(clf, seeded.mlp_state_shapes(64, 10), seed + 1)
    ae.eval()
    clf.eval()
    x = seeded.seeded_images(batch, seed)
    with torch.no_grad():
        z = ae.enc(x)
        logits = clf(z)
    out = {"meta": np.array([batch, seed])}
    pack("z", z, out)
    pack("logits", logits, out)
    out["argmax"] = logits.argmax(1).numpy()
    np.savez_compressed(os.path.join(OUT, tag + ".npz"), **out)


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)   # fixed reduction order
    cls = load_reference_classes()
    golden_ae_eval(cls, 64, 4, 1, "ae_eval_L64_B4")
    golden_ae_eval(cls, 128, 3, 2, "ae_eval_L128_B3")
    golden_ae_train(cls, 64, 6, 3, "ae_train_L64_B6")
    golden_ae_train(cls, 64, 1 + 32, 4, "ae_train_L64_B33", steps=1)
    golden_ae_probe(cls, 128, 5, 5, "ae_probe_L128_B5")
    golden_mlp(cls, 16, 6, "mlp_train_B16")
    golden_encode_predict(cls, 8, 7, "encode_predict_B8")
    print("golden vectors written to", OUT)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
