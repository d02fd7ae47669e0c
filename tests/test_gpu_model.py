"""Model-level parity on the GPU: the drop-in modules against the committed golden vectors (produced by the
reference's own notebook cells) and against the CPU oracle on the same seeded inputs.

Tolerances are BASELINE.json's: fp32 rel <= 1e-4 on logits and latents, bf16 rel <= 1e-2, argmax identical;
rel = max|a-b| / max|b|.  Gradients get 10x the forward tolerance (they pass through twice as many layers).
"""
import numpy as np
import pytest
import torch
import torch.nn as nn

import ae_b200
from ae_b200 import _lib
from oracle import seeded, torch_port as tp
from tests import golden_util as gu_gold
from tests import gpu_util as gu

pytestmark = pytest.mark.gpu

CONFIGS = [(b, p) for b in gu.BACKENDS for p in gu.PRECISIONS if not (b == "simt" and p == "bf16")]


def _model(latent, backend, prec):
    return ae_b200.SupervisedAutoencoder(latent, 10, precision=prec, backend=backend)


@pytest.mark.parametrize("backend,prec", CONFIGS)
@pytest.mark.parametrize("tag", ["ae_eval_L64_B4", "ae_eval_L128_B3"])
def test_ae_eval_forward_golden(tag, backend, prec):
    g = gu_gold.load(tag)
    latent, batch, seed = [int(v) for v in g["meta"]]
    m = _model(latent, backend, prec)
    gu.load_ae(m, seed, latent)
    m = m.to(gu.dev()).eval()
    x = seeded.seeded_images(batch, seed).to(gu.dev())
    with torch.no_grad():
        x_hat, logits, z = m(x)
        z2 = m.enc(x)
    tol = gu.TOL[prec]
    gu_gold.check(g, "z", z.cpu().numpy(), tol)
    gu_gold.check(g, "logits", logits.cpu().numpy(), tol)
    gu_gold.check(g, "x_hat", x_hat.cpu().numpy(), tol)
    assert torch.equal(z, z2)


@pytest.mark.parametrize("backend,prec", CONFIGS)
@pytest.mark.parametrize("tag", ["ae_train_L64_B6", "ae_train_L64_B33"])
def test_ae_dropin_training_loop_golden(tag, backend, prec):
    """The reference's loop body NB:2676-2684, verbatim, on the drop-in modules."""
    g = gu_gold.load(tag)
    latent, batch, seed, steps = [int(v) for v in g["meta"]]
    alpha, lr = [float(v) for v in g["hyper"]]
    model = _model(latent, backend, prec)
    gu.load_ae(model, seed, latent)
    model = model.to(gu.dev())
    criterion_recon = nn.MSELoss()
    criterion_class = nn.CrossEntropyLoss()
    model.train()
    model(seeded.seeded_images(2, 99).to(gu.dev()))      # materialise flat storage (ae_b200.Adam needs it) ...
    gu.load_ae(model, seed, latent)                      # ... then restore the BN buffers that forward touched
    optimizer = ae_b200.Adam(model.parameters(), lr=lr)
    tol = gu.TOL[prec]
    for s in range(steps):
        imgs = seeded.seeded_images(batch, seed + 10 * s).to(gu.dev())
        labels = seeded.seeded_labels(batch, seed + 10 * s).to(gu.dev())
        optimizer.zero_grad()
        x_hat, logits, z = model(imgs)
        loss_recon = criterion_recon(x_hat, imgs)
        loss_class = criterion_class(logits, labels)
        loss = alpha * loss_recon + loss_class
        loss.backward()
        grads = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
        optimizer.step()
        ref_loss = g[f"s{s}/loss"]
        assert abs(loss.item() - ref_loss[0]) <= tol * 10 * abs(ref_loss[0])
        # after the first optimizer step the comparison is conditioned by Adam, not by the kernels: its first step is
        # lr * sign(g), so an entry whose gradient is ~0 moves by +-lr in either arithmetic
        later = max(tol * 20, 1e-2)
        gu_gold.check(g, f"s{s}/z", z.detach().cpu().numpy(), tol if s == 0 else later)
        gu_gold.check(g, f"s{s}/logits", logits.detach().cpu().numpy(), tol if s == 0 else later)
        gu_gold.check(g, f"s{s}/x_hat", x_hat.detach().cpu().numpy(), tol if s == 0 else later)
        if s == 0:
            for k, gr in grads.items():
                # max-norm on sampled entries: bounded loosely (single ReLU-branch flips move single entries)
                rt, at = gu_gold.grad_tolerances(k, max(2e-2, 5 * gu.GRAD_TOL[(backend, prec)]), 1.0)
                gu_gold.check(g, f"s{s}/grad/{k}", gr.cpu().numpy(), rt, atol=at)
            for k, v in model.state_dict().items():
                if k.endswith("running_mean") or k.endswith("running_var") or k.endswith("num_batches_tracked"):
                    rt, at = gu_gold.state_tolerances(k, tol * 10, lr, s)
                    gu_gold.check(g, f"s{s}/state/{k}", v.cpu().numpy(), rt, atol=at)
    sd = model.state_dict()
    assert int(sd["enc.encoder.1.num_batches_tracked"]) == steps
    assert int(sd["dec.decoder.8.num_batches_tracked"]) == steps


@pytest.mark.parametrize("backend,prec", CONFIGS)
@pytest.mark.parametrize("batch", [1, 33, 256])
def test_ae_fused_train_step_vs_oracle(batch, backend, prec):
    """ae_train_step (forward + alpha*MSE + CE + backward in the library) against the oracle's autograd.

    Gradients are compared with the oracle evaluated in fp64, in relative L2 norm.  The gradient is a
    discontinuous function of the activations: a pre-activation within one ulp of zero takes the other ReLU branch
    in another arithmetic, which moves individual weight-gradient entries by ~1/sqrt(pixels) (measured: the fp32
    CPU reference itself differs from fp64 by up to 4e-3 of max|g| on single entries at batch 256, see DESIGN.md
    'gradient conditioning').  The max-norm is therefore only bounded loosely; losses use the fp32 oracle."""
    seed, alpha = 11, 35.0
    st = seeded.seeded_state(seeded.ae_state_shapes(64, 10), seed)
    x, y = seeded.seeded_images(batch, seed), seeded.seeded_labels(batch, seed)
    ref_state = {k: v.clone() for k, v in st.items()}
    loss, lrec, lcls, grads32, (x_hat, logits, z) = tp.ae_train_step(ref_state, {}, x, y, alpha, 5e-3)
    st64 = {k: (v.double() if v.dtype == torch.float32 else v.clone()) for k, v in st.items()}
    _, _, _, grads, _ = tp.ae_train_step(st64, {}, x.double(), y, alpha, 5e-3)
    model = _model(64, backend, prec)
    model.load_state_dict(st)
    model = model.to(gu.dev()).train()
    out = model.train_step_grads(x.to(gu.dev()), y.to(gu.dev()), alpha)
    torch.cuda.synchronize()
    tol = gu.TOL[prec]
    got = out.cpu()
    assert abs(float(got[0]) - float(loss)) <= tol * abs(float(loss))
    assert abs(float(got[1]) - float(lrec)) <= tol * abs(float(lrec))
    assert abs(float(got[2]) - float(lcls)) <= tol * abs(float(lcls))
    worst = ("", 0.0)
    worst32, worst_max, worst_cos = 0.0, 0.0, 1.0
    for k, p in model.named_parameters():
        if k in gu_gold.NOISE_BIAS:
            assert float(p.grad.abs().max()) == 0.0     # mathematically zero; the library writes exact zeros
            continue
        r = gu.rel_l2(p.grad, grads[k])
        worst32 = max(worst32, gu.rel_l2(grads32[k], grads[k]))
        if r > worst[1]:
            worst = (k, r)
        if batch > 1:
            worst_max = max(worst_max, gu.rel(p.grad, grads[k]))
            worst_cos = min(worst_cos, float(torch.nn.functional.cosine_similarity(p.grad.double().cpu().flatten(), grads[k].flatten(), dim=0)))
            assert gu.rel(p.grad, grads[k]) <= gu.GRAD_MAX_TOL[(backend, prec)], k
    print(f"batch {batch} {backend}/{prec}: worst gradient rel-L2 vs fp64 oracle: ours {worst[1]:.2e} ({worst[0]}), fp32 CPU reference {worst32:.2e}; "
          f"worst max-norm {worst_max:.2e}, worst cosine {worst_cos:.5f}")
    if batch > 1:
        assert worst_cos >= gu.GRAD_COS[(backend, prec)], worst_cos
    if batch > 1:      # batch 1: BatchNorm over 16..1024 pixels of one image only -- degenerate conditioning
        # calibrated by the reference arithmetic itself: the fp32 CPU oracle is `worst32` away from the fp64 one
        assert worst[1] <= max(gu.GRAD_TOL[(backend, prec)], 3 * worst32), (worst, worst32)
    if batch > 1:
        for k, v in model.state_dict().items():
            if k.endswith("running_var") or k.endswith("running_mean"):
                assert gu.rel(v, ref_state[k]) <= tol, k


@pytest.mark.parametrize("backend,prec", CONFIGS)
def test_train_step_graph_matches_oracle_over_steps(backend, prec):
    """Whole-step CUDA graph (train step + Adam + re-pack) against the oracle loop.  After ONE step the parameters
    must agree tightly.  Later steps are compared through the loss only: with lr = 5e-3 the first Adam steps move
    every weight by +-lr, the loss jumps from 5.9 to 9.5, and the fp32 and fp64 CPU oracles themselves drift apart
    (25 % of conv1's weights differ by > 0.1 lr after 4 steps; measured, DESIGN.md 'gradient conditioning')."""
    seed, alpha, lr, batch, steps = 21, 35.0, 5e-3, 16, 4
    st = seeded.seeded_state(seeded.ae_state_shapes(64, 10), seed)
    ref_state = {k: v.clone() for k, v in st.items()}
    ref64_state = {k: (v.double() if v.dtype == torch.float32 else v.clone()) for k, v in st.items()}
    opt = {}
    x0, y0 = seeded.seeded_images(batch, seed), seeded.seeded_labels(batch, seed)
    tp.ae_train_step(ref64_state, {}, x0.double(), y0, alpha, lr)      # calibration: the same first step in fp64
    model = _model(64, backend, prec)
    model.load_state_dict(st)
    model = model.to(gu.dev()).train()
    model.engine().prepare(gu.dev(), batch)
    optimizer = ae_b200.Adam(model.parameters(), lr=lr)
    stepper = ae_b200.TrainStep(model, optimizer, alpha, batch)
    assert stepper.num_kernels > 0
    tol = gu.TOL[prec]
    for s in range(steps):
        x, y = seeded.seeded_images(batch, seed + s), seeded.seeded_labels(batch, seed + s)
        loss, *_ = tp.ae_train_step(ref_state, opt, x, y, alpha, lr)
        got = stepper(x.pin_memory(), y.pin_memory())
        assert abs(float(got[0]) - float(loss)) <= max(tol * 50, 2e-3) * abs(float(loss)), (s, float(got[0]), float(loss))
        if s == 0:
            # first Adam step: p -= lr * g / (|g| + eps) ~ lr * sign(g).  Entries whose gradient is ~0 may go either
            # way in ANY arithmetic, so the count of disagreeing entries is calibrated by the CPU oracle itself
            # (fp32 vs fp64 run of the same step) and bounded over all parameters together.
            flips, flips_cal, total = 0, 0, 0
            for k, p in model.named_parameters():
                if k in gu_gold.NOISE_BIAS:
                    continue
                d = (p.detach().cpu() - ref_state[k]).abs()
                assert float(d.max()) <= 2.02 * lr, (k, float(d.max()))
                flips += int((d > 0.1 * lr).sum())
                flips_cal += int(((ref64_state[k].float() - ref_state[k]).abs() > 0.1 * lr).sum())
                total += d.numel()
            frac, cal = flips / total, flips_cal / total
            print(f"first Adam step {backend}/{prec}: {flips}/{total} entries moved the other way ({frac:.2e}); "
                  f"fp32-vs-fp64 oracle: {cal:.2e}")
            assert frac <= 3 * cal + (2e-3 if prec == "fp32" else 5e-2), (frac, cal)


@pytest.mark.parametrize("backend,prec", CONFIGS)
def test_loss_curves_overlap_100_steps(backend, prec):
    """BASELINE north_star: loss curves overlapping over 100 steps.  Class-structured synthetic data, the reference's
    batch size 64 (NB:418) and its smallest learning rate 1e-4 (NB:2630), alpha = 35.  'Overlap' is calibrated by the
    CPU oracle itself: our curve may deviate from the fp32 oracle by at most 3x what the fp64 oracle deviates from it
    (plus 0.5 %), pointwise, and the 10-step moving averages must agree within 1 %."""
    seed, alpha, lr, batch, steps = 5, 35.0, 1e-4, 64, 100
    st = seeded.seeded_state(seeded.ae_state_shapes(64, 10), seed)
    ref_state = {k: v.clone() for k, v in st.items()}
    ref64 = {k: (v.double() if v.dtype == torch.float32 else v.clone()) for k, v in st.items()}
    opt, opt64 = {}, {}
    model = _model(64, backend, prec)
    model.load_state_dict(st)
    model = model.to(gu.dev()).train()
    model.engine().prepare(gu.dev(), batch)
    optimizer = ae_b200.Adam(model.parameters(), lr=lr)
    stepper = ae_b200.TrainStep(model, optimizer, alpha, batch)
    ref_curve, ref64_curve, got_curve = [], [], []
    for s in range(steps):
        y = seeded.seeded_labels(batch, 1000 + s)
        x = seeded.structured_images(y, 1000 + s)
        loss, *_ = tp.ae_train_step(ref_state, opt, x, y, alpha, lr)
        loss64, *_ = tp.ae_train_step(ref64, opt64, x.double(), y, alpha, lr)
        got = stepper(x, y)
        ref_curve.append(float(loss))
        ref64_curve.append(float(loss64))
        got_curve.append(got.clone())
    torch.cuda.synchronize()
    got_curve = np.array([float(v[0]) for v in got_curve])
    ref_curve, ref64_curve = np.array(ref_curve), np.array(ref64_curve)
    assert ref_curve[-10:].mean() < 0.9 * ref_curve[:10].mean()            # the synthetic task actually trains
    dev_ = np.abs(got_curve - ref_curve) / np.abs(ref_curve)
    cal = np.abs(ref64_curve - ref_curve) / np.abs(ref_curve)
    print(f"loss-curve {backend}/{prec}: max dev {dev_.max():.3e} mean {dev_.mean():.3e}; fp64-vs-fp32 oracle max {cal.max():.3e}; "
          f"first {ref_curve[0]:.4f} last {ref_curve[-1]:.4f} ours last {got_curve[-1]:.4f}")
    bound = 3 * cal.max() + (5e-3 if prec == "fp32" else 3e-2)
    assert float(dev_.max()) <= bound, (float(dev_.max()), int(dev_.argmax()), bound)
    ma = lambda a: np.convolve(a, np.ones(10) / 10, mode="valid")
    assert float(np.max(np.abs(ma(got_curve) - ma(ref_curve)) / ma(ref_curve))) <= (1e-2 if prec == "fp32" else 3e-2)


@pytest.mark.parametrize("backend,prec", CONFIGS)
def test_encode_predict_golden_and_argmax(backend, prec):
    g = gu_gold.load("encode_predict_B8")
    batch, seed = [int(v) for v in g["meta"]]
    ae = _model(64, backend, prec)
    gu.load_ae(ae, seed)
    clf = ae_b200.MLP(64, 10)
    gu.load_mlp(clf, seed + 1)
    ae, clf = ae.to(gu.dev()), clf.to(gu.dev())
    x = seeded.seeded_images(batch, seed).to(gu.dev())
    z, logits, am = ae_b200.encode_predict(ae.enc, clf, x)
    tol = gu.TOL[prec]
    gu_gold.check(g, "z", z.cpu().numpy(), tol)
    gu_gold.check(g, "logits", logits.cpu().numpy(), tol)
    assert np.array_equal(am.cpu().numpy(), g["argmax"])
    clf.eval()
    with torch.no_grad():
        assert torch.equal(clf(z), logits)


def test_argmax_identical_on_structured_data_after_training():
    """Class-structured data + a short training run on the GPU, then clf(enc(x)).argmax(1) must equal the oracle's on
    the SAME trained weights (SURVEY hard-part: argmax parity is meaningless at random init)."""
    seed, alpha, lr, batch = 8, 35.0, 2e-3, 64
    st = seeded.seeded_state(seeded.ae_state_shapes(64, 10), seed)
    model = ae_b200.SupervisedAutoencoder(64, 10, backend=gu.BACKENDS[-1])
    model.load_state_dict(st)
    model = model.to(gu.dev()).train()
    model.engine().prepare(gu.dev(), batch)
    optimizer = ae_b200.Adam(model.parameters(), lr=lr)
    stepper = ae_b200.TrainStep(model, optimizer, alpha, batch)
    first = None
    for s in range(40):
        y = seeded.seeded_labels(batch, 2000 + s)
        x = seeded.structured_images(y, 2000 + s)
        out = stepper(x, y)
        if first is None:
            first = out.clone()
    torch.cuda.synchronize()
    assert float(out[2]) < float(first[2])               # cross-entropy went down: the classes are learnable
    # a classifier on the latents: train the MLP for a few steps on the GPU too
    clf = ae_b200.MLP(64, 10).to(gu.dev())
    gu.load_mlp(clf, seed + 1)
    clf = clf.to(gu.dev()).train()
    model.eval()
    clf(torch.zeros(4, 64, device=gu.dev()))
    gu.load_mlp(clf, seed + 1)
    copt = ae_b200.Adam(clf.parameters(), lr=1e-2, weight_decay=1e-4)
    for s in range(60):
        y = seeded.seeded_labels(256, 2500 + s)
        x = seeded.structured_images(y, 2500 + s).to(gu.dev())
        with torch.no_grad():
            zt = model.enc(x)
        copt.zero_grad()
        clf.fused_step_grads(zt, y.to(gu.dev()))
        copt.step()
    y = seeded.seeded_labels(512, 3000)
    x = seeded.structured_images(y, 3000)
    ref_ae = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    ref_clf = {k: v.detach().cpu().clone() for k, v in clf.state_dict().items()}
    zr, lr_ = tp.encode_predict(ref_ae, ref_clf, x)
    z, logits, am = ae_b200.encode_predict(model.enc, clf, x.to(gu.dev()))
    assert gu.rel(z, zr) <= 1e-4 and gu.rel(logits, lr_) <= 1e-4
    ref_am = lr_.argmax(1)
    assert len(set(ref_am.tolist())) >= 5, "degenerate predictions: the test would not discriminate"
    acc = float((ref_am == y).float().mean())
    assert acc > 0.5, acc
    top2 = lr_.topk(2, dim=1).values
    safe = (top2[:, 0] - top2[:, 1]) > 1e-4 * float(lr_.abs().max())      # exclude numerical ties
    assert int(safe.sum()) >= 500
    assert torch.equal(am.cpu()[safe], ref_am[safe])


# ------------------------------------------------------------------------------------------------------
# MLP (NB:2970-2987) -- one cluster kernel
# ------------------------------------------------------------------------------------------------------
def test_mlp_dropin_training_loop_golden():
    g = gu_gold.load("mlp_train_B16")
    batch, seed, steps = [int(v) for v in g["meta"]]
    lr, wd = [float(v) for v in g["hyper"]]
    clf = ae_b200.MLP(input_dim=64, num_classes=10)
    gu.load_mlp(clf, seed)
    clf = clf.to(gu.dev())
    clf.eval()
    with torch.no_grad():
        clf(torch.zeros(2, 64, device=gu.dev()))
    optimizer = ae_b200.Adam(clf.parameters(), lr=lr, weight_decay=wd)
    criterion = torch.nn.CrossEntropyLoss()
    clf.train()
    for s in range(steps):
        xb = torch.from_numpy(g[f"s{s}/x"]).to(gu.dev())
        yb = seeded.seeded_labels(batch, seed + s).to(gu.dev())
        clf.set_dropout_keep_mask(torch.from_numpy(g[f"s{s}/keep"]))
        optimizer.zero_grad()
        logits = clf(xb)
        loss = criterion(logits, yb)
        loss.backward()
        grads = {k: p.grad.detach().clone() for k, p in clf.named_parameters()}
        optimizer.step()
        assert abs(loss.item() - g[f"s{s}/loss"][0]) <= 1e-5 * abs(g[f"s{s}/loss"][0])
        gu_gold.check(g, f"s{s}/logits", logits.detach().cpu().numpy(), 1e-4)
        for k, gr in grads.items():
            rt, at = gu_gold.grad_tolerances(k, 1e-3, 1.0)
            gu_gold.check(g, f"s{s}/grad/{k}", gr.cpu().numpy(), rt, atol=at + 1e-7)
        for k, v in clf.state_dict().items():
            rt, at = gu_gold.state_tolerances(k, 1e-3, lr, s)
            gu_gold.check(g, f"s{s}/state/{k}", v.cpu().numpy(), rt, atol=at + 1e-6)
    clf.set_dropout_keep_mask(None)
    clf.eval()
    with torch.no_grad():
        le = clf(torch.from_numpy(g["eval/x"]).to(gu.dev()))
    gu_gold.check(g, "eval/logits", le.cpu().numpy(), 1e-3)
    assert np.array_equal(le.argmax(1).cpu().numpy(), g["eval/argmax"])


@pytest.mark.parametrize("batch", [1, 5, 64, 256, 300])
def test_mlp_fused_step_vs_oracle(batch):
    seed = 31
    st = seeded.seeded_state(seeded.mlp_state_shapes(64, 10), seed)
    rs = np.random.RandomState(batch)
    x = torch.from_numpy(rs.standard_normal((batch, 64)).astype(np.float32))
    y = seeded.seeded_labels(batch, seed)
    keep = torch.from_numpy((rs.random_sample((batch, 128)) >= 0.3).astype(np.float32))
    ref_state = {k: v.clone() for k, v in st.items()}
    if batch > 1:
        loss, grads, logits = tp.mlp_train_step(ref_state, {}, x, y, 1e-3, 1e-4, keep)
        clf = ae_b200.MLP(64, 10)
        clf.load_state_dict(st)
        clf = clf.to(gu.dev()).train()
        clf.set_dropout_keep_mask(keep)
        gl, gc, glogits = clf.fused_step_grads(x.to(gu.dev()), y.to(gu.dev()))
        torch.cuda.synchronize()
    if batch == 1:
        # BatchNorm over one sample: torch raises in training mode; only check that the kernel runs and is finite
        clf = ae_b200.MLP(64, 10); clf.load_state_dict(st); clf = clf.to(gu.dev()).train()
        gl, gc, glogits = clf.fused_step_grads(x.to(gu.dev()), y.to(gu.dev()))
        assert torch.isfinite(gl).all() and torch.isfinite(glogits).all()
        return
    assert abs(float(gl) - float(loss)) <= 1e-5 * max(1.0, abs(float(loss)))
    assert gu.rel(glogits, logits) <= 1e-4
    assert int(gc) == int((logits.argmax(1) == y).sum())
    for k, p in clf.named_parameters():
        if k in gu_gold.NOISE_BIAS:
            assert float(p.grad.abs().max()) == 0.0
            continue
        assert gu.rel(p.grad, grads[k]) <= 1e-3, k
    for k, v in clf.state_dict().items():
        if "running" in k:
            assert gu.rel(v, ref_state[k]) <= 1e-5, k


def _mlp_keep_mask(clf, batch):
    """The keep mask [B,128] the last training launch drew and saved for its backward half: the last region of the MLP
    workspace (mlp_carve in csrc/mlp.cu: ... bnc, keep[B*128], 256 bytes of slack)."""
    import ctypes as C
    st = clf._state
    total = _lib.load().ae_mlp_workspace_bytes(batch, clf.input_dim, clf.num_classes)
    off = (st.ws_ptr.value - st.workspace.data_ptr()) + total - 256 - batch * 128
    return st.workspace[off:off + batch * 128].view(batch, 128).clone()


def test_mlp_dropout_stream_statistics():
    """Dropout(0.3) (NB:2977) on the kernel's OWN random stream: keep rate 0.7, no structure across units or rows, a new
    mask per seed, and the saved mask is the one the forward used (replaying it explicitly reproduces the launch)."""
    b = 4096
    clf = ae_b200.MLP(64, 10).to(gu.dev())
    gu.load_mlp(clf, 3)
    clf = clf.to(gu.dev()).train()
    x = torch.randn(b, 64, device=gu.dev())
    y = torch.randint(0, 10, (b,), device=gu.dev())
    loss1, _, logits1 = clf.fused_step_grads(x, y)
    m1 = _mlp_keep_mask(clf, b)
    loss1, logits1 = loss1.clone(), logits1.clone()
    loss2, _, _ = clf.fused_step_grads(x, y)
    m2 = _mlp_keep_mask(clf, b)
    torch.cuda.synchronize()
    assert set(m1.unique().tolist()) <= {0, 1}
    k1 = m1.double()
    n = k1.numel()
    sigma = (0.7 * 0.3 / n) ** 0.5
    assert abs(float(k1.mean()) - 0.7) <= 5 * sigma, float(k1.mean())                      # 0.7 +- 0.003
    per_unit, per_row = k1.mean(0), k1.mean(1)
    assert float((per_unit - 0.7).abs().max()) <= 5 * (0.21 / b) ** 0.5                    # every unit: 0.7 +- 0.036
    assert float((per_row - 0.7).abs().max()) <= 6 * (0.21 / 128) ** 0.5                   # every row
    c = k1 - k1.mean()
    var = float((c * c).mean())
    assert abs(float((c[:, 1:] * c[:, :-1]).mean()) / var) <= 0.01                         # neighbouring units
    assert abs(float((c[1:] * c[:-1]).mean()) / var) <= 0.01                               # neighbouring rows
    agree = float((m1 == m2).double().mean())                                              # another seed: 0.49 + 0.09 = 0.58
    assert abs(agree - 0.58) <= 0.01, agree
    # the saved mask is the one the launch used
    clf.set_dropout_keep_mask(m1)
    loss3, _, logits3 = clf.fused_step_grads(x, y)
    torch.cuda.synchronize()
    assert torch.equal(logits3, logits1) and torch.equal(loss3, loss1)


def test_mlp_eval_large_batch():
    clf = ae_b200.MLP(64, 10).to(gu.dev())
    gu.load_mlp(clf, 3)
    clf = clf.to(gu.dev()).eval()
    xl = torch.randn(70000, 64, device=gu.dev())
    logits, am = clf.predict(xl)
    ref_state = {k: v.detach().cpu().clone() for k, v in clf.state_dict().items()}
    with torch.no_grad():
        ref = tp.mlp_forward(ref_state, xl.cpu(), False)
    assert gu.rel(logits, ref) <= 1e-4
    top2 = ref.topk(2, dim=1).values
    safe = (top2[:, 0] - top2[:, 1]) > 1e-4
    assert torch.equal(am.cpu()[safe], ref.argmax(1)[safe])


def test_no_cpu_fallback_and_errors():
    m = ae_b200.SupervisedAutoencoder(64)
    with pytest.raises(RuntimeError):
        m(torch.zeros(2, 3, 64, 64))
    m = m.to(gu.dev())
    with pytest.raises(RuntimeError):
        m(torch.zeros(2, 3, 32, 32, device=gu.dev()))
    with pytest.raises(RuntimeError):
        ae_b200.SupervisedAutoencoder(60).to(gu.dev())(torch.zeros(2, 3, 64, 64, device=gu.dev()))


def test_frozen_encoder_and_feature_extraction():
    """NB:3434-3441: freeze best_ae.enc, extract latents over a loader of ragged batches."""
    ae = ae_b200.SupervisedAutoencoder(64, backend=gu.BACKENDS[-1]).to(gu.dev())
    st = gu.load_ae(ae, 12)
    for p in ae.enc.parameters():
        p.requires_grad = False
    ae.enc.eval()
    xs = [seeded.seeded_images(b, 40 + i) for i, b in enumerate([64, 64, 48])]
    ys = [seeded.seeded_labels(b, 40 + i) for i, b in enumerate([64, 64, 48])]
    X, Y = ae_b200.extract_features(list(zip(xs, ys)), ae.enc, device=gu.dev())
    assert X.shape == (176, 64) and Y.shape == (176,) and X.is_cuda
    with torch.no_grad():
        ref = tp.encoder_forward(st, torch.cat(xs), False)
    assert gu.rel(X, ref) <= 1e-4
    assert torch.equal(Y.cpu(), torch.cat(ys))


def test_eval_encoder_walks_large_batches_in_chunks(monkeypatch):
    """Inference over more images than AE_B200_EVAL_CHUNK (BASELINE configs[4]: batches up to 64k) walks the batch in
    chunks of the engine's workspace size; images are independent in eval mode, so the latents must not depend on it."""
    ae = ae_b200.SupervisedAutoencoder(64, backend=gu.BACKENDS[-1]).to(gu.dev())
    st = gu.load_ae(ae, 13)
    ae.enc.eval()
    x = seeded.seeded_images(150, 77)
    with torch.no_grad():
        z_one = ae.enc(x.to(gu.dev())).clone()                     # one engine call (150 <= default chunk of 4096)
        monkeypatch.setenv("AE_B200_EVAL_CHUNK", "64")
        z_chunked = ae.enc(x.to(gu.dev()))                         # 64 + 64 + 22
        ref = tp.encoder_forward(st, x, False)
    torch.cuda.synchronize()
    assert gu.rel(z_chunked, z_one) <= 1e-6
    assert gu.rel(z_chunked, ref) <= 1e-4


@pytest.mark.parametrize("latent,batch", [(16, 5), (48, 130), (64, 300), (128, 129), (256, 257)])
@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_eval_bottleneck_on_tensor_cores_matches_oracle_and_cuda_core_path(latent, batch, prec, monkeypatch):
    """Eval-mode Flatten + Linear(4096, latent) (NB:520-521) runs on tcgen05 over the planes conv4's epilogue emits
    (dense_tc.cu): every latent width the kernel accepts, ragged 128-row tiles, against the oracle and against the CUDA-core
    kernel (AE_B200_DENSE_TC=0, read when the engine is created)."""
    x = seeded.seeded_images(batch, 31 + latent)
    zs = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("AE_B200_DENSE_TC", flag)
        ae = ae_b200.SupervisedAutoencoder(latent, precision=prec, backend="tc").to(gu.dev())
        st = gu.load_ae(ae, 17, latent)
        ae.enc.eval()
        with torch.no_grad():
            zs[flag] = ae.enc(x.to(gu.dev())).clone()
    torch.cuda.synchronize()
    ref = tp.encoder_forward(st, x, False)
    assert gu.rel(zs["1"], ref) <= gu.TOL[prec]
    assert gu.rel(zs["1"], zs["0"]) <= gu.TOL[prec]


def test_backward_after_another_forward_raises():
    """The activations of a forward live in the engine's workspace: a backward whose forward was overwritten by a later
    forward must raise instead of silently using the newer activations (gradient accumulation over two forwards, retained
    graphs)."""
    torch.manual_seed(0)
    ae = ae_b200.SupervisedAutoencoder(64, backend=gu.BACKENDS[-1]).to(gu.dev()).train()
    x1, x2 = torch.rand(4, 3, 64, 64, device=gu.dev()), torch.rand(4, 3, 64, 64, device=gu.dev())
    z1 = ae.enc(x1)
    ae.enc(x2)
    with pytest.raises(RuntimeError, match="another forward"):
        z1.sum().backward()
    z3 = ae.enc(x1)                     # a matching pair still works
    z3.sum().backward()
    assert all(p.grad is not None for p in ae.enc.parameters())
    ae.eval()
    big = torch.rand(130, 3, 64, 64, device=gu.dev())
    import os
    os.environ["AE_B200_EVAL_CHUNK"] = "64"
    try:
        assert not ae.enc(big).requires_grad           # chunked inference pass: only the last chunk's activations exist
    finally:
        del os.environ["AE_B200_EVAL_CHUNK"]


def test_optimizer_state_survives_a_larger_batch_and_frozen_parameters_do_not_move():
    """A batch larger than the engine's capacity re-creates the native engine (new workspace) but must keep the flat
    parameter buffers, so Adam's moments and step counter carry on (they used to restart silently).  Parameters without
    a gradient are skipped like torch.optim.Adam skips them: value and moments untouched."""
    torch.manual_seed(1)
    ae = ae_b200.SupervisedAutoencoder(64, backend=gu.BACKENDS[-1]).to(gu.dev()).train()
    x, y = torch.rand(16, 3, 64, 64, device=gu.dev()), torch.randint(0, 10, (16,), device=gu.dev())
    ae.engine().prepare(gu.dev(), 16)
    opt = ae_b200.Adam(ae.parameters(), lr=1e-3)
    ae.train_step_grads(x, y, 35.0)
    opt.step()
    eng = ae.engine()
    flat, inst = eng.flat, eng.instance
    st = opt.flat_state(flat)
    m_before, step_before = st["m"].clone(), int(st["step"][0])
    with torch.no_grad():
        ae.eval()
        ae.enc(torch.rand(200, 3, 64, 64, device=gu.dev()))       # 200 > capacity 64: the native engine is re-created
        ae.train()
    assert eng.instance == inst + 1 and eng.flat is flat and flat.aliased()
    assert opt.flat_state(eng.flat) is st and torch.equal(st["m"], m_before) and int(st["step"][0]) == step_before
    # second step: encoder frozen by dropping its gradients
    ae.train_step_grads(x, y, 35.0)
    enc_before = {k: p.detach().clone() for k, p in ae.enc.named_parameters()}
    dec_before = ae.dec.decoder_input.weight.detach().clone()
    for p in ae.enc.parameters():
        p.grad = None
    lo = ae.enc.encoder[0].weight._ae_flat[1]
    m_enc = st["m"][lo:lo + 864].clone()
    opt.step()
    torch.cuda.synchronize()
    for k, p in ae.enc.named_parameters():
        assert torch.equal(p.detach(), enc_before[k]), k
    assert torch.equal(st["m"][lo:lo + 864], m_enc)
    assert not torch.equal(ae.dec.decoder_input.weight.detach(), dec_before)
    assert int(st["step"][0]) == step_before + 1


def test_invalidate_after_raw_parameter_writes():
    """Writes through .data (dist.broadcast(p.data), p.data.copy_) bump no version counter; engine().invalidate() (called by
    dp.broadcast_parameters) makes the next forward re-derive the packed weights."""
    torch.manual_seed(2)
    a = ae_b200.SupervisedAutoencoder(64, backend=gu.BACKENDS[-1]).to(gu.dev()).eval()
    b = ae_b200.SupervisedAutoencoder(64, backend=gu.BACKENDS[-1]).to(gu.dev()).eval()
    x = torch.rand(8, 3, 64, 64, device=gu.dev())
    with torch.no_grad():
        za, zb = a.enc(x).clone(), b.enc(x).clone()                # packs both models
        assert not torch.allclose(za, zb)
        for (_, pa), (_, pb) in zip(a.state_dict().items(), b.state_dict().items()):
            pa.data.copy_(pb.data)                                 # raw write: no version bump
        a.engine().invalidate()
        assert torch.equal(a.enc(x), zb)


@pytest.mark.parametrize("cluster", ["1", "2", "4", "8"])
@pytest.mark.parametrize("batch", [64, 37])
def test_mlp_train_step_graph_matches_oracle_loop(batch, cluster, monkeypatch):
    """MLPTrainStep (NB:3476-3482 as a replayed two-launch graph: cluster kernel + fused Adam with weight decay 1e-4) against
    the oracle loop over several steps, dropout switched off so both sides see the same network; every cluster size of the
    kernel (the library picks it by batch size; AE_B200_MLP_CLUSTER forces it)."""
    monkeypatch.setenv("AE_B200_MLP_CLUSTER", cluster)
    seed, lr, wd, steps = 41, 1e-3, 1e-4, 5
    st = seeded.seeded_state(seeded.mlp_state_shapes(64, 10), seed)
    rs = np.random.RandomState(batch)
    xs = [torch.from_numpy(rs.standard_normal((batch, 64)).astype(np.float32)) for _ in range(steps)]
    ys = [seeded.seeded_labels(batch, seed + i) for i in range(steps)]
    keep = None                                   # the oracle's 'no dropout' (an explicit mask would be scaled by 1/0.7)
    ref_state, opt_ref = {k: v.clone() for k, v in st.items()}, {}
    ref_losses, ref_after_first = [], None
    for i in range(steps):
        ref_losses.append(float(tp.mlp_train_step(ref_state, opt_ref, xs[i], ys[i], lr, wd, keep)[0]))
        if i == 0:
            ref_after_first = {k: v.clone() for k, v in ref_state.items()}
    clf = ae_b200.MLP(64, 10)
    clf.load_state_dict(st)
    clf = clf.to(gu.dev()).train()
    clf.net[3].p = 0.0
    clf._state.prepare(gu.dev(), batch)
    opt = ae_b200.Adam(clf.parameters(), lr=lr, weight_decay=wd)
    step = ae_b200.MLPTrainStep(clf, opt, batch)
    losses, after_first = [], None
    for i in range(steps):
        step.x.copy_(xs[i]); step.y.copy_(ys[i])
        step.run()
        losses.append(float(step.loss))
        if i == 0:
            after_first = {k: v.detach().clone() for k, v in clf.state_dict().items()}
    torch.cuda.synchronize()
    # step 0 sees identical weights; afterwards Adam's first updates are ~lr * sign(g), so entries with |g| ~ 0 may move the
    # other way in another arithmetic and the trajectories separate slightly (DESIGN.md 'Numerics')
    assert abs(losses[0] - ref_losses[0]) <= 1e-5 * max(1.0, abs(ref_losses[0])), (losses, ref_losses)
    for a, b in zip(losses, ref_losses):
        assert abs(a - b) <= 5e-3 * max(1.0, abs(b)), (losses, ref_losses)
    sd = clf.state_dict()
    assert int(sd["net.1.num_batches_tracked"]) == steps and int(sd["net.5.num_batches_tracked"]) == steps
    for k in ("net.1.running_mean", "net.1.running_var", "net.5.running_mean", "net.5.running_var"):
        assert gu.rel(after_first[k], ref_after_first[k]) <= 1e-4, k       # the first step's batch statistics
    # the first Adam step moves every weight by ~lr: all of them within 2 lr of the oracle, nearly all of them the same way
    d = (after_first["net.4.weight"].cpu() - ref_after_first["net.4.weight"]).abs()
    assert float(d.max()) <= 2.01 * lr and float((d > 0.1 * lr).float().mean()) <= 1e-2


def test_mlp_epoch_mode_matches_stepping_by_hand():
    """fit.train_epoch_mlp replays one captured step per batch that gathers its own rows through the epoch's permutation and
    records (loss, correct) on the device (ae_mlp_train_step_indexed).  With dropout off it must reproduce, bit for bit, an
    MLPTrainStep fed by hand with the same batches -- ragged tail batch included -- and a second epoch must re-use the graphs."""
    from ae_b200 import fit
    seed, lr, wd, n, bs = 23, 1e-3, 1e-4, 150, 64
    st = seeded.seeded_state(seeded.mlp_state_shapes(64, 10), seed)
    rs = np.random.RandomState(7)
    X = torch.from_numpy(rs.standard_normal((n, 64)).astype(np.float32)).to(gu.dev())
    y = seeded.seeded_labels(n, seed).to(gu.dev())

    def make():
        clf = ae_b200.MLP(64, 10)
        clf.load_state_dict(st)
        clf = clf.to(gu.dev()).train()
        clf.net[3].p = 0.0
        clf._state.prepare(gu.dev(), bs)
        return clf, ae_b200.Adam(clf.parameters(), lr=lr, weight_decay=wd)

    gen = torch.Generator(device=gu.dev()).manual_seed(5)
    orders = [torch.randperm(n, device=gu.dev(), generator=gen) for _ in range(2)]
    # by hand
    clf_a, opt_a = make()
    steps = {b: ae_b200.MLPTrainStep(clf_a, opt_a, b) for b in (64, 22)}
    hand = []
    for order in orders:
        tot_l, tot_c = 0.0, 0
        for k in range(0, n, bs):
            idx = order[k:k + bs]
            s = steps[int(idx.numel())]
            s.x.copy_(X[idx]); s.y.copy_(y[idx])
            s.run()
            tot_l += float(s.loss) * int(idx.numel()); tot_c += int(s.correct)
        hand.append((tot_l / n, tot_c / n))
    # epoch mode with the same permutations
    clf_b, opt_b = make()
    gen = torch.Generator(device=gu.dev()).manual_seed(5)
    got = [fit.train_epoch_mlp(clf_b, opt_b, X, y, bs, True, gen) for _ in range(2)]
    torch.cuda.synchronize()
    assert len(clf_b._train_steps) == 1 and sorted(next(iter(clf_b._train_steps.values()))["steps"]) == [22, 64]
    for (la, ca), (lb, cb) in zip(hand, got):
        assert abs(la - lb) <= 1e-6 * max(1.0, abs(la)) and ca == cb
    sa, sb = clf_a.state_dict(), clf_b.state_dict()
    for k in sa:
        assert torch.equal(sa[k], sb[k]), k
