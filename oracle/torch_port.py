"""Functional CPU restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY.

Every function cites the reference lines it restates (NB:n = line n of the notebook's raw
JSON).  The arithmetic is expressed with ``torch.nn.functional`` on CPU
fp32 tensors held in a plain ``state`` dict keyed exactly like the reference
``state_dict()``; there are no ``nn.Module`` objects here.  Pinned against the reference's
own cells by ``oracle/make_golden.py`` + ``tests/test_oracle_pin.py``.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

BN_EPS = 1e-5       # torch default, used by every BatchNorm in NB:505-517, NB:617-625, NB:2974-2980
BN_MOMENTUM = 0.1
DROPOUT_P = 0.3     # NB:2977


def _bn(state, prefix, x, training, new_buffers):
    """BatchNorm2d/1d: batch statistics + running update in training, running stats in eval."""
    rm, rv = state[prefix + ".running_mean"], state[prefix + ".running_var"]
    if training:
        rm, rv = rm.clone(), rv.clone()
        y = F.batch_norm(x, rm, rv, state[prefix + ".weight"], state[prefix + ".bias"], True, BN_MOMENTUM, BN_EPS)
        new_buffers[prefix + ".running_mean"] = rm
        new_buffers[prefix + ".running_var"] = rv
        new_buffers[prefix + ".num_batches_tracked"] = state[prefix + ".num_batches_tracked"] + 1
        return y
    return F.batch_norm(x, rm, rv, state[prefix + ".weight"], state[prefix + ".bias"], False, BN_MOMENTUM, BN_EPS)


def encoder_forward(state, x, training: bool, new_buffers=None, prefix: str = "enc."):
    """Encoder.forward, NB:499-525: 4 x [Conv2d k3 s2 p1, BatchNorm2d, ReLU], Flatten, Linear."""
    nb = {} if new_buffers is None else new_buffers
    h = x
    for i in range(4):
        p = f"{prefix}encoder.{3 * i}"
        h = F.conv2d(h, state[p + ".weight"], state[p + ".bias"], stride=2, padding=1)   # NB:504,508,512,516
        h = _bn(state, f"{prefix}encoder.{3 * i + 1}", h, training, nb)                  # NB:505,509,513,517
        h = F.relu(h)                                                                    # NB:506..518
    h = h.flatten(1)                                                                     # NB:520
    return F.linear(h, state[prefix + "encoder.13.weight"], state[prefix + "encoder.13.bias"])  # NB:521


def decoder_forward(state, z, training: bool, new_buffers=None, prefix: str = "dec."):
    """Decoder.forward, NB:607-635: Linear, Unflatten(256,4,4), 3 x [ConvT, BN, ReLU], ConvT, Sigmoid."""
    nb = {} if new_buffers is None else new_buffers
    h = F.linear(z, state[prefix + "decoder_input.weight"], state[prefix + "decoder_input.bias"])  # NB:611,633
    h = h.view(-1, 256, 4, 4)                                                            # NB:614
    for i in range(4):
        p = f"{prefix}decoder.{3 * i + 1}"
        h = F.conv_transpose2d(h, state[p + ".weight"], state[p + ".bias"], stride=2, padding=1,
                               output_padding=1)                                         # NB:616,620,624,628
        if i < 3:
            h = _bn(state, f"{prefix}decoder.{3 * i + 2}", h, training, nb)              # NB:617,621,625
            h = F.relu(h)
    return torch.sigmoid(h)                                                              # NB:629


def head_forward(state, z, prefix: str = "classifier."):
    """SupervisedAutoencoder.classifier, NB:692-696: Linear, ReLU, Linear."""
    h = F.relu(F.linear(z, state[prefix + "0.weight"], state[prefix + "0.bias"]))
    return F.linear(h, state[prefix + "2.weight"], state[prefix + "2.bias"])


def ae_forward(state, x, training: bool, new_buffers=None):
    """SupervisedAutoencoder.forward, NB:698-702: returns (x_hat, logits, z)."""
    z = encoder_forward(state, x, training, new_buffers)
    x_hat = decoder_forward(state, z, training, new_buffers)
    logits = head_forward(state, z)
    return x_hat, logits, z


def ae_loss(x_hat, logits, x, labels, alpha: float):
    """NB:2679-2681: loss = alpha * MSELoss()(x_hat, imgs) + CrossEntropyLoss()(logits, labels)."""
    loss_recon = F.mse_loss(x_hat, x)
    loss_class = F.cross_entropy(logits, labels)
    return alpha * loss_recon + loss_class, loss_recon, loss_class


def mlp_forward(state, x, training: bool, dropout_keep: Optional[torch.Tensor] = None, new_buffers=None,
                prefix: str = "net."):
    """MLP.forward, NB:2970-2987.  ``dropout_keep`` is the {0,1} keep mask of Dropout(0.3)
    (kept values are scaled by 1/0.7); ``None`` means no dropout (eval, or p forced to 0)."""
    nb = {} if new_buffers is None else new_buffers
    h = F.linear(x, state[prefix + "0.weight"], state[prefix + "0.bias"])
    h = F.relu(_bn(state, prefix + "1", h, training, nb))
    if training and dropout_keep is not None:
        h = h * dropout_keep * (1.0 / (1.0 - DROPOUT_P))
    h = F.linear(h, state[prefix + "4.weight"], state[prefix + "4.bias"])
    h = F.relu(_bn(state, prefix + "5", h, training, nb))
    return F.linear(h, state[prefix + "7.weight"], state[prefix + "7.bias"])


def is_param(key: str) -> bool:
    leaf = key.rsplit(".", 1)[1]
    return leaf in ("weight", "bias")


def param_keys(state):
    return [k for k in state if is_param(k)]


def adam_update(p, g, m, v, step: int, lr: float, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0):
    """torch.optim.Adam single-tensor update (NB:2654 / NB:3461; amsgrad off, coupled L2).
    Returns (p, m, v).  Same op order as torch's _single_tensor_adam."""
    if weight_decay != 0.0:
        g = g + weight_decay * p
    m = m + (g - m) * (1.0 - beta1)                      # exp_avg.lerp_(grad, 1 - beta1)
    v = v * beta2 + (g * g) * (1.0 - beta2)              # exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    step_size = lr / bc1
    denom = v.sqrt() / math.sqrt(bc2) + eps
    p = p - step_size * (m / denom)
    return p, m, v


def ae_train_step(state, opt, x, labels, alpha: float, lr: float):
    """One iteration of the AE training loop body, NB:2676-2684 (zero_grad, forward, loss,
    backward, Adam.step).  ``opt`` = {"step": int, "m": {k: t}, "v": {k: t}} (created when empty).
    Returns (loss, loss_recon, loss_class, grads, outputs); mutates ``state`` and ``opt``."""
    keys = param_keys(state)
    leaves = {k: state[k].detach().clone().requires_grad_(True) for k in keys}
    work = dict(state)
    work.update(leaves)
    nb: Dict[str, torch.Tensor] = {}
    x_hat, logits, z = ae_forward(work, x, True, nb)
    loss, lr_, lc_ = ae_loss(x_hat, logits, x, labels, alpha)
    grads_t = torch.autograd.grad(loss, [leaves[k] for k in keys])
    grads = OrderedDict((k, g.detach()) for k, g in zip(keys, grads_t))
    if not opt:
        opt.update({"step": 0, "m": {k: torch.zeros_like(state[k]) for k in keys},
                    "v": {k: torch.zeros_like(state[k]) for k in keys}})
    opt["step"] += 1
    for k in keys:
        p, m, v = adam_update(state[k], grads[k], opt["m"][k], opt["v"][k], opt["step"], lr)
        state[k], opt["m"][k], opt["v"][k] = p, m, v
    state.update(nb)
    return (loss.detach(), lr_.detach(), lc_.detach(), grads,
            (x_hat.detach(), logits.detach(), z.detach()))


def mlp_train_step(state, opt, x, labels, lr: float, weight_decay: float = 1e-4,
                   dropout_keep: Optional[torch.Tensor] = None):
    """One iteration of the MLP loop body, NB:3477-3482 (CE loss, Adam with coupled L2 1e-4)."""
    keys = param_keys(state)
    leaves = {k: state[k].detach().clone().requires_grad_(True) for k in keys}
    work = dict(state)
    work.update(leaves)
    nb: Dict[str, torch.Tensor] = {}
    logits = mlp_forward(work, x, True, dropout_keep, nb)
    loss = F.cross_entropy(logits, labels)
    grads_t = torch.autograd.grad(loss, [leaves[k] for k in keys])
    grads = OrderedDict((k, g.detach()) for k, g in zip(keys, grads_t))
    if not opt:
        opt.update({"step": 0, "m": {k: torch.zeros_like(state[k]) for k in keys},
                    "v": {k: torch.zeros_like(state[k]) for k in keys}})
    opt["step"] += 1
    for k in keys:
        p, m, v = adam_update(state[k], grads[k], opt["m"][k], opt["v"][k], opt["step"], lr,
                              weight_decay=weight_decay)
        state[k], opt["m"][k], opt["v"][k] = p, m, v
    state.update(nb)
    return loss.detach(), grads, logits.detach()


def encode_predict(ae_state, mlp_state, x) -> Tuple[torch.Tensor, torch.Tensor]:
    """BASELINE config 1/5: clf(enc(x)) in eval mode (NB:2908 composed with NB:3702)."""
    with torch.no_grad():
        z = encoder_forward(ae_state, x, False)
        logits = mlp_forward(mlp_state, z, False)
    return z, logits
