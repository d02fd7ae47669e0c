"""Fused flat-buffer Adam with the torch.optim.Adam surface the reference uses.

NB:2654  torch.optim.Adam(model.parameters(), lr=learning_rate)              (no weight decay)
NB:3461  torch.optim.Adam(clf.parameters(), lr=lr, weight_decay=1e-4)        (coupled L2)
NB:2676 / NB:2684  optimizer.zero_grad(); optimizer.step()

All parameters of an ae_b200 model live in one flat fp32 buffer; ``step()`` is ONE launch of the
vectorised 128-bit kernel ae_adam_step_flat over that buffer (plus a gather when the gradients were
produced by the autograd path and are not already the flat gradient buffer).
"""
from __future__ import annotations

import ctypes as C
import weakref

import torch

from . import _lib
from ._lib import check, ptr, stream_ptr


class Adam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1) or not (0 <= betas[1] < 1) or weight_decay < 0:
            raise ValueError("invalid Adam hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._flat_state = weakref.WeakKeyDictionary()   # flat buffer object -> dict(m, v, step); dies with the buffer
        self.grad_scale = 1.0     # data parallel: 1/world after a sum-allreduce
        self._on_step = []        # callbacks run after the parameters changed (weight re-pack)

    def _flat_of(self, group):
        flats = {}
        for p in group["params"]:
            fl = getattr(p, "_ae_flat", None)
            if fl is None or not fl[0].aliased():
                raise RuntimeError("ae_b200.Adam: parameters must belong to an ae_b200 model that has run at least one "
                                   "forward on the GPU (flat storage); there is no per-tensor fallback")
            flats.setdefault(id(fl[0]), (fl[0], []))[1].append((p, fl[1]))
        return list(flats.values())

    def flat_state(self, flat):
        st = self._flat_state.get(flat)
        if st is None:
            dev = flat.data.device
            st = dict(m=torch.zeros_like(flat.data), v=torch.zeros_like(flat.data),
                      step=torch.zeros(2, dtype=torch.int32, device=dev))
            self._flat_state[flat] = st
        return st

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        for group in self.param_groups:
            for flat, plist in self._flat_of(group):
                st = self.flat_state(flat)
                base = flat.grad.data_ptr()
                # gradients produced by autograd live elsewhere: gather them into the flat gradient buffer
                srcs, dsts = [], []
                stepped = set()
                for p, off in plist:
                    n = p.numel()
                    g = p.grad
                    if g is None:
                        continue          # torch.optim.Adam skips parameters without a gradient: frozen for this step
                    stepped.add(id(p))
                    if g.data_ptr() != base + 4 * off or not g.is_contiguous():
                        srcs.append(g.reshape(-1))
                        dsts.append(flat.grad[off:off + n])
                if srcs:
                    torch._foreach_copy_(dsts, srcs)
                # the kernel walks the whole flat buffer: parameters that must not move (no gradient this step, frozen, or
                # not in this group) are restored afterwards together with their moments
                keep = [(o, q.numel(), q.detach().clone(), st["m"][o:o + q.numel()].clone(), st["v"][o:o + q.numel()].clone())
                        for q, o in zip(flat.params, flat.offsets) if id(q) not in stepped]
                b1, b2 = group["betas"]
                check(lib.ae_adam_step_flat(ptr(flat.data), ptr(flat.grad), ptr(st["m"]), ptr(st["v"]), flat.len,
                                            float(group["lr"]), float(b1), float(b2), float(group["eps"]),
                                            float(group["weight_decay"]), float(self.grad_scale), ptr(st["step"]),
                                            stream_ptr()))
                flat.generation += 1
                for o, n, val, m0, v0 in keep:
                    flat.data[o:o + n].copy_(val.reshape(-1))
                    st["m"][o:o + n].copy_(m0)
                    st["v"][o:o + n].copy_(v0)
        for cb in self._on_step:
            cb()
        return loss
