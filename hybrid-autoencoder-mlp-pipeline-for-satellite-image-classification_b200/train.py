"""Whole-step fast path: [H2D] -> CUDA graph {train step -> (NCCL allreduce) -> Adam -> weight re-pack}.

The graph is captured once by the library (ae_step_graph_capture) on fixed device buffers; every call
copies the batch into those buffers and replays the graph.  Equivalent to the loop body NB:2673-2684.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import check, ptr
from .modules import SupervisedAutoencoder
from .optim import Adam


class TrainStep:
    def __init__(self, model: SupervisedAutoencoder, optimizer: Adam, alpha: float, batch: int, comm=None,
                 device=None):
        lib = _lib.load()
        self.model, self.opt, self.alpha, self.batch, self.comm = model, optimizer, float(alpha), int(batch), comm
        dev = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.device = dev
        eng = model.engine()
        eng.prepare(dev, batch)
        self.x = torch.zeros(batch, 3, 64, 64, dtype=torch.float32, device=dev)
        self.y = torch.zeros(batch, dtype=torch.int64, device=dev)
        self.loss = torch.zeros(4, dtype=torch.float32, device=dev)
        flat = eng.flat
        group = optimizer.param_groups[0]
        st = optimizer.flat_state(flat)
        b1, b2 = group["betas"]
        cfg = _lib.AdamConfig(float(group["lr"]), float(b1), float(b2), float(group["eps"]), float(group["weight_decay"]))
        self.stream = torch.cuda.Stream(device=dev)
        self.stream.wait_stream(torch.cuda.current_stream(dev))
        h = C.c_void_p()
        with torch.cuda.stream(self.stream):
            check(lib.ae_step_graph_capture(eng.handle, ptr(self.x), ptr(self.y), batch, self.alpha, ptr(self.loss),
                                            ptr(flat.data), ptr(flat.grad), ptr(st["m"]), ptr(st["v"]), flat.len,
                                            C.byref(cfg), ptr(st["step"]), comm.handle if comm is not None else None,
                                            C.c_void_p(self.stream.cuda_stream), C.byref(h)))
        self.handle = h
        self.stream.synchronize()
        self._eng = eng
        self.num_kernels = lib.ae_step_graph_num_kernels(h)

    def load(self, imgs: torch.Tensor, labels: torch.Tensor):
        """Async copy of one batch (host pinned or device) into the graph's input buffers."""
        with torch.cuda.stream(self.stream):
            self.x.copy_(imgs, non_blocking=True)
            self.y.copy_(labels, non_blocking=True)

    def run(self):
        check(_lib.load().ae_step_graph_launch(self.handle, C.c_void_p(self.stream.cuda_stream)))
        self._eng.flat.generation += 1
        self._eng.mark_packed()      # the graph re-packs the weights itself

    def __call__(self, imgs, labels):
        """One step with torch stream semantics: inputs produced on the current stream are waited for, and the
        returned loss tensor [loss, mse, ce] is safe to read from the current stream."""
        cur = torch.cuda.current_stream(self.device)
        self.stream.wait_stream(cur)
        self.load(imgs, labels)
        self.run()
        cur.wait_stream(self.stream)
        return self.loss[:3]

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                _lib.load().ae_step_graph_destroy(self.handle)
                self.handle = None
        except Exception:
            pass
