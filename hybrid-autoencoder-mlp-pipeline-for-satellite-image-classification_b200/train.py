"""Whole-step fast path: [H2D] -> CUDA graph {train step -> (NCCL allreduce) -> Adam -> weight re-pack}.

The graph is captured once by the library (ae_step_graph_capture) on fixed device buffers; every call
copies the batch into those buffers and replays the graph.  Equivalent to the loop body NB:2673-2684.
"""
from __future__ import annotations

import ctypes as C
import gc

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
        if comm is not None:
            from . import dp
            dp.attach_peers(comm, model)         # collective; no-op when already attached or AE_B200_DP_FUSED=0
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
        self._instance = (eng.instance, eng.flat_instance)
        self.num_kernels = lib.ae_step_graph_num_kernels(h)

    def load(self, imgs: torch.Tensor, labels: torch.Tensor):
        """Async copy of one batch (host pinned or device) into the graph's input buffers."""
        with torch.cuda.stream(self.stream):
            self.x.copy_(imgs, non_blocking=True)
            self.y.copy_(labels, non_blocking=True)

    def run(self, stream=None):
        """Replay the captured step on `stream` (default: this TrainStep's own stream)."""
        if (self._eng.instance, self._eng.flat_instance) != self._instance:
            # the captured graph addresses the workspace and the flat buffers of the engine it was captured on
            raise RuntimeError("ae_b200: the model's engine or parameter storage was re-created (a larger batch was run "
                               "through it, or the model was moved) after this TrainStep was captured; build a new TrainStep")
        check(_lib.load().ae_step_graph_launch(self.handle, C.c_void_p((stream or self.stream).cuda_stream)))
        self._eng.flat.generation += 1
        self._eng.mark_packed()      # the graph re-packs the weights itself
        for part in (_lib.PART_ENC, _lib.PART_DEC, _lib.PART_HEAD):
            self._eng.stamp_forward(part)   # the workspace now holds this step's activations

    def __call__(self, imgs, labels):
        """One step with torch stream semantics: inputs produced on the current stream are waited for, and the
        returned loss tensor [loss, mse, ce] is safe to read from the current stream."""
        cur = torch.cuda.current_stream(self.device)
        self.stream.wait_stream(cur)
        self.load(imgs, labels)
        self.run()
        cur.wait_stream(self.stream)
        return self.loss[:3]

    def run_batches(self, batches, depth: int = 2):
        """The loop NB:2672-2688 over an iterable of (imgs, labels) HOST batches (pinned memory for real overlap),
        software-pipelined: the H2D copy of batch i+1 runs on a copy stream while the graph of batch i executes, and
        the loss of batch i is read back (D2H, pinned) while batch i+1 runs.  Every batch is copied host->device and
        every loss device->host; nothing is skipped.  Returns the list of [loss, mse, ce] host tensors."""
        dev = self.device
        if depth < 2:
            raise ValueError("TrainStep.run_batches: depth must be >= 2 (one staging slot is filled while the other is read)")
        if not hasattr(self, "_copy_stream") or len(self._stage) != depth:
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._stage = [(torch.empty_like(self.x), torch.empty_like(self.y)) for _ in range(depth)]
            self._staged = [torch.cuda.Event() for _ in range(depth)]
            self._consumed = [torch.cuda.Event() for _ in range(depth)]
            self._loss_host = [torch.zeros(4).pin_memory() for _ in range(depth)]
            self._loss_done = [torch.cuda.Event() for _ in range(depth)]
        cs, ms = self._copy_stream, self.stream
        cs.wait_stream(torch.cuda.current_stream(dev))
        ms.wait_stream(torch.cuda.current_stream(dev))
        out, pending = [], []
        it = iter(batches)

        def stage(i, batch):
            sx, sy = self._stage[i % depth]
            with torch.cuda.stream(cs):
                if i >= depth:
                    cs.wait_event(self._consumed[i % depth])      # the step that used this staging slot has copied it out
                sx.copy_(batch[0], non_blocking=True)
                sy.copy_(batch[1], non_blocking=True)
                self._staged[i % depth].record(cs)

        nxt = next(it, None)
        i = 0
        if nxt is not None:
            stage(0, nxt)
        while nxt is not None:
            cur_i = i
            nxt = next(it, None)
            if nxt is not None:
                stage(cur_i + 1, nxt)                              # overlaps the graph launched below
            sx, sy = self._stage[cur_i % depth]
            with torch.cuda.stream(ms):
                ms.wait_event(self._staged[cur_i % depth])
                self.x.copy_(sx, non_blocking=True)                # device-to-device, a few microseconds
                self.y.copy_(sy, non_blocking=True)
                self._consumed[cur_i % depth].record(ms)
                self.run()
                if len(pending) >= depth:                          # the host slot is free once its loss was consumed
                    j = pending.pop(0)
                    self._loss_done[j % depth].synchronize()
                    out.append(self._loss_host[j % depth][:3].clone())
                self._loss_host[cur_i % depth].copy_(self.loss, non_blocking=True)
                self._loss_done[cur_i % depth].record(ms)
            pending.append(cur_i)
            i += 1
        for j in pending:
            self._loss_done[j % depth].synchronize()
            out.append(self._loss_host[j % depth][:3].clone())
        torch.cuda.current_stream(dev).wait_stream(ms)
        return out

    def _sibling(self, batch: int) -> "TrainStep":
        """The same step captured for another batch size (the loader's last, smaller batch); shares model, optimizer
        state and workspace."""
        if not hasattr(self, "_siblings"):
            self._siblings = {}
        sib = self._siblings.get(batch)
        if sib is None:
            if batch > self.batch:
                raise RuntimeError("ae_b200: a loader batch larger than the TrainStep's batch size")
            sib = TrainStep(self.model, self.opt, self.alpha, batch, comm=self.comm, device=self.device)
            self._siblings[batch] = sib
        return sib

    def run_loader(self, loader):
        """One training epoch, NB:2672-2688, over an ``ae_b200.data.DeviceLoader``: every batch is gathered / augmented
        straight into the graph's input buffer by one kernel, the step graph is replayed, and the step's loss is kept in
        a device-side history; the host synchronises ONCE, at the end (the reference synchronises every step through
        ``loss.item()``, NB:2687).  Returns (losses [steps,3] host tensor of [loss, mse, ce], batch sizes list)."""
        dev, ds = self.device, loader.dataset
        hist = torch.zeros(len(loader), 4, dtype=torch.float32, device=dev)
        sizes = []
        self.stream.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(self.stream):
            for k, idx in enumerate(loader.batches()):
                b = int(idx.numel())
                st = self if b == self.batch else self._sibling(b)
                loader.transform(ds.images, idx, out=st.x)
                torch.index_select(ds.labels, 0, idx, out=st.y)
                st.run(self.stream)
                hist[k].copy_(st.loss, non_blocking=True)
                sizes.append(b)
        self.stream.synchronize()
        return hist[:len(sizes), :3].cpu(), sizes

    def close(self):
        """Destroy the captured graph (required before the NCCL communicator it references is destroyed)."""
        for sib in getattr(self, "_siblings", {}).values():
            sib.close()
        if getattr(self, "handle", None):
            self.stream.synchronize()
            _lib.load().ae_step_graph_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MLPTrainStep:
    """One MLP training step of the reference loop (NB:3476-3482: zero_grad, forward, cross-entropy, backward,
    Adam.step) captured ONCE as a CUDA graph of two launches -- forward + BatchNorm1d + dropout + CE + backward in one kernel
    (one CTA with every intermediate in shared memory for batches up to 64 rows, a thread-block cluster above that) and the
    fused flat Adam -- on fixed device buffers.  Nothing that changes between steps is baked in: the
    dropout seed advances on the device.  `load` copies / gathers a batch into the buffers, `run` replays."""

    def __init__(self, clf, optimizer: Adam, batch: int, device=None, epoch=None):
        """`epoch` = (X, y, order, cursor, hist): epoch mode -- the captured step reads its batch as rows
        order[cursor[0] ..] of the device-resident X / y, stores (loss, correct) in hist[cursor[1]] and advances the cursor
        itself (ae_mlp_train_step_indexed), so an epoch is one `run()` per batch and nothing else on the host."""
        lib = _lib.load()
        self.clf, self.opt, self.batch = clf, optimizer, int(batch)
        self.epoch = epoch
        dev = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.device = dev
        st = clf._state
        st.prepare(dev, batch)
        self._st, self._flat = st, st.flat
        d, c = clf.input_dim, clf.num_classes
        self.x = torch.zeros(batch, d, dtype=torch.float32, device=dev)
        self.y = torch.zeros(batch, dtype=torch.int64, device=dev)
        self.loss = torch.zeros(1, dtype=torch.float32, device=dev)
        self.correct = torch.zeros(1, dtype=torch.int32, device=dev)
        self.logits = torch.empty(batch, c, dtype=torch.float32, device=dev)
        self.seed_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        group = optimizer.param_groups[0]
        ast = optimizer.flat_state(st.flat)
        b1, b2 = group["betas"]
        self.key = (float(group["lr"]), float(b1), float(b2), float(group["eps"]), float(group["weight_decay"]))
        cfg = _lib.AdamConfig(*self.key)
        seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        self._ws = st.workspace                      # the graph addresses this workspace: keep it alive, detect a re-allocation
        stream = torch.cuda.Stream(device=dev)
        stream.wait_stream(torch.cuda.current_stream(dev))

        def call():
            cur = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            if epoch is not None:
                X, y, order, cursor, hist = epoch
                assert X.dtype == torch.float32 and y.dtype == torch.int64 and X.is_contiguous() and y.is_contiguous()
                assert order.dtype == torch.int64 and cursor.dtype == torch.int64 and cursor.numel() == 2 and hist.dtype == torch.float32
                check(lib.ae_mlp_train_step_indexed(ptr(st.flat.data), ptr(st.flat.grad), ptr(st.running), ptr(st.steps), ptr(X), ptr(y),
                                                    ptr(order), ptr(cursor), ptr(hist), seed, ptr(self.seed_dev), float(clf.net[3].p),
                                                    batch, d, c, ptr(self.logits), ptr(self.loss), ptr(self.correct), st.ws_ptr,
                                                    st.ws_bytes, C.byref(cfg), ptr(ast["m"]), ptr(ast["v"]), ptr(ast["step"]), cur))
                return
            check(lib.ae_mlp_train_step(ptr(st.flat.data), ptr(st.flat.grad), ptr(st.running), ptr(st.steps), ptr(self.x), ptr(self.y),
                                        seed, ptr(self.seed_dev), float(clf.net[3].p), batch, d, c, ptr(self.logits), ptr(self.loss),
                                        ptr(self.correct), st.ws_ptr, st.ws_bytes, C.byref(cfg), ptr(ast["m"]), ptr(ast["v"]),
                                        ptr(ast["step"]), cur))
        # Collect garbage BEFORE the capture and capture in relaxed mode: a cyclic-GC pass that happens to run inside the capture
        # window may finalise an old TrainStep, whose close() synchronises its own stream -- in the default (global) mode that
        # unrelated call invalidates this capture.
        gc.collect()
        with torch.cuda.stream(stream):
            # (no warm-up launch: it would be a real optimizer step; the kernels allocate nothing)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, stream=stream, capture_error_mode="relaxed"):
                call()
        torch.cuda.current_stream(dev).wait_stream(stream)
        for p_, g in zip(st.flat.params, st.flat.grad_views(st.flat.grad)):
            if p_.requires_grad:
                p_.grad = g

    def valid(self) -> bool:
        st = self._st
        group = self.opt.param_groups[0]
        b1, b2 = group["betas"]
        key = (float(group["lr"]), float(b1), float(b2), float(group["eps"]), float(group["weight_decay"]))
        return st.flat is self._flat and st.workspace is self._ws and st.flat.aliased() and key == self.key

    def run(self):
        """Replay on the current stream; loss [1], correct [1] and logits are valid afterwards (stream-ordered)."""
        self.graph.replay()
        self._flat.generation += 1
