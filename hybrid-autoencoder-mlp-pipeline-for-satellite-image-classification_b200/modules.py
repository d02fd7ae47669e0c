"""nn.Module shells with the reference's constructors, attributes, forward outputs and state_dict keys.

Reference surface mirrored here (NB:n = line n of the notebook's raw JSON):
  Encoder(latent_dim)                       NB:499-525    forward(x) -> z
  Decoder(latent_dim)                       NB:607-635    forward(z) -> x_hat
  SupervisedAutoencoder(latent_dim, nc=10)  NB:685-702    forward(x) -> (x_hat, logits, z)
  MLP(input_dim, num_classes=10)            NB:2970-2987  forward(x) -> logits

The shells hold their parameters and BatchNorm buffers in stock torch layer objects (so default
initialisation, ``state_dict()`` keys, ``load_state_dict``, ``.to()``, ``requires_grad`` and checkpoints
behave exactly like the reference), but those layer objects are never called: ``forward`` re-points the
parameters into one flat fp32 buffer per model and runs the hand-written sm_100a kernels of
libae_b200.so through ``torch.autograd.Function``s.  There is no PyTorch / CPU fallback: a forward on a
non-CUDA tensor raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional

import torch
import torch.nn as nn

from . import _lib
from ._lib import check, ptr, stream_ptr

_PREC = {"fp32": _lib.PREC_FP32, "bf16": _lib.PREC_BF16}
_BACKEND = {"tc": _lib.BACKEND_TC, "simt": _lib.BACKEND_SIMT}


def default_precision() -> str:
    return os.environ.get("AE_B200_PRECISION", "fp32")


def default_backend() -> str:
    return os.environ.get("AE_B200_BACKEND", "tc")


def eval_chunk() -> int:
    """Images per engine call of an eval-mode encoder pass over a larger batch."""
    return max(64, int(os.environ.get("AE_B200_EVAL_CHUNK", "4096")))


def _require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise RuntimeError(f"ae_b200: {what} must be a CUDA tensor (no CPU fallback); got device {t.device}")
    if t.dtype != torch.float32:
        raise RuntimeError(f"ae_b200: {what} must be float32, got {t.dtype}")


# ------------------------------------------------------------------------------------------------
# flat storage
# ------------------------------------------------------------------------------------------------
class _Flat:
    """One flat fp32 buffer holding a list of parameters at given (4-aligned) offsets."""

    def __init__(self, params: List[nn.Parameter], offsets: List[int], flat_len: int, device):
        self.params = params
        self.offsets = offsets
        self.len = flat_len
        self.data = torch.zeros(flat_len, dtype=torch.float32, device=device)
        self.grad = torch.zeros(flat_len, dtype=torch.float32, device=device)
        with torch.no_grad():
            for p, off in zip(params, offsets):
                n = p.numel()
                self.data[off:off + n].copy_(p.detach().reshape(-1).to(device=device, dtype=torch.float32))
                p.data = self.data[off:off + n].view(p.shape)
                p._ae_flat = (self, off)
        self.generation = 0      # bumped by code that changes the parameters through raw pointers (fused Adam)

    def aliased(self) -> bool:
        base = self.data.data_ptr()
        return all(p.data_ptr() == base + 4 * off for p, off in zip(self.params, self.offsets))

    def version_sum(self) -> int:
        return sum(p._version for p in self.params) + (self.generation << 32)

    def grad_views(self, src: torch.Tensor):
        return [src[off:off + p.numel()].view(p.shape) for p, off in zip(self.params, self.offsets)]


class _Engine:
    """Owns the ae_engine_t handle, the workspace and the flat buffers of up to three parts."""

    def __init__(self, latent_dim: int, num_classes: int, precision: str, backend: str):
        self.latent_dim, self.num_classes = latent_dim, num_classes
        self.precision, self.backend = precision, backend
        self.handle = None
        self.max_batch = 0
        self.device = None
        self.workspace = None
        self.parts = {}          # part id -> dict(params=[...], bns=[...])
        self.flat: Optional[_Flat] = None
        self.running = None      # flat fp32 running stats
        self.steps = None        # flat int64 num_batches_tracked
        self.part_off = {}
        self.run_off = {}
        self.step_off = {}
        self.packed_version = None
        self.instance = 0        # bumped whenever the native engine (and its workspace) is re-created
        self.flat_instance = 0   # bumped whenever the flat parameter / gradient buffers are re-allocated
        self.fwd_seq = {}        # part id -> sequence number of the last forward (its activations live in the workspace)

    def register(self, part: int, params: List[nn.Parameter], bns: List[nn.Module]):
        self.parts[part] = dict(params=params, bns=bns)

    def destroy(self):
        if self.handle is not None:
            _lib.load().ae_engine_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass

    # -- (re)build ------------------------------------------------------------------------------
    def _create(self, device, batch: int):
        lib = _lib.load()
        if not lib.ae_device_supported(device.index if device.index is not None else torch.cuda.current_device()):
            raise RuntimeError("ae_b200: this library is built for sm_100a (B200) only")
        self.destroy()
        cap = 64
        while cap < batch:
            cap *= 2
        cfg = _lib.EngineConfig(self.latent_dim, self.num_classes, cap, _PREC[self.precision], _BACKEND[self.backend])
        h = C.c_void_p()
        check(lib.ae_engine_create(C.byref(cfg), C.byref(h)))
        self.handle, self.max_batch, self.device = h, cap, device
        self.instance += 1
        nbytes = lib.ae_engine_workspace_bytes(h)
        self.workspace = torch.empty(nbytes + 256, dtype=torch.uint8, device=device)
        base = self.workspace.data_ptr()
        self.ws_ptr = C.c_void_p((base + 255) & ~255)
        check(lib.ae_engine_bind_workspace(h, self.ws_ptr, nbytes))
        # the flat parameter / gradient buffers (and with them every optimizer's moments) survive the re-creation: only
        # the workspace and the native handle are new
        if self.flat is not None and self.flat.data.device == device and self.flat.aliased() and self._buffers_aliased():
            self._bind_parts()
        else:
            self.flat = None
        self.packed_version = None
        self.fwd_seq = {}

    def _flatten(self):
        lib = _lib.load()
        params, offsets, total = [], [], 0
        run_total, step_total = 0, 0
        self.part_off, self.run_off, self.step_off = {}, {}, {}
        layouts = {}
        for part in sorted(self.parts):
            n = len(self.parts[part]["params"])
            offs = (C.c_int64 * 64)()
            sizes = (C.c_int64 * 64)()
            flen = C.c_int64()
            cnt = lib.ae_engine_param_layout(self.handle, part, offs, sizes, C.byref(flen))
            if cnt != n:
                raise RuntimeError(f"ae_b200: part {part} has {n} parameters, the engine expects {cnt}")
            for p, o, s in zip(self.parts[part]["params"], offs, sizes):
                if p.numel() != s:
                    raise RuntimeError(f"ae_b200: parameter of {p.numel()} elements where the engine expects {s}")
                params.append(p)
                offsets.append(total + o)
            self.part_off[part] = total
            total += flen.value
            ch = (C.c_int * 16)()
            nb = lib.ae_engine_bn_layout(self.handle, part, ch)
            layouts[part] = [ch[i] for i in range(nb)]
            self.run_off[part] = run_total
            self.step_off[part] = step_total
            run_total += 2 * sum(layouts[part])
            step_total += nb
        self.flat = _Flat(params, offsets, total, self.device)
        self.flat_instance += 1
        self.running = torch.zeros(max(run_total, 4), dtype=torch.float32, device=self.device)
        self.steps = torch.zeros(max(step_total, 1), dtype=torch.int64, device=self.device)
        with torch.no_grad():
            for part in sorted(self.parts):
                ro, so = self.run_off[part], self.step_off[part]
                for i, bn in enumerate(self.parts[part]["bns"]):
                    c = layouts[part][i]
                    self.running[ro:ro + c].copy_(bn.running_mean.to(self.device))
                    self.running[ro + c:ro + 2 * c].copy_(bn.running_var.to(self.device))
                    self.steps[so + i].copy_(bn.num_batches_tracked.to(self.device))
                    bn._buffers["running_mean"] = self.running[ro:ro + c]
                    bn._buffers["running_var"] = self.running[ro + c:ro + 2 * c]
                    bn._buffers["num_batches_tracked"] = self.steps[so + i]
                    ro += 2 * c
        self._bind_parts()
        self.packed_version = None

    def _bind_parts(self):
        lib = _lib.load()
        for part in sorted(self.parts):
            po = self.part_off[part]
            has_bn = len(self.parts[part]["bns"]) > 0
            check(lib.ae_engine_bind_part(
                self.handle, part, C.c_void_p(self.flat.data.data_ptr() + 4 * po),
                C.c_void_p(self.flat.grad.data_ptr() + 4 * po),
                C.c_void_p(self.running.data_ptr() + 4 * self.run_off[part]) if has_bn else None,
                C.c_void_p(self.steps.data_ptr() + 8 * self.step_off[part]) if has_bn else None))

    def _buffers_aliased(self) -> bool:
        for part in self.parts:
            ro = self.run_off[part]
            for bn in self.parts[part]["bns"]:
                if bn.running_mean.data_ptr() != self.running.data_ptr() + 4 * ro:
                    return False
                ro += 2 * bn.running_mean.numel()
        return True

    def prepare(self, device, batch: int):
        """Make sure handle, flat storage and packed weights are current.  Cheap when nothing changed."""
        if self.handle is None or self.device != device or batch > self.max_batch:
            self._create(device, batch)
        if self.flat is None or not self.flat.aliased() or not self._buffers_aliased():
            self._flatten()
        v = self.flat.version_sum()
        if self.packed_version != v:
            self.pack()
            self.packed_version = self.flat.version_sum()

    def pack(self):
        lib = _lib.load()
        for part in sorted(self.parts):
            check(lib.ae_engine_pack_weights(self.handle, part, stream_ptr()))

    def mark_packed(self):
        self.packed_version = self.flat.version_sum()

    def invalidate(self):
        """Call after changing parameters in a way torch's version counters do not see (``p.data.copy_(...)``,
        ``dist.broadcast(p.data)``): the next forward re-derives the packed weights."""
        if self.flat is not None:
            self.flat.generation += 1

    # -- forward / backward pairing: the activations of a forward live in the engine's workspace, not in ctx -------------
    def stamp_forward(self, part: int) -> int:
        self.fwd_seq[part] = self.fwd_seq.get(part, 0) + 1
        return self.fwd_seq[part]

    def check_backward(self, part: int, seq: int, what: str):
        if self.fwd_seq.get(part) != seq:
            raise RuntimeError(
                f"ae_b200: backward of a {what} forward whose activations are gone: another forward of the same module ran in "
                "between (the engine keeps ONE set of activations per module; run backward before the next forward, e.g. no "
                "gradient accumulation over two forwards and no retained graphs)")


def _part_grads(engine: "_Engine", part: int):
    """Snapshot of one part's slice of the flat gradient buffer as per-parameter views (autograd accumulates what a
    Function returns, so the engine's own buffer must not be handed out)."""
    plist = engine.parts[part]["params"]
    lo = plist[0]._ae_flat[1]
    hi = plist[-1]._ae_flat[1] + plist[-1].numel()
    snap = engine.flat.grad[lo:hi].clone()
    return [snap[p._ae_flat[1] - lo:p._ae_flat[1] - lo + p.numel()].view(p.shape) for p in plist]


# ------------------------------------------------------------------------------------------------
# autograd functions (one per part)
# ------------------------------------------------------------------------------------------------
class _EncoderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, engine: _Engine, training: bool, *params):
        _require_cuda(x, "input images")
        if x.dim() != 4 or tuple(x.shape[1:]) != (3, 64, 64):
            raise RuntimeError(f"ae_b200: expected images of shape [B,3,64,64], got {tuple(x.shape)}")
        x = x.contiguous()
        b = x.shape[0]
        z = torch.empty(b, engine.latent_dim, dtype=torch.float32, device=x.device)
        if not training and b > eval_chunk():
            # inference over a large batch (BASELINE config 5: 1k-64k images): images are independent in eval mode, so
            # the batch is walked in chunks and the workspace stays bounded (2.5 MB per image of the chunk)
            n = eval_chunk()
            engine.prepare(x.device, n)
            lib = _lib.load()
            for i in range(0, b, n):
                m = min(n, b - i)
                check(lib.ae_encoder_forward(engine.handle, ptr(x[i:i + m]), m, 0, ptr(z[i:i + m]), stream_ptr()))
            engine.stamp_forward(_lib.PART_ENC)
            ctx.mark_non_differentiable(z)         # only the last chunk's activations exist: inference only
            return z
        engine.prepare(x.device, b)
        check(_lib.load().ae_encoder_forward(engine.handle, ptr(x), b, int(training), ptr(z), stream_ptr()))
        ctx.engine, ctx.b, ctx.training = engine, b, training
        ctx.seq = engine.stamp_forward(_lib.PART_ENC)
        ctx.save_for_backward(x)
        return z

    @staticmethod
    def backward(ctx, dz):
        engine = ctx.engine
        if not ctx.training:
            raise RuntimeError("ae_b200: backward through an eval-mode encoder forward is not supported (the engine keeps "
                               "batch statistics only in training mode); call .train() or wrap the forward in no_grad()")
        engine.check_backward(_lib.PART_ENC, ctx.seq, "encoder")
        dz = dz.contiguous()
        check(_lib.load().ae_encoder_backward(engine.handle, ptr(dz), ctx.b, stream_ptr()))
        grads = _part_grads(engine, _lib.PART_ENC)
        grads = [g if need else None for g, need in zip(grads, ctx.needs_input_grad[3:])]
        return (None, None, None, *grads)


class _DecoderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, engine: _Engine, training: bool, *params):
        _require_cuda(z, "latent")
        z = z.contiguous()
        b = z.shape[0]
        engine.prepare(z.device, b)
        x_hat = torch.empty(b, 3, 64, 64, dtype=torch.float32, device=z.device)
        check(_lib.load().ae_decoder_forward(engine.handle, ptr(z), b, int(training), ptr(x_hat), stream_ptr()))
        ctx.engine, ctx.b, ctx.training = engine, b, training
        ctx.seq = engine.stamp_forward(_lib.PART_DEC)
        ctx.save_for_backward(z)
        return x_hat

    @staticmethod
    def backward(ctx, d_xhat):
        engine = ctx.engine
        if not ctx.training:
            raise RuntimeError("ae_b200: backward through an eval-mode decoder forward is not supported; call .train() or "
                               "wrap the forward in no_grad()")
        engine.check_backward(_lib.PART_DEC, ctx.seq, "decoder")
        d_xhat = d_xhat.contiguous()
        dz = torch.empty(ctx.b, engine.latent_dim, dtype=torch.float32, device=d_xhat.device)
        check(_lib.load().ae_decoder_backward(engine.handle, ptr(d_xhat), ctx.b, ptr(dz), stream_ptr()))
        grads = _part_grads(engine, _lib.PART_DEC)
        grads = [g if need else None for g, need in zip(grads, ctx.needs_input_grad[3:])]
        return (dz if ctx.needs_input_grad[0] else None, None, None, *grads)


class _HeadFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, engine: _Engine, *params):
        _require_cuda(z, "latent")
        z = z.contiguous()
        b = z.shape[0]
        engine.prepare(z.device, b)
        logits = torch.empty(b, engine.num_classes, dtype=torch.float32, device=z.device)
        check(_lib.load().ae_head_forward(engine.handle, ptr(z), b, ptr(logits), stream_ptr()))
        ctx.engine, ctx.b = engine, b
        ctx.seq = engine.stamp_forward(_lib.PART_HEAD)
        ctx.save_for_backward(z)
        return logits

    @staticmethod
    def backward(ctx, d_logits):
        engine = ctx.engine
        engine.check_backward(_lib.PART_HEAD, ctx.seq, "classifier-head")
        d_logits = d_logits.contiguous()
        dz = torch.empty(ctx.b, engine.latent_dim, dtype=torch.float32, device=d_logits.device)
        check(_lib.load().ae_head_backward(engine.handle, ptr(d_logits), ctx.b, ptr(dz), stream_ptr()))
        grads = _part_grads(engine, _lib.PART_HEAD)
        grads = [g if need else None for g, need in zip(grads, ctx.needs_input_grad[2:])]
        return (dz if ctx.needs_input_grad[0] else None, None, *grads)


# ------------------------------------------------------------------------------------------------
# shells
# ------------------------------------------------------------------------------------------------
class _PartModule(nn.Module):
    _part = -1

    def _container_params(self) -> List[nn.Parameter]:
        return list(self.parameters())

    def _bns(self):
        return [m for m in self.modules() if isinstance(m, (nn.BatchNorm2d, nn.BatchNorm1d))]

    def _own_engine(self, latent_dim, num_classes=10, precision=None, backend=None):
        eng = _Engine(latent_dim, num_classes, precision or default_precision(), backend or default_backend())
        object.__setattr__(self, "_engine", eng)
        eng.register(self._part, self._container_params(), self._bns())

    def _adopt(self, engine: _Engine):
        object.__setattr__(self, "_engine", engine)
        engine.register(self._part, self._container_params(), self._bns())


class Encoder(_PartModule):
    """NB:499-525.  4 x [Conv2d k3 s2 p1, BatchNorm2d, ReLU], Flatten, Linear(4096, latent_dim)."""
    _part = _lib.PART_ENC

    def __init__(self, latent_dim, *, precision=None, backend=None):
        super().__init__()
        layers = []
        ch = [3, 32, 64, 128, 256]
        for i in range(4):
            layers += [nn.Conv2d(ch[i], ch[i + 1], 3, stride=2, padding=1), nn.BatchNorm2d(ch[i + 1]), nn.ReLU()]
        layers += [nn.Flatten(), nn.Linear(256 * 4 * 4, latent_dim)]
        self.encoder = nn.Sequential(*layers)     # parameter containers only; never called
        self.latent_dim = latent_dim
        self._own_engine(latent_dim, precision=precision, backend=backend)

    def forward(self, x):
        return _EncoderFn.apply(x, self._engine, self.training, *self._container_params())


class Decoder(_PartModule):
    """NB:607-635.  Linear(latent,4096), Unflatten(256,4,4), 3 x [ConvT k3 s2 p1 op1, BN, ReLU], ConvT(32,3), Sigmoid."""
    _part = _lib.PART_DEC

    def __init__(self, latent_dim, *, precision=None, backend=None):
        super().__init__()
        self.decoder_input = nn.Linear(latent_dim, 256 * 4 * 4)
        layers = [nn.Unflatten(1, (256, 4, 4))]
        ch = [256, 128, 64, 32]
        for i in range(3):
            layers += [nn.ConvTranspose2d(ch[i], ch[i + 1], 3, stride=2, padding=1, output_padding=1),
                       nn.BatchNorm2d(ch[i + 1]), nn.ReLU()]
        layers += [nn.ConvTranspose2d(32, 3, 3, stride=2, padding=1, output_padding=1), nn.Sigmoid()]
        self.decoder = nn.Sequential(*layers)     # parameter containers only; never called
        self.latent_dim = latent_dim
        self._own_engine(latent_dim, precision=precision, backend=backend)

    def forward(self, z):
        return _DecoderFn.apply(z, self._engine, self.training, *self._container_params())


class _Head(_PartModule, nn.Sequential):
    """NB:692-696: Linear(latent,128), ReLU, Linear(128,num_classes) -- keys classifier.{0,2}.{weight,bias}."""
    _part = _lib.PART_HEAD

    def __init__(self, latent_dim, num_classes):
        nn.Sequential.__init__(self, nn.Linear(latent_dim, 128), nn.ReLU(), nn.Linear(128, num_classes))

    def forward(self, z):
        return _HeadFn.apply(z, self._engine, *self._container_params())


class SupervisedAutoencoder(nn.Module):
    """NB:685-702.  forward(x) -> (x_hat [B,3,64,64], logits [B,num_classes], z [B,latent_dim])."""

    def __init__(self, latent_dim, num_classes=10, *, precision=None, backend=None):
        super().__init__()
        self.enc = Encoder(latent_dim, precision=precision, backend=backend)
        self.dec = Decoder(latent_dim, precision=precision, backend=backend)
        self.classifier = _Head(latent_dim, num_classes)
        self.latent_dim, self.num_classes = latent_dim, num_classes
        eng = _Engine(latent_dim, num_classes, precision or default_precision(), backend or default_backend())
        object.__setattr__(self, "_engine", eng)
        self.enc._adopt(eng)
        self.dec._adopt(eng)
        self.classifier._adopt(eng)

    def forward(self, x):
        z = self.enc(x)
        x_hat = self.dec(z)
        logits = self.classifier(z)
        return x_hat, logits, z

    # ---- fused fast paths (not part of the reference surface) ----
    def engine(self) -> _Engine:
        return self._engine

    def train_step_grads(self, imgs, labels, alpha: float):
        """NB:2676-2683 in one library call: forward, alpha*MSE + CE, backward.  Gradients land in
        ``param.grad`` (views of the flat gradient buffer); returns a device tensor [loss, mse, ce]."""
        _require_cuda(imgs, "input images")
        eng = self._engine
        imgs = imgs.contiguous()
        b = imgs.shape[0]
        eng.prepare(imgs.device, b)
        loss = torch.empty(4, dtype=torch.float32, device=imgs.device)
        check(_lib.load().ae_train_step(eng.handle, ptr(imgs), ptr(labels.contiguous()), b, float(alpha), ptr(loss),
                                        stream_ptr()))
        for part in (_lib.PART_ENC, _lib.PART_DEC, _lib.PART_HEAD):
            eng.stamp_forward(part)
        for p, off in zip(eng.flat.params, eng.flat.offsets):
            if p.requires_grad:
                p.grad = eng.flat.grad[off:off + p.numel()].view(p.shape)
        return loss[:3]

    @torch.no_grad()
    def eval_step(self, imgs, labels, alpha: float):
        """NB:2694-2714: eval-mode forward + loss.  Returns (loss[3], x_hat, logits, z)."""
        _require_cuda(imgs, "input images")
        eng = self._engine
        imgs = imgs.contiguous()
        b = imgs.shape[0]
        eng.prepare(imgs.device, b)
        dev = imgs.device
        loss = torch.empty(4, dtype=torch.float32, device=dev)
        x_hat = torch.empty(b, 3, 64, 64, dtype=torch.float32, device=dev)
        logits = torch.empty(b, self.num_classes, dtype=torch.float32, device=dev)
        z = torch.empty(b, self.latent_dim, dtype=torch.float32, device=dev)
        check(_lib.load().ae_eval_step(eng.handle, ptr(imgs), ptr(labels.contiguous()), b, float(alpha), ptr(loss),
                                       ptr(x_hat), ptr(logits), ptr(z), stream_ptr()))
        for part in (_lib.PART_ENC, _lib.PART_DEC, _lib.PART_HEAD):
            eng.stamp_forward(part)
        return loss[:3], x_hat, logits, z


# ------------------------------------------------------------------------------------------------
# MLP
# ------------------------------------------------------------------------------------------------
class _MLPState:
    def __init__(self, module: "MLP"):
        self.module = module
        self.flat: Optional[_Flat] = None
        self.running = None
        self.steps = None
        self.workspace = None
        self.ws_batch = 0

    def prepare(self, device, batch):
        m = self.module
        lib = _lib.load()
        params = list(m.parameters())
        if self.flat is None or self.flat.data.device != device or not self.flat.aliased() or \
                m.net[1].running_mean.data_ptr() != (self.running.data_ptr() if self.running is not None else 0):
            offs = (C.c_int64 * 10)()
            sizes = (C.c_int64 * 10)()
            total = lib.ae_mlp_param_layout(m.input_dim, m.num_classes, offs, sizes)
            for p, s in zip(params, sizes):
                if p.numel() != s:
                    raise RuntimeError("ae_b200: unexpected MLP parameter shape")
            self.flat = _Flat(params, list(offs), total, device)
            self.running = torch.zeros(2 * 128 + 2 * 64, dtype=torch.float32, device=device)
            self.steps = torch.zeros(2, dtype=torch.int64, device=device)
            with torch.no_grad():
                o = 0
                for i, bn in enumerate((m.net[1], m.net[5])):
                    c = bn.num_features
                    self.running[o:o + c].copy_(bn.running_mean.to(device))
                    self.running[o + c:o + 2 * c].copy_(bn.running_var.to(device))
                    self.steps[i].copy_(bn.num_batches_tracked.to(device))
                    bn._buffers["running_mean"] = self.running[o:o + c]
                    bn._buffers["running_var"] = self.running[o + c:o + 2 * c]
                    bn._buffers["num_batches_tracked"] = self.steps[i]
                    o += 2 * c
        if self.workspace is None or batch > self.ws_batch or self.workspace.device != device:
            cap = max(64, 1 << (batch - 1).bit_length())
            n = lib.ae_mlp_workspace_bytes(cap, m.input_dim, m.num_classes)
            self.workspace = torch.empty(n + 256, dtype=torch.uint8, device=device)
            self.ws_batch = cap
            self.ws_bytes = n
        base = self.workspace.data_ptr()
        self.ws_ptr = C.c_void_p((base + 255) & ~255)


class _MLPFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, state: _MLPState, training: bool, *params):
        _require_cuda(x, "MLP input")
        m = state.module
        x = x.contiguous()
        b = x.shape[0]
        if x.dim() != 2 or x.shape[1] != m.input_dim:
            raise RuntimeError(f"ae_b200: expected MLP input [B,{m.input_dim}], got {tuple(x.shape)}")
        state.prepare(x.device, b)
        lib = _lib.load()
        logits = torch.empty(b, m.num_classes, dtype=torch.float32, device=x.device)
        if not training:
            check(lib.ae_mlp_forward_eval(ptr(state.flat.data), ptr(state.running), ptr(x), b, m.input_dim, m.num_classes,
                                          ptr(logits), None, stream_ptr()))
            return logits
        keep = m._dropout_keep_override
        if keep is not None:
            keep = keep.to(device=x.device, dtype=torch.uint8).contiguous()
        seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if keep is None else 0
        p_drop = float(m.net[3].p)
        check(lib.ae_mlp_fwd_bwd_ce(ptr(state.flat.data), None, ptr(state.running), ptr(x), None, ptr(keep), seed, p_drop,
                                    b, m.input_dim, m.num_classes, 1, ptr(logits), None, None, state.ws_ptr,
                                    state.ws_bytes, stream_ptr()))
        state.steps += 1
        ctx.state, ctx.b, ctx.p_drop = state, b, p_drop
        ctx.save_for_backward(x)
        return logits

    @staticmethod
    def backward(ctx, d_logits):
        state = ctx.state
        m = state.module
        (x,) = ctx.saved_tensors
        d_logits = d_logits.contiguous()
        check(_lib.load().ae_mlp_backward(ptr(state.flat.data), ptr(state.flat.grad), ptr(x), ptr(d_logits), ctx.p_drop,
                                          ctx.b, m.input_dim, m.num_classes, state.ws_ptr, state.ws_bytes, stream_ptr()))
        snap = state.flat.grad.clone()
        grads = state.flat.grad_views(snap)
        grads = [g if need else None for g, need in zip(grads, ctx.needs_input_grad[3:])]
        return (None, None, None, *grads)


class MLP(nn.Module):
    """NB:2970-2987.  Linear(D,128) BN1d ReLU Dropout(0.3) Linear(128,64) BN1d ReLU Linear(64,num_classes)."""

    def __init__(self, input_dim, num_classes=10):
        super().__init__()
        self.net = nn.Sequential(
            nn.Linear(input_dim, 128), nn.BatchNorm1d(128), nn.ReLU(), nn.Dropout(0.3),
            nn.Linear(128, 64), nn.BatchNorm1d(64), nn.ReLU(),
            nn.Linear(64, num_classes))           # parameter containers only; never called
        self.input_dim, self.num_classes = input_dim, num_classes
        object.__setattr__(self, "_state", _MLPState(self))
        object.__setattr__(self, "_dropout_keep_override", None)

    def forward(self, x):
        return _MLPFn.apply(x, self._state, self.training, *self.parameters())

    def set_dropout_keep_mask(self, keep: Optional[torch.Tensor]):
        """Test hook: use this {0,1} keep mask [B,128] instead of the kernel's own random stream."""
        object.__setattr__(self, "_dropout_keep_override", keep)

    def fused_step_grads(self, x, labels):
        """NB:3477-3481 in one kernel launch: forward + cross-entropy + backward.  Returns (loss, correct) device
        tensors; gradients land in ``param.grad``."""
        _require_cuda(x, "MLP input")
        st = self._state
        x = x.contiguous()
        b = x.shape[0]
        st.prepare(x.device, b)
        keep = self._dropout_keep_override
        if keep is not None:
            keep = keep.to(device=x.device, dtype=torch.uint8).contiguous()
        seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if keep is None else 0
        loss = torch.empty(1, dtype=torch.float32, device=x.device)
        correct = torch.empty(1, dtype=torch.int32, device=x.device)
        logits = torch.empty(b, self.num_classes, dtype=torch.float32, device=x.device)
        check(_lib.load().ae_mlp_fwd_bwd_ce(ptr(st.flat.data), ptr(st.flat.grad), ptr(st.running), ptr(x),
                                            ptr(labels.contiguous()), ptr(keep), seed, float(self.net[3].p), b,
                                            self.input_dim, self.num_classes, 1, ptr(logits), ptr(loss), ptr(correct),
                                            st.ws_ptr, st.ws_bytes, stream_ptr()))
        st.steps += 1
        for p, g in zip(st.flat.params, st.flat.grad_views(st.flat.grad)):
            if p.requires_grad:
                p.grad = g
        return loss, correct, logits

    @torch.no_grad()
    def predict(self, x):
        """clf.eval(); clf(x).argmax(1) (NB:3702) in one kernel."""
        _require_cuda(x, "MLP input")
        st = self._state
        x = x.contiguous()
        b = x.shape[0]
        st.prepare(x.device, b)
        logits = torch.empty(b, self.num_classes, dtype=torch.float32, device=x.device)
        am = torch.empty(b, dtype=torch.int64, device=x.device)
        check(_lib.load().ae_mlp_forward_eval(ptr(st.flat.data), ptr(st.running), ptr(x), b, self.input_dim,
                                              self.num_classes, ptr(logits), ptr(am), stream_ptr()))
        return logits, am
