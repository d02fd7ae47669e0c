#!/usr/bin/env python
"""Device time of the four 3-channel-layer launches (thin.cu) through the C ABI, each replayed in a CUDA graph over
rotating buffers larger than L2:  python scripts/thin_bench.py [batch ...]   (default 256 4096)

  conv1 forward (bias + statistics), conv1 forward with the eval epilogue (BatchNorm + ReLU + split planes),
  conv1 weight gradient, ConvTranspose2d(32,3) forward + sigmoid + squared error, its fused backward.
Prints per launch: microseconds, algorithmic GB/s (bytes the launch must read + write / time) and FMA rate.
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import ae_b200  # noqa: F401
from ae_b200 import _lib
from tests import gpu_util as gu


def timed(fn, nrot, iters=int(os.environ.get("THIN_ITERS", "20"))):
    """fn(i) launches on the current stream with buffer set i % nrot; returns microseconds per launch (graph replay)."""
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for i in range(nrot):
            fn(i)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for i in range(nrot):
                fn(i)
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (iters * nrot)


def run(batch):
    d = gu.dev()
    lib = gu.lib()
    rs = np.random.RandomState(1)
    img_b, wide_b = batch * 3 * 64 * 64 * 4, batch * 32 * 32 * 32 * 4
    nrot = max(2, int(300e6 // (img_b + wide_b)) + 1)
    pb = (gu.PREC["fp32"], gu.BACK["simt"])
    xs = [torch.rand(batch, 3, 64, 64, device=d) for _ in range(nrot)]
    xh = [torch.rand(batch, 3, 64, 64, device=d) for _ in range(nrot)]
    wide = [torch.randn(batch, 32, 32, 32, device=d) for _ in range(nrot)]
    outw = [torch.empty(batch, 32, 32, 32, device=d) for _ in range(nrot)]
    w = torch.from_numpy((rs.standard_normal((32, 3, 3, 3)) / 5).astype(np.float32)).to(d)
    b32 = torch.zeros(32, device=d)
    b3 = torch.zeros(3, device=d)
    bnc = gu.make_bnc(32, rs, d)
    stats = torch.zeros(64, dtype=torch.float64, device=d)
    sse = torch.zeros(2, dtype=torch.float64, device=d)
    nb = lib.ae_thin_wgrad_workspace_bytes(batch)
    part = torch.empty(nb, dtype=torch.uint8, device=d)
    dw, db = torch.empty(32, 3, 3, 3, device=d), torch.empty(3, device=d)
    fma = batch * 1024 * 32 * 27

    def st():
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def conv1_fwd(i):
        op = gu.operand(xs[i % nrot])
        ep = gu.epilogue(_lib.EPI_BIAS_STATS, b32, None, None, stats)
        _lib.check(lib.ae_thin_gather_fwd(C.byref(op), gu.p(w), C.byref(ep), gu.p(outw[i % nrot]), batch, *pb, st()))

    def conv1_eval(i):
        op = gu.operand(xs[i % nrot])
        ep = gu.epilogue(_lib.EPI_BNRELU_SPLIT, b32, None, bnc)
        _lib.check(lib.ae_thin_gather_fwd(C.byref(op), gu.p(w), C.byref(ep), gu.p(outw[i % nrot]), batch, *pb, st()))

    def conv1_wgrad(i):
        opw = gu.operand(wide[i % nrot], outw[i % nrot], bnc, 0.0, _lib.OP_BNBWD)
        opt = gu.operand(xs[i % nrot])
        _lib.check(lib.ae_thin_wgrad(C.byref(opw), C.byref(opt), gu.p(dw), None, gu.p(part), nb, batch, *pb, st()))

    def convt_fwd(i):
        opw = gu.operand(wide[i % nrot], None, bnc, 0.0, _lib.OP_BNRELU)
        _lib.check(lib.ae_thin_scatter_sigmoid_fwd(C.byref(opw), gu.p(w), gu.p(b3), gu.p(xh[i % nrot]), gu.p(xs[i % nrot]),
                                                   gu.p(sse), batch, *pb, st()))

    def convt_bwd(i):
        opw = gu.operand(wide[i % nrot], None, bnc, 0.0, _lib.OP_BNRELU)
        opt = gu.operand(xs[i % nrot], xh[i % nrot], None, 1e-3, _lib.OP_SIGMOID_BWD)
        ep = gu.epilogue(_lib.EPI_RELUBWD_STATS, None, wide[i % nrot], bnc, stats)
        _lib.check(lib.ae_thin_bwd_fused(C.byref(opw), C.byref(opt), gu.p(w), C.byref(ep), gu.p(outw[i % nrot]), gu.p(dw),
                                         gu.p(db), gu.p(part), nb, batch, *pb, st()))

    rows = [("conv1 forward (bias, statistics)", conv1_fwd, img_b + wide_b, fma),
            ("conv1 forward (eval: BN+ReLU+planes)", conv1_eval, img_b + wide_b, fma),
            ("conv1 weight gradient", conv1_wgrad, img_b + 2 * wide_b, fma),
            ("convT(32,3) forward + sigmoid + SSE", convt_fwd, 2 * img_b + wide_b, fma),
            ("convT(32,3) fused backward", convt_bwd, 2 * img_b + 3 * wide_b, 2 * fma)]
    only = os.environ.get("THIN_ONLY")                      # e.g. THIN_ONLY=3 profiles one row under ncu
    if only:
        rows = [rows[int(i)] for i in only.split(",")]
    for name, fn, nbytes, nfma in rows:
        us = timed(fn, nrot)
        print(f"batch {batch:5d}  {name:40s} {us:8.2f} us   {nbytes / us / 1e3:7.1f} GB/s   {nfma / us / 1e6:6.2f} TFMA/s")


if __name__ == "__main__":
    for b in ([int(a) for a in sys.argv[1:]] or [256, 4096]):
        run(b)
