"""Per-operator parity: every C-ABI operator against the CPU oracle (torch fp32 functional ops, the
same arithmetic the reference's nn.Modules dispatch to) on seeded inputs.  Run with -m gpu on a B200."""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from ae_b200 import _lib
from tests import gpu_util as gu

pytestmark = pytest.mark.gpu

SHAPES = [  # (batch, hs, cb, cs): the three mid-layer geometries of the model (a2-a4 / a7-a9) + ragged batches
    (2, 16, 32, 64), (3, 8, 64, 128), (5, 4, 128, 256), (1, 4, 128, 256), (33, 16, 32, 64),
]


def _geom(b, hs, cb, cs):
    return _lib.ConvGeom(b, hs, hs, cb, cs)


@pytest.fixture(params=["auto", "gen1", "gen2"])
def rowgemm_gen(request, monkeypatch):
    """Both generations of the tcgen05 row GEMM on every shape: "auto" is the library's own choice (the second generation
    from ~100 tiles up), "gen2" forces the second generation wherever it supports the problem (AE_B200_ROWGEMM, read per call)."""
    if request.param != "auto":
        monkeypatch.setenv("AE_B200_ROWGEMM", request.param[-1])
    return request.param


def _apply_operand_cpu(mode, src, src2, bnc):
    """NCHW cpu tensors; bnc [8][C] cpu."""
    v = lambda r: bnc[r].view(1, -1, 1, 1)
    if mode == _lib.OP_RAW:
        return src
    if mode == _lib.OP_BNRELU:
        return torch.relu(src * v(0) + v(1))
    return v(4) * src + v(5) * (src2 - v(2)) + v(6)


@pytest.mark.parametrize("backend", gu.BACKENDS)
@pytest.mark.parametrize("prec", gu.PRECISIONS)
@pytest.mark.parametrize("mode", [_lib.OP_RAW, _lib.OP_BNRELU, _lib.OP_BNBWD])
@pytest.mark.parametrize("shape", SHAPES)
def test_conv_fwd(shape, mode, prec, backend, rowgemm_gen):
    if backend == "simt" and prec == "bf16":
        pytest.skip("the CUDA-core backend always computes in fp32")
    if backend == "simt" and rowgemm_gen != "auto":
        pytest.skip("row-GEMM generations are a tcgen05-path switch")
    b, hs, cb, cs = shape
    rs = np.random.RandomState(hash((shape, mode)) % 2 ** 31)
    d = gu.dev()
    big = torch.from_numpy(rs.standard_normal((b, cb, 2 * hs, 2 * hs)).astype(np.float32))
    big2 = torch.from_numpy(rs.standard_normal((b, cb, 2 * hs, 2 * hs)).astype(np.float32))
    w = torch.from_numpy((rs.standard_normal((cs, cb, 3, 3)) / np.sqrt(9 * cb)).astype(np.float32))
    bias = torch.from_numpy(rs.uniform(-0.1, 0.1, cs).astype(np.float32))
    bnc = gu.make_bnc(cb, rs, "cpu")
    ref = F.conv2d(_apply_operand_cpu(mode, big, big2, bnc), w, bias, stride=2, padding=1)
    wd = w.to(d)
    pf, _ = gu.pack_conv(wd, cs, cb, prec, backend)
    xb, xb2, bncd, biasd = gu.nhwc(big).to(d), gu.nhwc(big2).to(d), bnc.to(d), bias.to(d)
    out = torch.empty(b, hs, hs, cs, device=d)
    stats = torch.zeros(2 * cs, dtype=torch.float64, device=d)
    g = _geom(b, hs, cb, cs)
    op, _keep = gu.conv_operand(backend, prec, cb, xb, xb2, bncd, mode)
    ep = gu.epilogue(_lib.EPI_BIAS_STATS, biasd, None, None, stats)
    _lib.check(gu.lib().ae_conv2d_s2_fwd(C.byref(g), C.byref(op), gu.p(pf), C.byref(ep), gu.p(out), gu.PREC[prec],
                                        gu.BACK[backend], gu.stream()))
    torch.cuda.synchronize()
    got = gu.nchw(out).cpu()
    assert gu.rel(got, ref) <= gu.TOL[prec]
    s1 = ref.double().sum(dim=(0, 2, 3))
    s2 = (ref.double() ** 2).sum(dim=(0, 2, 3))
    assert gu.rel(stats[:cs], s1) <= gu.TOL[prec] * 10
    assert gu.rel(stats[cs:], s2) <= gu.TOL[prec] * 10


@pytest.mark.parametrize("backend", gu.BACKENDS)
@pytest.mark.parametrize("prec", gu.PRECISIONS)
@pytest.mark.parametrize("mode", [_lib.OP_RAW, _lib.OP_BNRELU, _lib.OP_BNBWD])
@pytest.mark.parametrize("shape", SHAPES)
def test_conv_dgrad_is_transposed_conv(shape, mode, prec, backend, rowgemm_gen):
    if backend == "simt" and prec == "bf16":
        pytest.skip("the CUDA-core backend always computes in fp32")
    if backend == "simt" and rowgemm_gen != "auto":
        pytest.skip("row-GEMM generations are a tcgen05-path switch")
    b, hs, cb, cs = shape
    rs = np.random.RandomState(hash((shape, mode, 1)) % 2 ** 31)
    d = gu.dev()
    small = torch.from_numpy(rs.standard_normal((b, cs, hs, hs)).astype(np.float32))
    small2 = torch.from_numpy(rs.standard_normal((b, cs, hs, hs)).astype(np.float32))
    w = torch.from_numpy((rs.standard_normal((cs, cb, 3, 3)) / np.sqrt(9 * cs / 4)).astype(np.float32))
    bias = torch.from_numpy(rs.uniform(-0.1, 0.1, cb).astype(np.float32))
    bnc = gu.make_bnc(cs, rs, "cpu")
    a = _apply_operand_cpu(mode, small, small2, bnc)
    ref = F.conv_transpose2d(a, w, bias, stride=2, padding=1, output_padding=1)
    # RELUBWD epilogue: mask with the sign of scale*y+shift of the layer below, statistics of dz and dz*xhat
    yb = torch.from_numpy(rs.standard_normal((b, cb, 2 * hs, 2 * hs)).astype(np.float32))
    bnc_o = gu.make_bnc(cb, rs, "cpu")
    ref_nob = F.conv_transpose2d(a, w, None, stride=2, padding=1, output_padding=1)
    v = lambda r: bnc_o[r].view(1, -1, 1, 1)
    mask = (yb * v(0) + v(1)) > 0
    ref_dz = ref_nob * mask
    xhat = (yb - v(2)) * v(3)
    wd = w.to(d)
    _, pd = gu.pack_conv(wd, cs, cb, prec, backend)
    g = _geom(b, hs, cb, cs)
    sd, sd2, bncd = gu.nhwc(small).to(d), gu.nhwc(small2).to(d), bnc.to(d)   # keep alive: operand holds raw pointers
    op, _keep = gu.conv_operand(backend, prec, cs, sd, sd2, bncd, mode)
    out = torch.empty(b, 2 * hs, 2 * hs, cb, device=d)
    stats = torch.zeros(2 * cb, dtype=torch.float64, device=d)
    biasd = bias.to(d)
    ep = gu.epilogue(_lib.EPI_BIAS_STATS, biasd, None, None, stats)
    _lib.check(gu.lib().ae_conv2d_s2_dgrad(C.byref(g), C.byref(op), gu.p(pd), C.byref(ep), gu.p(out), gu.PREC[prec],
                                          gu.BACK[backend], gu.stream()))
    torch.cuda.synchronize()
    assert gu.rel(gu.nchw(out).cpu(), ref) <= gu.TOL[prec]
    assert gu.rel(stats[:cb], ref.double().sum(dim=(0, 2, 3))) <= gu.TOL[prec] * 10
    ybd, bncod = gu.nhwc(yb).to(d), bnc_o.to(d)
    stats.zero_()
    ep2 = gu.epilogue(_lib.EPI_RELUBWD_STATS, None, ybd, bncod, stats)
    _lib.check(gu.lib().ae_conv2d_s2_dgrad(C.byref(g), C.byref(op), gu.p(pd), C.byref(ep2), gu.p(out), gu.PREC[prec],
                                          gu.BACK[backend], gu.stream()))
    torch.cuda.synchronize()
    scale = float(ref_nob.abs().max())
    assert float((gu.nchw(out).cpu() - ref_dz).abs().max()) <= gu.TOL[prec] * scale
    s2 = (ref_dz.double() * xhat.double()).sum(dim=(0, 2, 3))
    denom = float((ref_dz.double() * xhat.double()).abs().sum(dim=(0, 2, 3)).max())
    assert float((stats[cb:].cpu() - s2).abs().max()) <= gu.TOL[prec] * denom


@pytest.mark.parametrize("prec", gu.PRECISIONS)
@pytest.mark.parametrize("shape", [(300, 8, 64, 128), (80, 16, 32, 64), (150, 4, 128, 256), (301, 4, 128, 256)])
def test_rowgemm_second_generation_many_tiles(shape, prec, monkeypatch):
    """More tiles than SMs: every CTA of the second-generation kernel walks several tiles (double-buffered accumulators,
    ring wrap-around, resident weight groups re-used across tiles, the 128-wide forward tiles), ragged last tile included."""
    if "tc" not in gu.BACKENDS:
        pytest.skip("tcgen05 path only")
    monkeypatch.setenv("AE_B200_ROWGEMM", "2")
    test_conv_fwd(shape, _lib.OP_RAW, prec, "tc", "gen2")
    test_conv_dgrad_is_transposed_conv(shape, _lib.OP_BNRELU, prec, "tc", "gen2")


@pytest.mark.parametrize("backend", gu.BACKENDS)
@pytest.mark.parametrize("prec", gu.PRECISIONS)
@pytest.mark.parametrize("shape", SHAPES + [(64, 16, 32, 64)])
def test_conv_wgrad(shape, prec, backend):
    if backend == "simt" and prec == "bf16":
        pytest.skip("the CUDA-core backend always computes in fp32")
    b, hs, cb, cs = shape
    rs = np.random.RandomState(hash((shape, 2)) % 2 ** 31)
    d = gu.dev()
    big = torch.from_numpy(rs.standard_normal((b, cb, 2 * hs, 2 * hs)).astype(np.float32))
    dz = torch.from_numpy(rs.standard_normal((b, cs, hs, hs)).astype(np.float32))
    y = torch.from_numpy(rs.standard_normal((b, cs, hs, hs)).astype(np.float32))
    bnc_b, bnc_s = gu.make_bnc(cb, rs, "cpu"), gu.make_bnc(cs, rs, "cpu")
    a = _apply_operand_cpu(_lib.OP_BNRELU, big, None, bnc_b)
    dy = _apply_operand_cpu(_lib.OP_BNBWD, dz, y, bnc_s)
    w = torch.zeros(cs, cb, 3, 3, requires_grad=True)
    F.conv2d(a, w, None, stride=2, padding=1).backward(dy)
    ref = w.grad
    g = _geom(b, hs, cb, cs)
    nbytes = gu.lib().ae_conv2d_s2_wgrad_workspace_bytes(C.byref(g), gu.PREC[prec], gu.BACK[backend])
    part = torch.empty(nbytes + 16, dtype=torch.uint8, device=d)
    dw = torch.full((cs, cb, 3, 3), float("nan"), device=d)
    bigd, bncbd, dzd, yd, bncsd = gu.nhwc(big).to(d), bnc_b.to(d), gu.nhwc(dz).to(d), gu.nhwc(y).to(d), bnc_s.to(d)
    opb, _k1 = gu.conv_operand(backend, prec, cb, bigd, None, bncbd, _lib.OP_BNRELU)
    ops, _k2 = gu.conv_operand(backend, prec, cs, dzd, yd, bncsd, _lib.OP_BNBWD)
    _lib.check(gu.lib().ae_conv2d_s2_wgrad(C.byref(g), C.byref(opb), C.byref(ops), gu.p(dw), gu.p(part), nbytes,
                                          gu.PREC[prec], gu.BACK[backend], gu.stream()))
    torch.cuda.synchronize()
    assert gu.rel(dw.cpu(), ref) <= gu.TOL[prec]


def _decode_planes(planes, shape, prec):
    """split-bf16 planes [nsplit][...shape] (uint8 buffer) -> fp32 tensor hi (+ lo)"""
    ns = 2 if prec == "fp32" else 1
    t = planes.view(torch.bfloat16)[: ns * int(np.prod(shape))].view(ns, *shape).float()
    return t.sum(0)


@pytest.mark.parametrize("prec", gu.PRECISIONS)
@pytest.mark.parametrize("family", ["fwd", "dgrad"])
@pytest.mark.parametrize("shape", SHAPES)
def test_conv_fused_bnrelu_split_epilogue(shape, family, prec):
    """Eval-mode fusion (AE_EPI_BNRELU_SPLIT): the epilogue writes split-bf16 planes of relu(bn(conv + bias)), i.e. exactly
    what ae_split_operand(BNRELU) makes of the fp32 output of the two-pass form."""
    if "tc" not in gu.BACKENDS:
        pytest.skip("tcgen05 path only")
    b, hs, cb, cs = shape
    rs = np.random.RandomState(hash((shape, family)) % 2 ** 31)
    d = gu.dev()
    w = torch.from_numpy((rs.standard_normal((cs, cb, 3, 3)) / np.sqrt(9 * cb)).astype(np.float32))
    pf, pd = gu.pack_conv(w.to(d), cs, cb, prec, "tc")
    g = _geom(b, hs, cb, cs)
    if family == "fwd":
        cin, cout, hin, hout = cb, cs, 2 * hs, hs
    else:
        cin, cout, hin, hout = cs, cb, hs, 2 * hs
    a = torch.from_numpy(rs.standard_normal((b, cin, hin, hin)).astype(np.float32))
    bias = torch.from_numpy(rs.uniform(-0.1, 0.1, cout).astype(np.float32))
    bnc = gu.make_bnc(cout, rs, "cpu")
    ad, biasd, bncd = gu.nhwc(a).to(d), bias.to(d), bnc.to(d)
    op, _keep = gu.conv_operand("tc", prec, cin, ad)
    fn = gu.lib().ae_conv2d_s2_fwd if family == "fwd" else gu.lib().ae_conv2d_s2_dgrad
    pk = pf if family == "fwd" else pd
    # two-pass form: fp32 output, then BatchNorm + ReLU + split
    y = torch.empty(b, hout, hout, cout, device=d)
    ep = gu.epilogue(_lib.EPI_STORE, biasd)
    _lib.check(fn(C.byref(g), C.byref(op), gu.p(pk), C.byref(ep), gu.p(y), gu.PREC[prec], gu.BACK["tc"], gu.stream()))
    want = torch.empty(gu.lib().ae_split_operand_bytes(y.numel(), gu.PREC[prec]), dtype=torch.uint8, device=d)
    op2 = gu.operand(y, None, bncd, 0.0, _lib.OP_BNRELU)
    _lib.check(gu.lib().ae_split_operand(C.byref(op2), cout, y.numel(), gu.p(want), gu.PREC[prec], gu.stream()))
    # fused form
    got = torch.full_like(want, 0x7f)
    epf = gu.epilogue(_lib.EPI_BNRELU_SPLIT, biasd, None, bncd)
    _lib.check(fn(C.byref(g), C.byref(op), gu.p(pk), C.byref(epf), gu.p(got), gu.PREC[prec], gu.BACK["tc"], gu.stream()))
    torch.cuda.synchronize()
    # the training-mode launch (second-generation kernel) accumulates the taps in another order than the fused one: equal up
    # to fp32 rounding of the accumulator, i.e. the decoded planes agree to ~2^-17 (one bf16 ulp in bf16 mode)
    gv, wv = _decode_planes(got, (b, hout, hout, cout), prec), _decode_planes(want, (b, hout, hout, cout), prec)
    assert float((gv - wv).abs().max()) <= (2e-5 if prec == "fp32" else 8e-3) * float(wv.abs().max()), \
        "fused epilogue planes differ from the two-pass planes"
    # and against the CPU reference of the layer
    if family == "fwd":
        ref = F.conv2d(a, w, bias, stride=2, padding=1)
    else:
        ref = F.conv_transpose2d(a, w, bias, stride=2, padding=1, output_padding=1)
    ref = _apply_operand_cpu(_lib.OP_BNRELU, ref, None, bnc)
    assert gu.rel(gu.nchw(_decode_planes(got, (b, hout, hout, cout), prec)).cpu(), ref) <= 2 * gu.TOL[prec]


@pytest.mark.parametrize("prec", gu.PRECISIONS)
@pytest.mark.parametrize("family", ["fwd", "dgrad"])
@pytest.mark.parametrize("shape", [(3, 8, 64, 128), (5, 4, 128, 256), (1, 4, 128, 256), (37, 8, 64, 128)])
def test_conv_wide_n_tiles_match_narrow(shape, family, prec, monkeypatch):
    """The 128-wide n-tile variant of the row GEMM (chosen for thousands of m-tiles, i.e. inference batches) reads the
    same weight pack two 64-row n-tiles at a time; every output element accumulates the same products in the same order,
    so it must reproduce the 64-wide result bit for bit (and the CPU reference within tolerance)."""
    if "tc" not in gu.BACKENDS:
        pytest.skip("tcgen05 path only")
    b, hs, cb, cs = shape
    rs = np.random.RandomState(hash((shape, family, "wide")) % 2 ** 31)
    d = gu.dev()
    w = torch.from_numpy((rs.standard_normal((cs, cb, 3, 3)) / np.sqrt(9 * cb)).astype(np.float32))
    pf, pd = gu.pack_conv(w.to(d), cs, cb, prec, "tc")
    g = _geom(b, hs, cb, cs)
    if family == "fwd":
        cin, cout, hin, hout = cb, cs, 2 * hs, hs
    else:
        cin, cout, hin, hout = cs, cb, hs, 2 * hs
    a = torch.from_numpy(rs.standard_normal((b, cin, hin, hin)).astype(np.float32))
    bias = torch.from_numpy(rs.uniform(-0.1, 0.1, cout).astype(np.float32))
    ad, biasd = gu.nhwc(a).to(d), bias.to(d)
    op, _keep = gu.conv_operand("tc", prec, cin, ad)
    fn = gu.lib().ae_conv2d_s2_fwd if family == "fwd" else gu.lib().ae_conv2d_s2_dgrad
    pk = pf if family == "fwd" else pd
    outs, stats = [], []
    for force in ("-1", "1"):
        monkeypatch.setenv("AE_B200_FORCE_NT128", force)
        y = torch.full((b, hout, hout, cout), float("nan"), device=d)
        st = torch.zeros(2 * cout, dtype=torch.float64, device=d)
        ep = gu.epilogue(_lib.EPI_BIAS_STATS, biasd, None, None, st)
        _lib.check(fn(C.byref(g), C.byref(op), gu.p(pk), C.byref(ep), gu.p(y), gu.PREC[prec], gu.BACK["tc"], gu.stream()))
        torch.cuda.synchronize()
        outs.append(y)
        stats.append(st)
    assert torch.equal(outs[0], outs[1])
    assert gu.rel(stats[1], stats[0]) <= 1e-6
    if family == "fwd":
        ref = F.conv2d(a, w, bias, stride=2, padding=1)
    else:
        ref = F.conv_transpose2d(a, w, bias, stride=2, padding=1, output_padding=1)
    assert gu.rel(gu.nchw(outs[1]).cpu(), ref) <= gu.TOL[prec]


@pytest.mark.parametrize("prec", gu.PRECISIONS)
@pytest.mark.parametrize("batch", [1, 5])
def test_thin_gather_fused_bnrelu_split_epilogue(batch, prec):
    """Conv2d(3,32) with the eval-mode epilogue: planes of relu(bn(conv1(x) + b)) bit-identical to the two-pass form."""
    rs = np.random.RandomState(300 + batch)
    d = gu.dev()
    x = torch.from_numpy(rs.random_sample((batch, 3, 64, 64)).astype(np.float32))
    w1 = torch.from_numpy((rs.standard_normal((32, 3, 3, 3)) / 5).astype(np.float32))
    b1 = torch.from_numpy(rs.uniform(-0.1, 0.1, 32).astype(np.float32))
    bnc = gu.make_bnc(32, rs, "cpu")
    xd, w1d, b1d, bncd = x.to(d), w1.to(d), b1.to(d), bnc.to(d)
    pb = (gu.PREC[prec], gu.BACK["simt"])
    op = gu.operand(xd)
    y = torch.empty(batch, 32, 32, 32, device=d)
    ep = gu.epilogue(_lib.EPI_STORE, b1d)
    _lib.check(gu.lib().ae_thin_gather_fwd(C.byref(op), gu.p(w1d), C.byref(ep), gu.p(y), batch, *pb, gu.stream()))
    want = torch.empty(gu.lib().ae_split_operand_bytes(y.numel(), gu.PREC[prec]), dtype=torch.uint8, device=d)
    op2 = gu.operand(y, None, bncd, 0.0, _lib.OP_BNRELU)
    _lib.check(gu.lib().ae_split_operand(C.byref(op2), 32, y.numel(), gu.p(want), gu.PREC[prec], gu.stream()))
    got = torch.full_like(want, 0x7f)
    epf = gu.epilogue(_lib.EPI_BNRELU_SPLIT, b1d, None, bncd)
    _lib.check(gu.lib().ae_thin_gather_fwd(C.byref(op), gu.p(w1d), C.byref(epf), gu.p(got), batch, *pb, gu.stream()))
    torch.cuda.synchronize()
    assert torch.equal(got, want)
    ref = _apply_operand_cpu(_lib.OP_BNRELU, F.conv2d(x, w1, b1, stride=2, padding=1), None, bnc)
    assert gu.rel(gu.nchw(_decode_planes(got, (batch, 32, 32, 32), prec)).cpu(), ref) <= gu.TOL[prec]


@pytest.mark.parametrize("backend", gu.BACKENDS)
@pytest.mark.parametrize("batch", [1, 3, 16])
def test_thin_layers(batch, backend):
    # CUDA cores: fp32 arithmetic.  tcgen05: 2-term bf16 operand split, the fp32-mode tolerance of the GEMM kernels.
    tol = 1e-5 if backend == "simt" else gu.TOL["fp32"]
    pb = (gu.PREC["fp32"], gu.BACK[backend])
    rs = np.random.RandomState(100 + batch)
    d = gu.dev()
    x = torch.from_numpy(rs.random_sample((batch, 3, 64, 64)).astype(np.float32))
    w1 = torch.from_numpy((rs.standard_normal((32, 3, 3, 3)) / 5).astype(np.float32))
    b1 = torch.from_numpy(rs.uniform(-0.1, 0.1, 32).astype(np.float32))
    # conv1 forward + statistics
    ref = F.conv2d(x, w1, b1, stride=2, padding=1)
    out = torch.empty(batch, 32, 32, 32, device=d)
    stats = torch.zeros(64, dtype=torch.float64, device=d)
    xd, w1d, b1d = x.to(d), w1.to(d), b1.to(d)
    op = gu.operand(xd)
    ep = gu.epilogue(_lib.EPI_BIAS_STATS, b1d, None, None, stats)
    _lib.check(gu.lib().ae_thin_gather_fwd(C.byref(op), gu.p(w1d), C.byref(ep), gu.p(out), batch, *pb, gu.stream()))
    torch.cuda.synchronize()
    assert gu.rel(gu.nchw(out).cpu(), ref) <= tol
    assert gu.rel(stats[:32], ref.double().sum(dim=(0, 2, 3))) <= tol
    assert gu.rel(stats[32:], (ref.double() ** 2).sum(dim=(0, 2, 3))) <= tol
    # convT4 forward + sigmoid + squared error
    t3 = torch.from_numpy(rs.standard_normal((batch, 32, 32, 32)).astype(np.float32))
    bnc = gu.make_bnc(32, rs, "cpu")
    a3 = _apply_operand_cpu(_lib.OP_BNRELU, t3, None, bnc)
    w4 = torch.from_numpy((rs.standard_normal((32, 3, 3, 3)) / 8).astype(np.float32))
    b4 = torch.from_numpy(rs.uniform(-0.1, 0.1, 3).astype(np.float32))
    ref_x = torch.sigmoid(F.conv_transpose2d(a3, w4, b4, stride=2, padding=1, output_padding=1))
    xh = torch.empty(batch, 3, 64, 64, device=d)
    sse = torch.zeros(2, dtype=torch.float64, device=d)
    t3d, bncd, w4d, b4d = gu.nhwc(t3).to(d), bnc.to(d), w4.to(d), b4.to(d)
    opw = gu.operand(t3d, None, bncd, 0.0, _lib.OP_BNRELU)
    _lib.check(gu.lib().ae_thin_scatter_sigmoid_fwd(C.byref(opw), gu.p(w4d), gu.p(b4d), gu.p(xh), gu.p(xd), gu.p(sse),
                                                   batch, *pb, gu.stream()))
    torch.cuda.synchronize()
    assert gu.rel(xh.cpu(), ref_x) <= tol
    assert abs(float(sse[0]) - float(((ref_x - x).double() ** 2).sum())) <= tol * float(((ref_x - x).double() ** 2).sum())
    # convT4 backward: fused-MSE upstream gradient, weight / bias gradient, data gradient with ReLU mask
    a3r = a3.clone().requires_grad_(True)
    w4r, b4r = w4.clone().requires_grad_(True), b4.clone().requires_grad_(True)
    xr = torch.sigmoid(F.conv_transpose2d(a3r, w4r, b4r, stride=2, padding=1, output_padding=1))
    alpha = 35.0
    (alpha * F.mse_loss(xr, x)).backward()
    scal = 2.0 * alpha / x.numel()
    opt = gu.operand(xd, xh, None, scal, _lib.OP_SIGMOID_BWD)
    nb = gu.lib().ae_thin_wgrad_workspace_bytes(batch)
    part = torch.empty(nb, dtype=torch.uint8, device=d)
    dw = torch.empty(32, 3, 3, 3, device=d)
    db = torch.empty(3, device=d)
    _lib.check(gu.lib().ae_thin_wgrad(C.byref(opw), C.byref(opt), gu.p(dw), gu.p(db), gu.p(part), nb, batch, *pb, gu.stream()))
    dz = torch.empty(batch, 32, 32, 32, device=d)
    stats.zero_()
    epb = gu.epilogue(_lib.EPI_RELUBWD_STATS, None, t3d, bncd, stats)
    _lib.check(gu.lib().ae_thin_gather_fwd(C.byref(opt), gu.p(w4d), C.byref(epb), gu.p(dz), batch, *pb, gu.stream()))
    torch.cuda.synchronize()
    assert gu.rel(dw.cpu(), w4r.grad) <= 2 * tol
    assert gu.rel(db.cpu(), b4r.grad) <= 2 * tol
    v = lambda r: bnc[r].view(1, -1, 1, 1)
    mask = (t3 * v(0) + v(1)) > 0
    ref_dz = a3r.grad * mask
    assert float((gu.nchw(dz).cpu() - ref_dz).abs().max()) <= 2 * tol * float(a3r.grad.abs().max())
    # the fused backward (one pass over x / x_hat) must reproduce both results
    dw2, db2, dz2 = torch.empty_like(dw), torch.empty_like(db), torch.empty_like(dz)
    stats2 = torch.zeros_like(stats)
    epf = gu.epilogue(_lib.EPI_RELUBWD_STATS, None, t3d, bncd, stats2)
    _lib.check(gu.lib().ae_thin_bwd_fused(C.byref(opw), C.byref(opt), gu.p(w4d), C.byref(epf), gu.p(dz2), gu.p(dw2), gu.p(db2),
                                         gu.p(part), nb, batch, *pb, gu.stream()))
    torch.cuda.synchronize()
    assert gu.rel(dw2, dw) <= 1e-6 and gu.rel(db2, db) <= 1e-6 and gu.rel(dz2, dz) <= 1e-6
    assert gu.rel(stats2, stats) <= 1e-6


def test_bn_finalize_and_backward_coefficients():
    rs = np.random.RandomState(5)
    d = gu.dev()
    Cn, count = 64, 6 * 16 * 16
    y = torch.from_numpy((rs.standard_normal((6, Cn, 16, 16)) * 2 + 0.5).astype(np.float32))
    gamma = torch.from_numpy(rs.uniform(0.5, 1.5, Cn).astype(np.float32))
    beta = torch.from_numpy(rs.uniform(-0.5, 0.5, Cn).astype(np.float32))
    rm = torch.from_numpy(rs.uniform(-0.5, 0.5, Cn).astype(np.float32))
    rv = torch.from_numpy(rs.uniform(0.5, 1.5, Cn).astype(np.float32))
    yr = y.clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    rm_ref, rv_ref = rm.clone(), rv.clone()
    out = F.relu(F.batch_norm(yr, rm_ref, rv_ref, gr, br, True, 0.1, 1e-5))
    dup = torch.from_numpy(rs.standard_normal(out.shape).astype(np.float32))
    out.backward(dup)
    stats = torch.stack([y.double().sum(dim=(0, 2, 3)), (y.double() ** 2).sum(dim=(0, 2, 3))]).reshape(-1).to(d)
    bnc = torch.zeros(8, Cn, device=d)
    gd, bd, rmd, rvd = gamma.to(d), beta.to(d), rm.to(d), rv.to(d)
    _lib.check(gu.lib().ae_bn_finalize(gu.p(stats), count, gu.p(gd), gu.p(bd), gu.p(rmd), gu.p(rvd), gu.p(bnc), Cn, 1, gu.stream()))
    torch.cuda.synchronize()
    assert gu.rel(rmd, rm_ref) <= 1e-6 and gu.rel(rvd, rv_ref) <= 1e-6
    act = torch.relu(y * bnc[0].cpu().view(1, -1, 1, 1) + bnc[1].cpu().view(1, -1, 1, 1))
    assert gu.rel(act, out) <= 1e-5
    # backward: dz = dup * mask; S1 = sum dz; S2 = sum dz*xhat
    mask = out > 0
    dz = dup * mask
    xhat = (y - bnc[2].cpu().view(1, -1, 1, 1)) * bnc[3].cpu().view(1, -1, 1, 1)
    st2 = torch.stack([dz.double().sum(dim=(0, 2, 3)), (dz.double() * xhat.double()).sum(dim=(0, 2, 3))]).reshape(-1).to(d)
    dg, db = torch.empty(Cn, device=d), torch.empty(Cn, device=d)
    _lib.check(gu.lib().ae_bn_bwd_reduce(gu.p(st2), count, gu.p(gd), gu.p(bnc), gu.p(dg), gu.p(db), Cn, gu.stream()))
    torch.cuda.synchronize()
    assert gu.rel(dg, gr.grad) <= 1e-5 and gu.rel(db, br.grad) <= 1e-5
    b = bnc.cpu()
    dy = b[4].view(1, -1, 1, 1) * dz + b[5].view(1, -1, 1, 1) * (y - b[2].view(1, -1, 1, 1)) + b[6].view(1, -1, 1, 1)
    assert gu.rel(dy, yr.grad) <= 2e-5
    # eval mode: coefficients from the running statistics
    bnc2 = torch.zeros(8, Cn, device=d)
    _lib.check(gu.lib().ae_bn_finalize(None, count, gu.p(gd), gu.p(bd), gu.p(rmd), gu.p(rvd), gu.p(bnc2), Cn, 0, gu.stream()))
    torch.cuda.synchronize()
    ref_eval = F.batch_norm(y, rm_ref, rv_ref, gamma, beta, False, 0.1, 1e-5)
    got_eval = y * bnc2[0].cpu().view(1, -1, 1, 1) + bnc2[1].cpu().view(1, -1, 1, 1)
    assert gu.rel(got_eval, ref_eval) <= 1e-5


@pytest.mark.parametrize("wd", [0.0, 1e-4])
def test_adam_matches_torch(wd):
    rs = np.random.RandomState(9)
    d = gu.dev()
    n = 4096 + 8
    p0 = torch.from_numpy(rs.standard_normal(n).astype(np.float32))
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref], lr=5e-3, weight_decay=wd)
    pd, m, v = p0.clone().to(d), torch.zeros(n, device=d), torch.zeros(n, device=d)
    step = torch.zeros(2, dtype=torch.int32, device=d)
    for it in range(5):
        g = torch.from_numpy((rs.standard_normal(n) * 10.0 ** rs.uniform(-6, 0, n)).astype(np.float32))
        ref.grad = g.clone()
        opt.step()
        gd = g.to(d)
        _lib.check(gu.lib().ae_adam_step_flat(gu.p(pd), gu.p(gd), gu.p(m), gu.p(v), n, 5e-3, 0.9, 0.999, 1e-8, wd, 1.0,
                                             gu.p(step), gu.stream()))
    torch.cuda.synchronize()
    assert int(step[0]) == 5 and int(step[1]) == 0
    assert float((pd.cpu() - ref.detach()).abs().max()) <= 2e-6


@pytest.mark.parametrize("batch", [1, 7, 256])
def test_softmax_ce(batch):
    rs = np.random.RandomState(batch)
    d = gu.dev()
    logits = torch.from_numpy((rs.standard_normal((batch, 10)) * 3).astype(np.float32))
    labels = torch.from_numpy(rs.randint(0, 10, batch).astype(np.int64))
    lr = logits.clone().requires_grad_(True)
    ref = F.cross_entropy(lr, labels)
    ref.backward()
    ld, yd = logits.to(d), labels.to(d)
    loss = torch.zeros(4, device=d)
    dl = torch.empty(batch, 10, device=d)
    corr = torch.zeros(1, dtype=torch.int32, device=d)
    _lib.check(gu.lib().ae_softmax_ce_fwd_bwd(gu.p(ld), gu.p(yd), batch, 10, 1.0, gu.p(loss), gu.p(dl), gu.p(corr), gu.stream()))
    torch.cuda.synchronize()
    assert abs(float(loss[0]) - float(ref)) <= 2e-6 * max(1.0, abs(float(ref)))
    assert gu.rel(dl, lr.grad) <= 1e-5
    assert int(corr[0]) == int((logits.argmax(1) == labels).sum())


def test_layout_roundtrip():
    d = gu.dev()
    x = torch.randn(3, 5, 6, 7, device=d)
    y = torch.empty(3, 6, 7, 5, device=d)
    _lib.check(gu.lib().ae_layout_nchw_f32_to_nhwc_f32(gu.p(x), gu.p(y), 3, 5, 6, 7, gu.stream()))
    assert torch.equal(y, x.permute(0, 2, 3, 1).contiguous())
    z = torch.empty_like(x)
    _lib.check(gu.lib().ae_layout_nhwc_f32_to_nchw_f32(gu.p(y), gu.p(z), 3, 5, 6, 7, gu.stream()))
    assert torch.equal(z, x)
    yb = torch.empty(3, 6, 7, 5, dtype=torch.bfloat16, device=d)
    _lib.check(gu.lib().ae_layout_nchw_f32_to_nhwc_bf16(gu.p(x), gu.p(yb), 3, 5, 6, 7, gu.stream()))
    assert torch.equal(yb, x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16))
    zb = torch.empty_like(x)
    _lib.check(gu.lib().ae_layout_nhwc_bf16_to_nchw_f32(gu.p(yb), gu.p(zb), 3, 5, 6, 7, gu.stream()))
    assert torch.equal(zb, yb.float().permute(0, 3, 1, 2).contiguous())


def test_linear_fwd_bwd_with_flatten_permutation():
    """enc.encoder.13 (NB:520-521): Flatten in (C,H,W) order + Linear, activation held NHWC."""
    rs = np.random.RandomState(77)
    d = gu.dev()
    B, L, K = 5, 64, 4096
    y4 = torch.from_numpy(rs.standard_normal((B, 256, 4, 4)).astype(np.float32))
    bnc = gu.make_bnc(256, rs, "cpu")
    w = torch.from_numpy((rs.standard_normal((L, K)) / 64).astype(np.float32))
    bias = torch.from_numpy(rs.uniform(-0.1, 0.1, L).astype(np.float32))
    a4 = _apply_operand_cpu(_lib.OP_BNRELU, y4, None, bnc).requires_grad_(True)
    wr, br = w.clone().requires_grad_(True), bias.clone().requires_grad_(True)
    z = F.linear(a4.flatten(1), wr, br)
    dz = torch.from_numpy(rs.standard_normal((B, L)).astype(np.float32))
    z.backward(dz)
    nws = gu.lib().ae_linear_workspace_bytes(B, L, K)
    ws = torch.empty(nws + 256, dtype=torch.uint8, device=d)
    wsp = C.c_void_p((ws.data_ptr() + 255) & ~255)
    y4d, bncd, wdv, bd, dzd = gu.nhwc(y4).to(d), bnc.to(d), w.to(d), bias.to(d), dz.to(d)
    op = gu.operand(y4d, None, bncd, 0.0, _lib.OP_BNRELU)
    out = torch.empty(B, L, device=d)
    _lib.check(gu.lib().ae_linear_fwd(C.byref(op), 256, gu.p(wdv), gu.p(bd), gu.p(out), B, L, K, 1, 0, wsp, nws, gu.stream()))
    torch.cuda.synchronize()
    assert gu.rel(out, z) <= 1e-5
    da = torch.empty(B, 4, 4, 256, device=d)
    dw = torch.empty(L, K, device=d)
    db = torch.empty(L, device=d)
    _lib.check(gu.lib().ae_linear_bwd(C.byref(op), 256, gu.p(wdv), gu.p(dzd), gu.p(da), None, 256, gu.p(dw), gu.p(db), B, L, K,
                                     1, 0, wsp, nws, gu.stream()))
    torch.cuda.synchronize()
    assert gu.rel(gu.nchw(da).cpu(), a4.grad) <= 1e-5
    assert gu.rel(dw, wr.grad) <= 1e-5
    assert gu.rel(db, br.grad) <= 1e-5


def _guarded(shape, dtype=torch.float32):
    """Tensor of `shape` inside a larger allocation whose 16 KB before and after it hold a sentinel: compute-sanitizer is
    closed on this pool (gpurun refuses it), so out-of-bounds writes of the TMA stores / epilogues are looked for this way."""
    n = int(np.prod(shape))
    itemsize = torch.empty(0, dtype=dtype).element_size()
    guard = 16384 // itemsize
    buf = torch.full((n + 2 * guard,), 7, dtype=torch.uint8 if dtype == torch.uint8 else dtype, device=gu.dev())
    if dtype != torch.uint8:
        buf.fill_(-12345.0)
    t = buf[guard:guard + n].view(*shape)

    def intact():
        ref = 7 if dtype == torch.uint8 else -12345.0
        return bool((buf[:guard] == ref).all()) and bool((buf[guard + n:] == ref).all())
    return t, intact


@pytest.mark.parametrize("gen", ["1", "2"])
@pytest.mark.parametrize("prec", gu.PRECISIONS)
@pytest.mark.parametrize("shape", [(3, 8, 64, 128), (5, 4, 128, 256), (33, 16, 32, 64), (1, 4, 128, 256), (150, 4, 128, 256)])
def test_rowgemm_writes_stay_inside_the_output(shape, prec, gen, monkeypatch):
    """Ragged last tiles (M not a multiple of 128 or 32 rows): the tensor-map stores of the second generation must clip, the
    per-lane stores of the first generation must be masked; every valid element must be written."""
    if "tc" not in gu.BACKENDS:
        pytest.skip("tcgen05 path only")
    monkeypatch.setenv("AE_B200_ROWGEMM", gen)
    b, hs, cb, cs = shape
    d = gu.dev()
    w = torch.randn(cs, cb, 3, 3, device=d) * 0.05
    pf, pd = gu.pack_conv(w, cs, cb, prec, "tc")
    g = _geom(b, hs, cb, cs)
    for fam in ("fwd", "dgrad"):
        cin, hin, oshape, oc = (cb, 2 * hs, (b, hs, hs, cs), cs) if fam == "fwd" else (cs, hs, (b, 2 * hs, 2 * hs, cb), cb)
        a = torch.randn(b, hin, hin, cin, device=d)
        op, _keep = gu.conv_operand("tc", prec, cin, a)
        fn = gu.lib().ae_conv2d_s2_fwd if fam == "fwd" else gu.lib().ae_conv2d_s2_dgrad
        out, intact = _guarded(oshape)
        bias = torch.zeros(oc, device=d)
        stats = torch.zeros(2 * oc, dtype=torch.float64, device=d)
        ep = gu.epilogue(_lib.EPI_BIAS_STATS, bias, None, None, stats)
        _lib.check(fn(C.byref(g), C.byref(op), gu.p(pf if fam == "fwd" else pd), C.byref(ep), gu.p(out), gu.PREC[prec], gu.BACK["tc"], gu.stream()))
        torch.cuda.synchronize()
        assert intact(), f"{fam}: wrote outside the output tensor"
        assert not bool((out == -12345.0).any()), f"{fam}: left output elements unwritten"
