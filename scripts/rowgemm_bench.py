#!/usr/bin/env python
"""Developer tool: device time of every row-GEMM launch of one training step (batch B), first generation
(AE_B200_ROWGEMM=1) against second generation, each timed as a CUDA graph of back-to-back launches rotating over
more buffers than fit in L2 (host-side launch cost -- tensor-map encoding -- is outside the graph replay).

    python scripts/rowgemm_bench.py [B]
"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from ae_b200 import _lib  # noqa: E402


def time_graph(launch, n_rot, reps=10):
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for i in range(n_rot):
            launch(i, st)
        st.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            for i in range(n_rot):
                launch(i, st)
        g.replay()
        st.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(reps):
            g.replay()
        e1.record(st)
        st.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * n_rot)


def case(family, B, hs, cb, cs, epi_mode, prec="fp32"):
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    P = _lib.PREC_FP32 if prec == "fp32" else _lib.PREC_BF16
    nsplit = 2 if prec == "fp32" else 1
    g = _lib.ConvGeom(B, hs, hs, cb, cs)
    w = torch.randn(cs, cb, 3, 3, device=dev) * 0.05
    nbytes = lib.ae_packed_weight_bytes(cs, cb, P, _lib.BACKEND_TC)
    raw = torch.zeros(2 * nbytes + 2048, dtype=torch.uint8, device=dev)
    base = (raw.data_ptr() + 1023) & ~1023
    pk_f, pk_d = C.c_void_p(base), C.c_void_p((base + nbytes + 1023) & ~1023)
    _lib.check(lib.ae_pack_conv_weight(_lib.ptr(w), cs, cb, pk_f, pk_d, P, _lib.BACKEND_TC, _lib.stream_ptr()))
    Ms = B * hs * hs
    if family == "dgrad":
        a_elems, o_shape, oc = Ms * cs, (B, 2 * hs, 2 * hs, cb), cb
    else:
        a_elems, o_shape, oc = 4 * Ms * cb, (B, hs, hs, cs), cs
    o_elems = o_shape[0] * o_shape[1] * o_shape[2] * o_shape[3]
    per = a_elems * 2 * nsplit + o_elems * 4 * (2 if epi_mode == "relubwd" else 1)
    n_rot = max(3, int(150e6 // per) + 1)
    planes = [(torch.randn(nsplit * a_elems, device=dev) * 0.5).to(torch.bfloat16) for _ in range(n_rot)]
    outs = [torch.empty(o_shape, device=dev) for _ in range(n_rot)]
    ys = [torch.randn(o_shape, device=dev) for _ in range(n_rot)] if epi_mode == "relubwd" else None
    bias = torch.zeros(oc, device=dev)
    stats = torch.zeros(2 * oc, dtype=torch.float64, device=dev)
    bnc = torch.ones(_lib.BNC_ROWS * oc, device=dev)
    fn = lib.ae_conv2d_s2_dgrad if family == "dgrad" else lib.ae_conv2d_s2_fwd
    pk = pk_d if family == "dgrad" else pk_f

    def launch(i, st):
        if epi_mode == "bias_stats":
            ep = _lib.Epilogue(_lib.EPI_BIAS_STATS, _lib.ptr(bias), None, None, _lib.ptr(stats))
        elif epi_mode == "relubwd":
            ep = _lib.Epilogue(_lib.EPI_RELUBWD_STATS, None, _lib.ptr(ys[i]), _lib.ptr(bnc), _lib.ptr(stats))
        else:
            ep = _lib.Epilogue(_lib.EPI_STORE, _lib.ptr(bias), None, None, None)
        op = _lib.Operand(_lib.ptr(planes[i]), None, None, 0.0, _lib.OP_SPLIT_BF16)
        _lib.check(fn(C.byref(g), C.byref(op), pk, C.byref(ep), _lib.ptr(outs[i]), P, _lib.BACKEND_TC, C.c_void_p(st.cuda_stream)))

    res = []
    for gen in ("1", "2"):
        os.environ["AE_B200_ROWGEMM"] = gen
        res.append(time_graph(launch, n_rot))
    os.environ.pop("AE_B200_ROWGEMM")
    flop = 2.0 * Ms * 9 * cs * cb
    print(f"{family:6s} hs={hs:2d} cb={cb:3d} cs={cs:3d} {epi_mode:10s} {prec}: gen1 {res[0]:6.2f} us   gen2 {res[1]:6.2f} us   "
          f"({flop / res[1] / 1e6:6.1f} TFLOP/s useful, rot {n_rot})")
    return res


if __name__ == "__main__":
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    tot = [0.0, 0.0]
    # encoder forward (FPROP, bias+stats), decoder forward (DGRAD, bias+stats), decoder backward (FPROP, relu-bwd; the
    # first decoder layer's input gradient is a plain store), encoder backward (DGRAD, relu-bwd)
    for fam, hs, cb, cs, ep in (("fprop", 16, 32, 64, "bias_stats"), ("fprop", 8, 64, 128, "bias_stats"), ("fprop", 4, 128, 256, "bias_stats"),
                                ("dgrad", 4, 128, 256, "bias_stats"), ("dgrad", 8, 64, 128, "bias_stats"), ("dgrad", 16, 32, 64, "bias_stats"),
                                ("fprop", 16, 32, 64, "relubwd"), ("fprop", 8, 64, 128, "relubwd"), ("fprop", 4, 128, 256, "store"),
                                ("dgrad", 4, 128, 256, "relubwd"), ("dgrad", 8, 64, 128, "relubwd"), ("dgrad", 16, 32, 64, "relubwd")):
        r = case(fam, B, hs, cb, cs, ep)
        tot[0] += r[0]; tot[1] += r[1]
    print(f"sum over the 12 launches of a step: gen1 {tot[0]:.1f} us, gen2 {tot[1]:.1f} us")
    if len(sys.argv) > 2:
        for fam, hs, cb, cs, ep in (("dgrad", 8, 64, 128, "bias_stats"), ("fprop", 8, 64, 128, "bias_stats")):
            case(fam, B, hs, cb, cs, ep, prec="bf16")
