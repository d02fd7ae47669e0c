#!/usr/bin/env python
"""Developer tool: where does a row-GEMM launch spend its time?  Needs the trace build of the library
(AE_B200_BUILD_VARIANT=trace python <pkg>/build.py  ->  libae_b200_trace.so, selected with AE_B200_LIB): every CTA of
k_tma_rowgemm writes %globaltimer at fixed points (kernel entry, set-up done, first / last TMA issued, first operands
landed, last MMA committed, first / last accumulator ready, last tile stored, statistics flushed, exit).  Prints, per
geometry, the launch as a timeline relative to the earliest CTA's entry.

    AE_B200_LIB=$PWD/<pkg>/libae_b200_trace.so python scripts/trace_rowgemm.py
"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from ae_b200 import _lib  # noqa: E402

SLOTS = ["entry", "setup", "tma_first", "tma_last", "ops_landed", "mma_last", "acc_first", "acc_last", "stored", "flushed", "exit",
         "epi_ld0", "epi_st0", "epi_stat0"]


def run(family, B, hs, cb, cs, epi_mode, prec="fp32", reps=5):
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
    if family == "dgrad":       # small -> big
        a_elems, o_shape, oc = Ms * cs, (B, 2 * hs, 2 * hs, cb), cb
    else:                       # big -> small
        a_elems, o_shape, oc = 4 * Ms * cb, (B, hs, hs, cs), cs
    planes = (torch.randn(nsplit * a_elems, device=dev) * 0.5).to(torch.bfloat16)
    out = torch.empty(o_shape, device=dev)
    bias = torch.zeros(oc, device=dev)
    stats = torch.zeros(2 * oc, dtype=torch.float64, device=dev)
    y = torch.randn(o_shape, device=dev)
    bnc = torch.ones(_lib.BNC_ROWS * oc, device=dev)
    if epi_mode == "bias_stats":
        ep = _lib.Epilogue(_lib.EPI_BIAS_STATS, _lib.ptr(bias), None, None, _lib.ptr(stats))
    elif epi_mode == "relubwd":
        ep = _lib.Epilogue(_lib.EPI_RELUBWD_STATS, None, _lib.ptr(y), _lib.ptr(bnc), _lib.ptr(stats))
    else:
        ep = _lib.Epilogue(_lib.EPI_STORE, _lib.ptr(bias), None, None, None)
    op = _lib.Operand(_lib.ptr(planes), None, None, 0.0, _lib.OP_SPLIT_BF16)
    fn = lib.ae_conv2d_s2_dgrad if family == "dgrad" else lib.ae_conv2d_s2_fwd
    pk = pk_d if family == "dgrad" else pk_f
    trace = torch.zeros(1024 * 16 + 512, dtype=torch.int64, device=dev)
    for fnn in ("ae_debug_set_trace", "ae_debug_set_trace2"):       # first / second generation kernel: only one of them runs
        getattr(lib, fnn).argtypes = [C.c_void_p]
        getattr(lib, fnn).restype = C.c_int
    flush = torch.empty(160 * 1024 * 1024 // 4, device=dev)
    for cold in (False, True):
        rows = []
        for _ in range(reps):
            if cold:
                flush.zero_()               # > L2: operands and weights come from DRAM
            trace.zero_()
            torch.cuda.synchronize()
            assert lib.ae_debug_set_trace(C.c_void_p(trace.data_ptr())) == 0
            assert lib.ae_debug_set_trace2(C.c_void_p(trace.data_ptr())) == 0
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.check(fn(C.byref(g), C.byref(op), pk, C.byref(ep), _lib.ptr(out), P, _lib.BACKEND_TC, _lib.stream_ptr()))
            e1.record()
            torch.cuda.synchronize()
            t = trace[:16384].view(-1, 16).cpu()
            t = t[t[:, 0] > 0]
            items = trace[16384:16384 + 240].view(-1, 4).cpu()
            prod = trace[16384 + 256:16384 + 256 + 240].view(-1, 4).cpu()
            rows.append((e0.elapsed_time(e1) * 1e3, t))
        us, t = rows[-1]
        t0 = int(t[:, 0].min())
        rel = (t[:, :len(SLOTS)] - t0).double() / 1e3
        print(f"--- {family} B={B} hs={hs} cb={cb} cs={cs} epi={epi_mode} {prec} {'L2-flushed' if cold else 'warm'}: "
              f"{t.shape[0]} CTAs, event time {us:.1f} us (all reps: {', '.join(f'{r[0]:.1f}' for r in rows)})")
        print("    slot          min     median     max   (us after the first CTA's entry)")
        for i, name in enumerate(SLOTS):
            col = rel[:, i]
            col = col[t[:, i] > 0]
            if col.numel():
                print(f"    {name:11s} {float(col.min()):7.2f} {float(col.median()):9.2f} {float(col.max()):7.2f}")
        it, pr = items.view(-1), prod.view(-1)
        if int(it[3]) > 0:      # second-generation kernel, CTA 0: cycle accounting of the two single-thread roles
            mhz = float(it[3]) / max(float(it[4]), 1.0) * 1e3
            n = max(int(it[5]), 1)
            print(f"    CTA0 MMA thread: {int(it[5])} A items, {int(it[3])} clk total at {mhz:.0f} MHz; per item: wait A {int(it[0]) / n:.0f} clk, "
                  f"wait W {int(it[1]) / n:.0f}, issue {int(it[2]) / n:.0f}, other {(int(it[3]) - int(it[0]) - int(it[1]) - int(it[2])) / n:.0f}")
            print(f"    CTA0 producer: {int(pr[2])} clk total; per item: wait A slot {int(pr[0]) / n:.0f} clk, wait W slot {int(pr[1]) / n:.0f}, "
                  f"other {(int(pr[2]) - int(pr[0]) - int(pr[1])) / n:.0f}")
        assert lib.ae_debug_set_trace(None) == 0 and lib.ae_debug_set_trace2(None) == 0


if __name__ == "__main__":
    if "trace" not in os.environ.get("AE_B200_LIB", ""):
        raise SystemExit("set AE_B200_LIB to the trace build (see the docstring)")
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    which = sys.argv[2] if len(sys.argv) > 2 else "all"
    run("dgrad", B, 8, 64, 128, "bias_stats")       # ConvTranspose2d 128->64 forward (the roofline kernel)
    run("fprop", B, 16, 32, 64, "bias_stats")       # Conv2d 32->64 forward
    run("fprop", B, 4, 128, 256, "bias_stats")      # Conv2d 128->256 forward
    if which == "all":
        run("dgrad", B, 8, 64, 128, "relubwd")          # Conv2d 64->128 data gradient
        run("dgrad", B, 8, 64, 128, "store")
        run("dgrad", B, 16, 32, 64, "bias_stats")       # ConvTranspose2d 64->32 forward
        run("dgrad", B, 4, 128, 256, "bias_stats")      # ConvTranspose2d 256->128 forward
        run("fprop", B, 8, 64, 128, "bias_stats")       # Conv2d 64->128 forward
        run("fprop", B, 16, 32, 64, "bias_stats")       # Conv2d 32->64 forward
        run("dgrad", B, 8, 64, 128, "bias_stats", prec="bf16")
