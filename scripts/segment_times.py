#!/usr/bin/env python
"""Developer tool: where the step's time goes, by segment.  Each segment (encoder forward, decoder forward, decoder
backward, encoder backward, Adam + re-pack) is captured into its own CUDA graph through the public C ABI and replayed
between CUDA events, so the numbers are device times without host launch cost.  The whole-step graph (TrainStep) is
printed beside the sum: the difference is what the step's parallel branches and fusions (head beside the decoder, fused
MSE, conv1's weight gradient beside Adam) are worth.

    python scripts/segment_times.py [batch]
"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import ae_b200  # noqa: E402
from ae_b200 import _lib  # noqa: E402
from ae_b200._lib import check, ptr  # noqa: E402


def timed_graph(fn, st, reps=30):
    with torch.cuda.stream(st):
        fn()
        st.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            fn()
        for _ in range(3):
            g.replay()
        st.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(reps):
            g.replay()
        e1.record(st)
        st.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    dev = torch.device("cuda", 0)
    lib = _lib.load()
    torch.manual_seed(0)
    model = ae_b200.SupervisedAutoencoder(64, 10).to(dev).train()
    eng = model.engine()
    eng.prepare(dev, B)
    h = eng.handle
    x = torch.rand(B, 3, 64, 64, device=dev)
    y = torch.randint(0, 10, (B,), device=dev)
    z = torch.empty(B, 64, device=dev)
    xh = torch.empty(B, 3, 64, 64, device=dev)
    dpre = torch.empty_like(xh)
    logits = torch.empty(B, 10, device=dev)
    dlog = torch.empty(B, 10, device=dev)
    dz1, dz2 = torch.empty(B, 64, device=dev), torch.empty(B, 64, device=dev)
    loss = torch.zeros(4, device=dev)
    st = torch.cuda.Stream()
    sp = lambda: C.c_void_p(st.cuda_stream)
    segs = [
        ("encoder forward", lambda: check(lib.ae_encoder_forward(h, ptr(x), B, 1, ptr(z), sp()))),
        ("decoder forward (+ separate head forward)", lambda: (check(lib.ae_decoder_forward(h, ptr(z), B, 1, ptr(xh), sp())),
                                                               check(lib.ae_head_forward(h, ptr(z), B, ptr(logits), sp())))),
        ("losses (MSE + CE, separate launches)", lambda: (check(lib.ae_sigmoid_mse_fwd_bwd(ptr(xh), ptr(x), xh.numel(), 35.0, ptr(loss), ptr(dpre), sp())),
                                                          check(lib.ae_softmax_ce_fwd_bwd(ptr(logits), ptr(y), B, 10, 1.0, ptr(loss[1:]), ptr(dlog), None, sp())))),
        ("decoder backward (+ head backward)", lambda: (check(lib.ae_decoder_backward(h, ptr(dpre), B, ptr(dz1), sp())),
                                                        check(lib.ae_head_backward(h, ptr(dlog), B, ptr(dz2), sp())))),
        ("encoder backward", lambda: check(lib.ae_encoder_backward(h, ptr(dz1), B, sp()))),
    ]
    tot = 0.0
    for name, fn in segs:
        t = timed_graph(fn, st)
        tot += t
        print(f"{name:45s} {t:8.1f} us")
    flat = eng.flat
    opt = ae_b200.Adam(model.parameters(), lr=5e-3)
    stt = opt.flat_state(flat)

    def adam_pack():
        check(lib.ae_adam_step_flat(ptr(flat.data), ptr(flat.grad), ptr(stt["m"]), ptr(stt["v"]), flat.len, 5e-3, 0.9, 0.999, 1e-8, 0.0, 1.0,
                                    ptr(stt["step"]), sp()))
        for part in (0, 1, 2):
            check(lib.ae_engine_pack_weights(h, part, sp()))
    t = timed_graph(adam_pack, st)
    tot += t
    print(f"{'Adam + weight re-pack (separate launches)':45s} {t:8.1f} us")
    print(f"{'sum of the segments':45s} {tot:8.1f} us")
    stepper = ae_b200.TrainStep(model, opt, 35.0, B)
    stepper.load(x, y)
    for _ in range(5):
        stepper.run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stepper.stream)
    for _ in range(50):
        stepper.run()
    e1.record(stepper.stream)
    torch.cuda.synchronize()
    print(f"{'whole-step graph (TrainStep)':45s} {e0.elapsed_time(e1) * 1e3 / 50:8.1f} us, {stepper.num_kernels} kernels")


if __name__ == "__main__":
    main()
