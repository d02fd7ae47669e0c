#!/usr/bin/env python
"""Small run of every kernel family for compute-sanitizer (scripts/gpu_sanitize.sh): one supervised train step through the
engine (both row-GEMM generations), the optimizer, one eval step, the MLP step, the augmentation kernel and a few direct
C-ABI launches with ragged / multi-tile shapes.  Small batches: the sanitizer slows kernels down 10-100x."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import ae_b200  # noqa: E402
from ae_b200 import _lib  # noqa: E402


def conv_launches(lib, dev, gen):
    os.environ["AE_B200_ROWGEMM"] = gen
    for (b, hs, cb, cs) in ((3, 8, 64, 128), (5, 4, 128, 256), (33, 16, 32, 64), (150, 4, 128, 256)):
        w = torch.randn(cs, cb, 3, 3, device=dev) * 0.05
        P = _lib.PREC_FP32
        nbytes = lib.ae_packed_weight_bytes(cs, cb, P, _lib.BACKEND_TC)
        raw = torch.zeros(2 * nbytes + 2048, dtype=torch.uint8, device=dev)
        base = (raw.data_ptr() + 1023) & ~1023
        pk_f, pk_d = C.c_void_p(base), C.c_void_p((base + nbytes + 1023) & ~1023)
        _lib.check(lib.ae_pack_conv_weight(_lib.ptr(w), cs, cb, pk_f, pk_d, P, _lib.BACKEND_TC, _lib.stream_ptr()))
        g = _lib.ConvGeom(b, hs, hs, cb, cs)
        for fam in ("fwd", "dgrad"):
            if fam == "fwd":
                a_elems, oshape, oc, fn, pk = 4 * b * hs * hs * cb, (b, hs, hs, cs), cs, lib.ae_conv2d_s2_fwd, pk_f
            else:
                a_elems, oshape, oc, fn, pk = b * hs * hs * cs, (b, 2 * hs, 2 * hs, cb), cb, lib.ae_conv2d_s2_dgrad, pk_d
            planes = (torch.randn(2 * a_elems, device=dev) * 0.5).to(torch.bfloat16)
            out = torch.empty(oshape, device=dev)
            y = torch.randn(oshape, device=dev)
            bias, stats, bnc = torch.zeros(oc, device=dev), torch.zeros(2 * oc, dtype=torch.float64, device=dev), torch.ones(8 * oc, device=dev)
            op = _lib.Operand(_lib.ptr(planes), None, None, 0.0, _lib.OP_SPLIT_BF16)
            for ep in (_lib.Epilogue(_lib.EPI_BIAS_STATS, _lib.ptr(bias), None, None, _lib.ptr(stats)),
                       _lib.Epilogue(_lib.EPI_RELUBWD_STATS, None, _lib.ptr(y), _lib.ptr(bnc), _lib.ptr(stats))):
                _lib.check(fn(C.byref(g), C.byref(op), pk, C.byref(ep), _lib.ptr(out), P, _lib.BACKEND_TC, _lib.stream_ptr()))
        torch.cuda.synchronize()
    os.environ.pop("AE_B200_ROWGEMM")


def main():
    dev = torch.device("cuda", 0)
    lib = _lib.load()
    torch.manual_seed(0)
    for prec in ("fp32", "bf16"):
        model = ae_b200.SupervisedAutoencoder(64, 10, precision=prec).to(dev).train()
        x, y = torch.rand(24, 3, 64, 64, device=dev), torch.randint(0, 10, (24,), device=dev)
        model.engine().prepare(dev, 24)
        opt = ae_b200.Adam(model.parameters(), lr=1e-3)
        for _ in range(2):
            loss = model.train_step_grads(x, y, 35.0)
            opt.step()
        st = ae_b200.TrainStep(model, opt, 35.0, 24)
        st(x, y)
        model.eval()
        model.eval_step(x, y, 35.0)
        torch.cuda.synchronize()
        print(prec, "train loss", [round(float(v), 4) for v in loss])
    conv_launches(lib, dev, "1")
    conv_launches(lib, dev, "2")
    clf = ae_b200.MLP(64, 10).to(dev).train()
    clf._state.prepare(dev, 64)
    copt = ae_b200.Adam(clf.parameters(), lr=1e-4, weight_decay=1e-4)
    step = ae_b200.MLPTrainStep(clf, copt, 64)
    step.x.normal_()
    step.run()
    clf.eval()
    clf.predict(torch.randn(100, 64, device=dev))
    imgs = torch.randint(0, 255, (8, 64, 64, 3), dtype=torch.uint8, device=dev)
    ae_b200.augment_u8(imgs, None, torch.tensor([1, 0] * 4), torch.tensor([0, 7] * 4), torch.tensor([8, 3] * 4), 4, None, noise_std=0.03, seed=5)
    torch.cuda.synchronize()
    print("sanitize target done")


if __name__ == "__main__":
    main()
