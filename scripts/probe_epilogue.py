import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ae_b200
from ae_b200 import _lib
lib = _lib.load(); dev = torch.device('cuda', 0)
def run(prec_name, epi_mode, B=256, hs=8, cb=64, cs=128, iters=40):
    prec = _lib.PREC_FP32 if prec_name == 'fp32' else _lib.PREC_BF16
    nsplit = 2 if prec_name == 'fp32' else 1
    M = B * hs * hs
    g = _lib.ConvGeom(B, hs, hs, cb, cs)
    w = torch.randn(cs, cb, 3, 3, device=dev) * 0.05
    nbytes = lib.ae_packed_weight_bytes(cs, cb, prec, 0)
    raw = torch.zeros(2 * nbytes + 2048, dtype=torch.uint8, device=dev)
    base = (raw.data_ptr() + 1023) & ~1023
    pk_f, pk_d = C.c_void_p(base), C.c_void_p((base + nbytes + 1023) & ~1023)
    _lib.check(lib.ae_pack_conv_weight(_lib.ptr(w), cs, cb, pk_f, pk_d, prec, 0, _lib.stream_ptr()))
    bias = torch.zeros(cb, device=dev); stats = torch.zeros(2 * cb, dtype=torch.float64, device=dev)
    n_rot = 7
    planes = [(torch.randn(nsplit * M * cs, device=dev) * 0.5).to(torch.bfloat16) for _ in range(n_rot)]
    outs = [torch.empty(B, 2 * hs, 2 * hs, cb, device=dev) for _ in range(n_rot)]
    ep = _lib.Epilogue(epi_mode, _lib.ptr(bias), None, None, _lib.ptr(stats) if epi_mode else None)
    def launch(i):
        op = _lib.Operand(_lib.ptr(planes[i % n_rot]), None, None, 0.0, _lib.OP_SPLIT_BF16)
        _lib.check(lib.ae_conv2d_s2_dgrad(C.byref(g), C.byref(op), pk_d, C.byref(ep), _lib.ptr(outs[i % n_rot]), prec, 0, _lib.stream_ptr()))
    for i in range(5): launch(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters): launch(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters
for shape in [(8, 64, 128), (16, 32, 64), (4, 128, 256)]:
    for prec in ('fp32', 'bf16'):
        for mode in (0, 1):
            t = run(prec, mode, hs=shape[0], cb=shape[1], cs=shape[2])
            print(f"dgrad hs={shape[0]} cb={shape[1]} cs={shape[2]} {prec} epilogue={'STORE' if mode == 0 else 'BIAS_STATS'}: {t:.1f} us (includes host launch + tensor-map encode)")
