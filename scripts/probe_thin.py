"""Conv2d(3,32) forward: fp32 CUDA-core kernel (thin.cu) vs the tcgen05 variant (thin_tc.cu) at training and inference batch sizes."""
import sys, os, ctypes as C; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ae_b200 import _lib
from tests import gpu_util as gu
lib = gu.lib()
d = gu.dev()
for B in (256, 4096):
    n_rot = 3 if B == 4096 else 12
    xs = [torch.rand(B, 3, 64, 64, device=d) for _ in range(n_rot)]
    w = torch.randn(32, 3, 3, 3, device=d) / 5
    b = torch.zeros(32, device=d)
    outs = [torch.empty(B, 32, 32, 32, device=d) for _ in range(2)]
    stats = torch.zeros(64, dtype=torch.float64, device=d)
    for backend in ("simt", "tc"):
        for prec in ("fp32", "bf16"):
            ep = gu.epilogue(_lib.EPI_BIAS_STATS, b, None, None, stats)
            def run(i):
                op = gu.operand(xs[i % n_rot])
                _lib.check(lib.ae_thin_gather_fwd(C.byref(op), gu.p(w), C.byref(ep), gu.p(outs[i % 2]), B, gu.PREC[prec], gu.BACK[backend], gu.stream()))
            for i in range(3): run(i)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(20): run(i)
            e1.record(); torch.cuda.synchronize()
            print(f"B={B} {backend}/{prec}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per launch", flush=True)
